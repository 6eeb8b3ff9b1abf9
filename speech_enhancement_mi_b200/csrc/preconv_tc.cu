// Pre-convolution block (CRN_ELU.py:337-339,375-376) on the tensor cores, fp16 operand mode: 5x5 frequency-dilated
// causal conv (5 -> 5 channels) + ELU + gated 1x1 + GlobalLayerNorm + residual + causal state roll, one persistent CTA
// per stream at a time.
//
// No im2col is ever materialised.  The whole zero-bordered input of the chunk sits in shared memory channels-last,
// X[25 frames][272 positions][8 halves] (one 16-byte unit per (frame, bin); bin f at position f + 2d), and the
// convolution is split as   out[t][f][co] = sum_kf Z_t[f + kf d][kf][co],   Z_t[pos][kf][co] = sum_{kt,ci} w[co][ci][kf][kt] X[t+kt][pos][ci]:
//   * Z_t is a plain GEMM over the UNSHIFTED positions (M = 256 positions, K = 5 frames x 8 channels, N = 5 taps x 8
//     channels = 48 with padding).  In the canonical K-major no-swizzle UMMA layout -- ((8, n), 2):((16 B, SBO), LBO) --
//     eight consecutive units of a frame row are one core matrix, SBO = 128 B walks the 8-row groups and LBO, the distance
//     between the two 16-byte K chunks of one MMA, is the ROW PITCH: one tcgen05.mma (M128 x N48 x K16) consumes two
//     frames.  3 MMAs per 128 positions, 126 per stream and layer (a first version with the taps in K needed 1092 MMAs
//     of N = 16 and was MMA-issue bound: ~100 cycles per instruction whatever its size).
//   * the frequency taps become a shift-and-add in the epilogue: Z_t goes TMEM -> registers -> shared memory, and the
//     thread of bin f adds the five rows f, f + d, ..., f + 4d.
// GlobalLayerNorm needs the statistics of the whole stream before anything can be written: the 4221 x 5 gated values wait
// in shared memory as fp16 planes; a second sweep normalises, adds the block input and stores the next layer's units.
#include <cuda_fp16.h>
#include <stdint.h>

#include "se_internal.h"

namespace se {
namespace {

constexpr int T = kFramesPerChunk;  // 21
constexpr int NB = 201;
constexpr int TP = T + 4;           // 25 frames: 4 carried + 21 new
constexpr int FPOS = PRECONV_TC_POS;  // 272 positions per frame row
constexpr int ROW_BYTES = FPOS * 16;
constexpr int X_BYTES = TP * ROW_BYTES;  // 108,800
constexpr int ZCOLS = 40;                // Z row: [kf][8 channels], 5 taps
constexpr int ZPITCH = 56;               // halves per Z row: 7 x 16 B, so that 8 consecutive rows hit 8 different bank groups
constexpr int ZBUF = 256 * ZPITCH;       // halves per buffer
constexpr int Z_BYTES = 2 * ZBUF * 2;    // 57,344: fp16, double-buffered over the frame parity (one barrier per frame)
constexpr int YPITCH = 208;              // bins per (channel, frame) row of the fp16 planes
constexpr int Y_BYTES = 5 * T * YPITCH * 2;  // 43,680
constexpr int NPAIR = 3;                 // frame pairs (0,1) (2,3) (4,-)
constexpr int WPAIR_BYTES = 2 * 6 * 128; // B tile of a pair: [k chunk (768 B)][n group of 8 rows (128 B)][8 rows x 16 B]
constexpr int W_BYTES = NPAIR * WPAIR_BYTES;
constexpr int NTILE = 2 * T;             // 2 tiles of 128 positions per frame
constexpr int kThreads = 10 * 32;        // warp 0: MMA issuer; warps 2..9: epilogue; all: loads and the final sweep
constexpr int kEpiWarp0 = 2;
constexpr int NACC = 8;                  // TMEM accumulators (64-column slots, 48 used): the MMA warp runs 4 frames ahead

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0 = 0;
    for (uint32_t spins = 1;; ++spins) {
        if (mbar_try_wait(bar, parity)) return;
        if ((spins & 1023u) == 0) {  // protocol bug: fail loudly (after ~2 s) instead of hanging the GPU
            const uint64_t t = global_ns();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 2000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, no swizzle (cute::UMMA::LayoutType::SWIZZLE_NONE = 0): start >> 4 | LBO >> 4 @16 | SBO >> 4 @32 | version 1 @46
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_elu(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }

__device__ __forceinline__ void tmem_ld8_nowait8(uint32_t taddr, float* r) {
    uint32_t v[8];
    tmem_ld8_nowait(taddr, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(v[i]);
}

__global__ void __launch_bounds__(kThreads, 1) preconv_tc_kernel(PreconvTcParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sx = smem;                                       // X
    __half* sz = reinterpret_cast<__half*>(smem + X_BYTES);         // Z of two frames [2][256][56] fp16
    __half* sy = reinterpret_cast<__half*>(smem + X_BYTES + Z_BYTES);  // gated values [5][21][208]
    unsigned char* swt = smem + X_BYTES + Z_BYTES + Y_BYTES;        // B tiles
    float* sp = reinterpret_cast<float*>(swt + W_BYTES);            // packed fp32 parameters (bias, gate, norm)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sp + PRECONV_W_FLOATS);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * NACC);
    double* s_red = reinterpret_cast<double*>(s_tmem + 2);  // [2][8 warps]
    float* s_co = reinterpret_cast<float*>(s_red + 16);     // mean, inv
    const uint32_t x_smem = smem_u32(sx), w_smem = smem_u32(swt), bar0 = smem_u32(s_bar);
    auto tfull = [&](uint32_t a) { return bar0 + 8u * a; };
    auto tempty = [&](uint32_t a) { return bar0 + 8u * (NACC + a); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = p.d;

    // ---- one-time set-up: parameters, B tiles (fp16, canonical no-swizzle K-major), barriers, TMEM ----------------------
    for (int i = tid; i < PRECONV_W_FLOATS; i += kThreads) sp[i] = __ldg(p.w + i);
    __syncthreads();
    // B[pair][kc][n][k]: chunk kc <-> frame tap kt = 2 pair + kc; row n = kf * 8 + co; k = input channel ci
    for (int u = tid; u < NPAIR * 2 * 48; u += kThreads) {
        const int n = u % 48, kc = (u / 48) & 1, pair = u / 96;
        const int kt = 2 * pair + kc, kf = n >> 3, co = n & 7;
        __align__(16) __half h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = 0.f;
            if (kt < 5 && kf < 5 && co < 5 && k < 5) v = sp[(kt * 5 + k) * 28 + kf * 5 + co];  // packed [(kt*5+ci)*28 + kf*5 + co]
            h[k] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4*>(swt + pair * WPAIR_BYTES + kc * 768 + (n >> 3) * 128 + (n & 7) * 16) =
            *reinterpret_cast<const uint4*>(h);
    }
    if (tid == 0) {
        for (int a = 0; a < NACC; ++a) {
            mbar_init(tfull(a), 1);
            mbar_init(tempty(a), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);

    uint32_t tile_it = 0;  // tiles issued / consumed so far by this CTA, same count in every role
    for (int stream = blockIdx.x; stream < p.B; stream += gridDim.x) {
        const int b = p.b0 + stream;
        __half* gx = p.in + (long long)b * p.in_sB;
        // ---- the chunk's input: 25 frames x 272 units, contiguous --------------------------------------------------------
        for (int i = tid; i < X_BYTES / 16; i += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(x_smem + 16u * i), "l"(reinterpret_cast<const uint4*>(gx) + i)
                         : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        // causal state: the last 4 frames of this chunk's input become frames 0..3 of the next chunk (CRN_ELU.py:246)
        for (int i = tid; i < 4 * FPOS; i += kThreads)
            reinterpret_cast<uint4*>(gx)[i] = reinterpret_cast<const uint4*>(sx + T * ROW_BYTES)[i];

        if (warp == 0) {
            // ============================ MMA issuer: Z of every frame, two tiles of 128 positions ============================
            if (lane == 0) {
                constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                for (int tile = 0; tile < NTILE; ++tile) {
                    const uint32_t it = tile_it + tile, acc = it % NACC;
                    mbar_wait(tempty(acc), ((it / NACC) & 1) ^ 1);
                    tc_fence_after();
                    const int t = tile >> 1, pos0 = (tile & 1) * 128;
#pragma unroll
                    for (int pair = 0; pair < NPAIR; ++pair) {
                        const uint32_t offa = (uint32_t)(((t + 2 * pair) * FPOS + pos0) * 16);
                        const uint32_t lbo = pair < 2 ? (uint32_t)ROW_BYTES : 16u;  // (4,-): the second chunk has zero weights
                        tc_mma_f16(tmem_base + acc * 64, make_desc_ns(x_smem + offa, lbo, 128u),
                                   make_desc_ns(w_smem + pair * WPAIR_BYTES, 768u, 128u), idesc, pair ? 1u : 0u);
                    }
                    tc_commit(tfull(acc));
                }
            }
            __syncwarp();
        } else if (warp >= kEpiWarp0) {
            // ============================ epilogue: 256 threads ============================
            const int ew = warp - kEpiWarp0, g = ew >> 2, q = warp & 3;
            const int pos = 128 * g + 32 * q + lane;  // Z row this thread moves out of TMEM
            const int f = ew * 32 + lane;             // output bin this thread finishes
            float psum = 0.f, psq = 0.f;
            // the 65 scalars of the cell (conv bias, the two 5x5 gate matrices and their biases) live in registers: read
            // from shared memory they were 70 dependent loads per output and half of the kernel's time
            float cb[5], wt[25], wg[25], bt[5], bg[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                cb[i] = sp[PRECONV_W_BIAS + i];
                bt[i] = sp[PRECONV_W_BT + i];
                bg[i] = sp[PRECONV_W_BG + i];
            }
#pragma unroll
            for (int i = 0; i < 25; ++i) {
                wt[i] = sp[PRECONV_W_WT + i];
                wg[i] = sp[PRECONV_W_WG + i];
            }
            for (int t = 0; t < T; ++t) {
                const uint32_t it = tile_it + 2 * t + g, acc = it % NACC;
                mbar_wait(tfull(acc), (it / NACC) & 1);
                tc_fence_after();
                float z[ZCOLS];
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64;
#pragma unroll
                for (int kf = 0; kf < 5; ++kf) tmem_ld8_nowait8(ta + 8 * kf, z + 8 * kf);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(tempty(acc));
                uint4* zr = reinterpret_cast<uint4*>(sz + (t & 1) * ZBUF + pos * ZPITCH);
#pragma unroll
                for (int kf = 0; kf < 5; ++kf) {
                    const __half2 h0 = __floats2half2_rn(z[8 * kf], z[8 * kf + 1]), h1 = __floats2half2_rn(z[8 * kf + 2], z[8 * kf + 3]),
                                  h2 = __floats2half2_rn(z[8 * kf + 4], 0.f);
                    uint4 u;
                    u.x = *reinterpret_cast<const unsigned*>(&h0);
                    u.y = *reinterpret_cast<const unsigned*>(&h1);
                    u.z = *reinterpret_cast<const unsigned*>(&h2);
                    u.w = 0u;
                    zr[kf] = u;
                }
                // Z of frame t complete.  One barrier per frame is enough: frame t+2 overwrites this buffer only after the
                // barrier of frame t+1, which every thread reaches after it has finished reading frame t
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (f < NB) {
                    float e[5];
#pragma unroll
                    for (int c = 0; c < 5; ++c) e[c] = cb[c];
#pragma unroll
                    for (int kf = 0; kf < 5; ++kf) {  // out[f] = sum_kf Z[f + kf d][kf]
                        const uint4 u = *reinterpret_cast<const uint4*>(sz + (t & 1) * ZBUF + (f + kf * d) * ZPITCH + 8 * kf);
                        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
                        const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
                        const float2 c4 = __half22float2(*reinterpret_cast<const __half2*>(&u.z));
                        e[0] += a.x;
                        e[1] += a.y;
                        e[2] += c2.x;
                        e[3] += c2.y;
                        e[4] += c4.x;
                    }
#pragma unroll
                    for (int c = 0; c < 5; ++c) e[c] = fast_elu(e[c]);
#pragma unroll
                    for (int co = 0; co < 5; ++co) {
                        float a = bt[co], gt = bg[co];
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            a = fmaf(wt[co * 5 + k], e[k], a);
                            gt = fmaf(wg[co * 5 + k], e[k], gt);
                        }
                        const float y = a * fast_sigmoid(gt);
                        psum += y;
                        psq += y * y;
                        sy[(co * T + t) * YPITCH + f] = __float2half_rn(y);
                    }
                }
            }
            // GlobalLayerNorm statistics of this stream (CRN_ELU.py:40-41), reduced in double
            double ds = psum, dq = psq;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                ds += __shfl_xor_sync(0xffffffffu, ds, off);
                dq += __shfl_xor_sync(0xffffffffu, dq, off);
            }
            if (lane == 0) {
                s_red[ew] = ds;
                s_red[8 + ew] = dq;
            }
        }
        tile_it += NTILE;
        __syncthreads();
        if (tid == 0) {
            double ds = 0.0, dq = 0.0;
            for (int w = 0; w < 8; ++w) {
                ds += s_red[w];
                dq += s_red[8 + w];
            }
            const double count = 5.0 * NB * T;
            const double mu = ds / count;
            double var = dq / count - mu * mu;
            if (var < 0.0) var = 0.0;
            const float varf = (float)var;
            const float den = p.student ? (sqrtf(varf) + 1e-8f) : (sqrtf(varf + 1e-8f) + 1e-8f);
            s_co[0] = (float)mu;
            s_co[1] = 1.0f / den;
        }
        __syncthreads();
        // ---- normalise, add the block input (CRN_ELU.py:376) and write the next layer's input units ---------------------------
        {
            const float mean = s_co[0], inv = s_co[1];
            for (int i = tid; i < T * NB; i += kThreads) {
                const int t = i / NB, f = i - t * NB;
                const __half* xin = reinterpret_cast<const __half*>(sx + ((t + 4) * FPOS + f + 2 * d) * 16);
                float o[5];
#pragma unroll
                for (int c = 0; c < 5; ++c)
                    o[c] = (__half2float(sy[(c * T + t) * YPITCH + f]) - mean) * inv * sp[PRECONV_W_NW + c] + sp[PRECONV_W_NB + c] +
                           __half2float(xin[c]);
                const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]), h2 = __floats2half2_rn(o[4], 0.f);
                uint4 u;
                u.x = *reinterpret_cast<const unsigned*>(&h0);
                u.y = *reinterpret_cast<const unsigned*>(&h1);
                u.z = *reinterpret_cast<const unsigned*>(&h2);
                u.w = 0u;
                *reinterpret_cast<uint4*>(p.out + (long long)b * p.oB + (long long)t * p.oT + (long long)f * p.oF) = u;
            }
        }
        __syncthreads();  // everyone is done with X and the planes before the next stream overwrites them
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

constexpr size_t kSmemBytes = X_BYTES + Z_BYTES + Y_BYTES + W_BYTES + PRECONV_W_FLOATS * 4 + 2 * NACC * 8 + 8 + 16 * 8 + 16;

}  // namespace

int launch_preconv_tc(const PreconvTcParams& p, cudaStream_t st) {
    if (p.B <= 0) return 0;
    SE_REQUIRE(p.d == 1 || p.d == 2 || p.d == 4, "preconv_tc: frequency dilation must be 1, 2 or 4 (CRN_ELU.py:336)");
    SE_DYN_SMEM(preconv_tc_kernel, kSmemBytes);
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    preconv_tc_kernel<<<p.B < num_sms ? p.B : num_sms, kThreads, kSmemBytes, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace se
