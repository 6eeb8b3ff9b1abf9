// Tensor-core gathered GEMM for sm_100a: tcgen05.mma kind::tf32 (fp32 operands read as TF32, fp32 accumulation in
// TMEM).  Same GemmParams contract as gemm_fp32.cu -- C[m][n] = sum_k A(m,k) W[n][k] with A gathered through the koff
// table (implicit im2col over zero-bordered channels-last activations) -- plus the fused GRU cell epilogue.
//
// One CTA computes a 128 x BN tile:
//   warps 0-3 : producers.  Per 32-float k-block, 8 consecutive lanes issue the 8 x 16-byte cp.async gathers of one
//               128-byte tile row (a warp instruction touches 4 rows = 4 cache lines) into the canonical K-major
//               SWIZZLE_128B layout (row r at (r/8)*1024 + (r%8)*128, 16-byte chunk j stored at j ^ (r%8)), the same
//               for the weight rows, then arrive on the stage's "full" mbarrier through
//               cp.async.mbarrier.arrive.noinc.  After the k loop the same threads run the epilogue: thread r reads
//               accumulator row r from TMEM (tcgen05.ld 32x32b), applies the epilogue function and parks the row in
//               the idle pipeline stages; each warp then writes its 32 rows out with coalesced 16-byte stores.
//   warp 4    : allocates TMEM; lane 0 issues the tcgen05.mma chain (4 MMAs of K=8 per k-block), commits each stage
//               to its "empty" mbarrier and the last one to the accumulator barrier.
// Every mbarrier wait is bounded (trap after ~2 s) so that a protocol bug cannot hang the GPU.
#include <stdint.h>

#include "se_internal.h"

namespace se {
namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // floats per k-block = one 128-byte swizzle atom
constexpr int A_STAGE_BYTES = BM * BK * 4;
constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
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
    const uint64_t t0 = global_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_ns() - t0 > 2000000000ull) __trap();  // protocol bug: fail loudly instead of hanging the GPU
    }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO(16B)=1 @16 |
// SBO(1024B)=64 @32 | version=1 @46 | layout SWIZZLE_128B=2 @61
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// per-stream (sum, sum of squares) of the rows a warp owns: rows are ordered by stream, so the warp holds a short
// monotone run of stream indices; one shuffle reduction and one pair of double atomics per distinct stream
__device__ __forceinline__ void stats_commit(double* stats, int b, float s, float ss) {
    const unsigned full = 0xffffffffu;
    int lo = b < 0 ? 0x7fffffff : b, hi = b;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo = min(lo, __shfl_xor_sync(full, lo, off));
        hi = max(hi, __shfl_xor_sync(full, hi, off));
    }
    for (int bb = lo; bb <= hi; ++bb) {  // hi < 0 (no valid row): lo = INT_MAX, loop does not run
        float a = b == bb ? s : 0.f, c = b == bb ? ss : 0.f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(full, a, off);
            c += __shfl_xor_sync(full, c, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(stats + 2 * bb, (double)a);
            atomicAdd(stats + 2 * bb + 1, (double)c);
        }
    }
}

constexpr int kW2Floats = 32 * 16 + 32;  // fused small gate: W2 [2*C2][16] + bias2 [2*C2], C2 <= 16

template <int BN, int STAGES>
struct TcSmem {
    static constexpr int B_STAGE_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int TILE_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SP = BN + 4;     // pitch (floats) of the epilogue staging rows
    static constexpr int KOFF_MAX = 512;  // K <= 2048
    static constexpr int OFF_KOFF = TILE_BYTES;
    static constexpr int OFF_BAR = OFF_KOFF + KOFF_MAX * 4;
    static constexpr int OFF_TMEM = OFF_BAR + (2 * STAGES + 1) * 8;
    static constexpr int OFF_AOFF = OFF_TMEM + 8;             // long long [BM] gather base offset of a row (-1: none)
    static constexpr int OFF_OOFF = OFF_AOFF + BM * 8;        // long long [BM] output offset of a row (-1: none)
    static constexpr int OFF_W2 = OFF_OOFF + BM * 8;          // float [kW2Floats]
    static constexpr int BYTES = 1024 /*align slack*/ + OFF_W2 + kW2Floats * 4;
    static_assert(BM * SP * 4 <= TILE_BYTES, "epilogue staging must fit in the pipeline stages");
    static_assert(STAGE_BYTES % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
};

// coalesced write-out of the 32 staging rows a warp owns: consecutive lanes write consecutive float4 of a row
template <int C4, int SP>
__device__ __forceinline__ void store_rows(const float* stg, int col0, const long long* s_ooff, float* out, int cnt,
                                           bool vec4, int warp, int lane) {
    if (vec4) {
        const int cnt4 = (cnt + 3) >> 2;
#pragma unroll 4
        for (int i = lane; i < 32 * C4; i += 32) {
            const int r = warp * 32 + i / C4;
            const int c4 = i % C4;
            const long long off = s_ooff[r];
            if (off >= 0 && c4 < cnt4)
                *reinterpret_cast<float4*>(out + off + 4 * c4) =
                    *reinterpret_cast<const float4*>(stg + r * SP + col0 + 4 * c4);
        }
    } else {
        for (int i = lane; i < 32 * C4 * 4; i += 32) {
            const int r = warp * 32 + i / (C4 * 4);
            const int c = i % (C4 * 4);
            const long long off = s_ooff[r];
            if (off >= 0 && c < cnt) out[off + c] = stg[r * SP + col0 + c];
        }
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads) gemm_tf32_kernel(GemmParams p) {
    using S = TcSmem<BN, STAGES>;
    constexpr int SP = S::SP;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* tiles_ptr = smem_raw + (tiles - raw);
    int* s_koff = reinterpret_cast<int*>(tiles_ptr + S::OFF_KOFF);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(tiles_ptr + S::OFF_BAR);
    // barriers: [0,STAGES) full, [STAGES,2*STAGES) empty, [2*STAGES] accumulator
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tiles_ptr + S::OFF_TMEM);
    long long* s_aoff = reinterpret_cast<long long*>(tiles_ptr + S::OFF_AOFF);
    long long* s_ooff = reinterpret_cast<long long*>(tiles_ptr + S::OFF_OOFF);
    float* s_w2 = reinterpret_cast<float*>(tiles_ptr + S::OFF_W2);
    const uint32_t bar0 = smem_u32(s_bar);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int nkb = (p.K + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

    for (int i = tid; i < p.K / 4; i += kThreads) s_koff[i] = __ldg(p.koff + i);
    if (BN == 16 && p.epi == EPI_ELU_GATE) {
        const int nw = 2 * p.C2 * 16;
        for (int i = tid; i < nw; i += kThreads) s_w2[i] = __ldg(p.W2 + i);
        for (int i = tid; i < 2 * p.C2; i += kThreads) s_w2[32 * 16 + i] = __ldg(p.bias2 + i);
    }
    // row r = tid of the tile: where it gathers from and where it writes to
    int b = -1;
    if (tid < BM) {
        const int m = m0 + tid;
        long long ao = -1, oo = -1;
        if (m < p.M) {
            const int rowsPerStream = p.Tn * p.Fo;
            b = m / rowsPerStream;
            const int rr = m - b * rowsPerStream;
            const int t = rr / p.Fo;
            const int f = rr - t * p.Fo;
            ao = b * p.sB + t * p.sT + f * p.sF;
            oo = b * p.oB + t * p.oT + f * p.oF;
        }
        s_aoff[tid] = ao;
        s_ooff[tid] = oo;
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), kProducerThreads);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);

    if (warp < 4) {
        // ============================ producers ============================
        // 8 consecutive lanes fetch the 8 x 16-byte chunks of one 128-byte tile row (coalesced: a warp instruction
        // touches 4 rows = 4 lines instead of 32); thread (g = tid/8, j = tid%8) serves rows g, g+16, g+32, ...
        const int j = tid & 7;
        const int g = tid >> 3;
        const uint32_t dst_gj = (uint32_t)((g >> 3) * 1024 + (g & 7) * 128 + ((j ^ (g & 7)) << 4));
        const float* arow[BM / 16];
        uint32_t avalid = 0;
#pragma unroll
        for (int i = 0; i < BM / 16; ++i) {
            const long long ao = s_aoff[i * 16 + g];
            arow[i] = p.A + (ao >= 0 ? ao : 0);
            avalid |= (ao >= 0 ? 1u : 0u) << i;
        }
        constexpr int B_ITERS = BN / 16;
        const float* wbase = p.W + (long long)(n0 + g) * p.K + 4 * j;

        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(empty_bar(s), ((kb / STAGES) & 1) ^ 1);
            const uint32_t stage = tiles + (uint32_t)s * S::STAGE_BYTES + dst_gj;
            const int k = kb * BK + 4 * j;
            const bool kin = k < p.K;
            const int ko = kin ? s_koff[k >> 2] : 0;
#pragma unroll
            for (int i = 0; i < BM / 16; ++i)
                cp_async16(stage + (uint32_t)i * 2048u, arow[i] + ko, (kin && ((avalid >> i) & 1u)) ? 16u : 0u);
#pragma unroll
            for (int i = 0; i < B_ITERS; ++i) {
                const bool ok = kin && (n0 + i * 16 + g) < p.Npad;
                cp_async16(stage + A_STAGE_BYTES + (uint32_t)i * 2048u,
                           ok ? wbase + (long long)i * 16 * p.K + kb * BK : p.W, ok ? 16u : 0u);
            }
            cp_async_mbar_arrive_noinc(full_bar(s));
        }

        // ============================ epilogue ============================
        // phase 1: thread r reads accumulator row r from TMEM, applies the epilogue function and parks the result in
        //          the (now idle) pipeline stages; phase 2: each warp writes its 32 rows out with coalesced stores.
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* stg = reinterpret_cast<float*>(tiles_ptr);
        float* srow = stg + tid * SP;
        const bool row_ok = b >= 0;
        float s_acc = 0.f, ss_acc = 0.f;

        if (p.epi == EPI_GRU) {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(srow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            __syncwarp();
            // tile columns: [r | z | n] of hidden units j0 .. j0+U; lane = unit, one row per iteration
            constexpr int U = BN / 3;
            if (lane < U) {
                const int ju = blockIdx.y * U + lane;
                const float* bias = p.bias + (long long)blockIdx.y * BN;
                const float br = __ldg(bias + lane), bz = __ldg(bias + U + lane), bn = __ldg(bias + 2 * U + lane);
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int r = warp * 32 + rr;
                    const int m = m0 + r;
                    if (m < p.M) {
                        const float* gi = p.gi + (long long)m * p.giB;
                        const float hp = p.hprev[(long long)m * p.hB + ju];
                        const float* sr = stg + r * SP;
                        const float rg = sigmoidf_(gi[ju] + sr[lane] + br);
                        const float zg = sigmoidf_(gi[p.H + ju] + sr[U + lane] + bz);
                        const float ng = tanhf(gi[2 * p.H + ju] + rg * (sr[2 * U + lane] + bn));
                        p.out[(long long)m * p.oB + ju] = (1.0f - zg) * ng + zg * hp;
                    }
                }
            }
        } else if (BN == 16 && p.epi == EPI_ELU_GATE) {
            // conv + ELU, then the gated 1x1 pair of CRN_ELU.py:240 in registers (C2 <= 16 channels), + statistics
            float e[16];
            tmem_ld16(trow, e);
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = elu1(e[i] + __ldg(p.bias + i));
            const float* b2 = s_w2 + 32 * 16;
            float y[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float a = 0.f, gt = 0.f;
                if (c < p.C2) {
                    a = b2[2 * c];
                    gt = b2[2 * c + 1];
                    const float4* wa = reinterpret_cast<const float4*>(s_w2 + (2 * c) * 16);
                    const float4* wg = reinterpret_cast<const float4*>(s_w2 + (2 * c + 1) * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 x = wa[q], z = wg[q];
                        a += x.x * e[4 * q] + x.y * e[4 * q + 1] + x.z * e[4 * q + 2] + x.w * e[4 * q + 3];
                        gt += z.x * e[4 * q] + z.y * e[4 * q + 1] + z.z * e[4 * q + 2] + z.w * e[4 * q + 3];
                    }
                    a *= sigmoidf_(gt);
                }
                y[c] = a;
                if (row_ok) {
                    s_acc += a;
                    ss_acc += a * a;
                }
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(srow + i) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
            __syncwarp();
            store_rows<4, SP>(stg, 0, s_ooff, p.out, p.C2, p.vec4 != 0, warp, lane);
            stats_commit(p.stats, b, s_acc, ss_acc);
        } else {
            const bool paired = (p.epi == EPI_GATE_STATS || p.epi == EPI_SKIP);
            const bool want_stats = (p.epi == EPI_ELU_STATS || p.epi == EPI_GATE_STATS || p.epi == EPI_SKIP);
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                const int n = n0 + c0;
                if (n < p.Npad) {  // bias is zero beyond N, weight rows beyond N are zero: padded columns come out 0
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + n + i);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (!paired) {
                    if (p.epi != EPI_BIAS) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = elu1(v[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(srow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            s_acc += v[i];
                            ss_acc += v[i] * v[i];
                        }
                    }
                } else {
                    float w[8];
                    if (p.epi == EPI_GATE_STATS) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = v[2 * i] * sigmoidf_(v[2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = v[2 * i];
                        const bool live = n < p.N;  // elu(0) = 0 for the padded pairs anyway
                        float* s2 = srow + BN / 2 + (c0 >> 1);
                        *reinterpret_cast<float4*>(s2) = make_float4(live ? elu1(v[1]) : 0.f, live ? elu1(v[3]) : 0.f,
                                                                      live ? elu1(v[5]) : 0.f, live ? elu1(v[7]) : 0.f);
                        *reinterpret_cast<float4*>(s2 + 4) =
                            make_float4(live ? elu1(v[9]) : 0.f, live ? elu1(v[11]) : 0.f, live ? elu1(v[13]) : 0.f,
                                        live ? elu1(v[15]) : 0.f);
                    }
                    float* s1 = srow + (c0 >> 1);
                    *reinterpret_cast<float4*>(s1) = make_float4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<float4*>(s1 + 4) = make_float4(w[4], w[5], w[6], w[7]);
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            s_acc += w[i];
                            ss_acc += w[i] * w[i];
                        }
                    }
                }
            }
            __syncwarp();
            const bool vec4 = p.vec4 != 0;
            if (!paired) {
                const int cnt = min(BN, p.N - n0);
                store_rows<BN / 4, SP>(stg, 0, s_ooff, p.out + n0, cnt, vec4, warp, lane);
            } else {
                const int cnt = min(BN / 2, (p.N - n0) >> 1);
                store_rows<BN / 8, SP>(stg, 0, s_ooff, p.out + (n0 >> 1), cnt, vec4, warp, lane);
                if (p.epi == EPI_SKIP) {
                    // out2 shares the row decomposition of out when its strides are equal (the only use: tmp_rm/tmp_rr)
                    store_rows<BN / 8, SP>(stg, BN / 2, s_ooff, p.out2 + (n0 >> 1), cnt, vec4, warp, lane);
                }
            }
            if (want_stats) stats_commit(p.stats, b, s_acc, ss_acc);
        }
        tc_fence_before();
    } else {
        // ============================ MMA issuer ============================
        // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
        if ((tid & 31) == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(full_bar(s), (kb / STAGES) & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async (generic) -> tensor core (async)
                tc_fence_after();
                const uint32_t stage = tiles + (uint32_t)s * S::STAGE_BYTES;
                const uint64_t adesc = make_desc(stage);
                const uint64_t bdesc = make_desc(stage + A_STAGE_BYTES);
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk) {
                    // advance 8 tf32 = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
                    tc_mma_tf32(tmem_base, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc,
                                (kb > 0 || kk > 0) ? 1u : 0u);
                }
                tc_commit(empty_bar(s));
            }
            tc_commit(acc_bar);
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                     : "memory");
    }
}

template <int BN, int STAGES>
int launch_tc(const GemmParams& p, cudaStream_t st) {
    using S = TcSmem<BN, STAGES>;
    static bool configured = false;
    if (!configured) {
        SE_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        S::BYTES));
        configured = true;
    }
    dim3 grid((p.M + BM - 1) / BM, (p.Npad + BN - 1) / BN);
    gemm_tf32_kernel<BN, STAGES><<<grid, kThreads, S::BYTES, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

bool gemm_tf32_supported(const GemmParams& p) {
    if (p.K % 4 != 0 || p.K < 8 || p.K > 2048) return false;
    if (p.epi == EPI_GRU) return p.H % 32 == 0;
    if (p.epi == EPI_ELU_GATE) return p.Npad == 16 && p.C2 >= 1 && p.C2 <= 16 && p.W2 && p.bias2;
    if (p.epi == EPI_SKIP && (p.o2B != p.oB || p.o2T != p.oT || p.o2F != p.oF)) return false;
    if (p.vec4 && ((p.oF % 4) || (p.oT % 4) || (p.oB % 4))) return false;
    return p.N >= 1;
}

int launch_gemm_tf32(const GemmParams& p, cudaStream_t st) {
    SE_REQUIRE(gemm_tf32_supported(p), "gemm_tf32: unsupported shape");
    if (p.M <= 0) return 0;
    if (p.epi == EPI_GRU) return launch_tc<96, 4>(p, st);
    if (p.Npad <= 16) return launch_tc<16, 4>(p, st);
    if (p.Npad <= 32) return launch_tc<32, 4>(p, st);
    if (p.Npad <= 64) return launch_tc<64, 4>(p, st);
    if (p.Npad <= 128 || p.Npad % 256 != 0) return launch_tc<128, 3>(p, st);
    return launch_tc<256, 3>(p, st);
}

}  // namespace se
