// Persistent tensor-core GRU recurrence for the streaming path (CRN_ELU.py:173, nn.GRU; fp16 operand mode): ONE launch
// per layer walks all T = 21 time steps of a chunk.  CTA (m-tile, n-tile) owns 128 streams x 32 hidden units for the
// whole chunk and keeps its slice of W_hh (96 gate rows x H, fp16, SWIZZLE_128B K-major) RESIDENT in shared memory;
// per step it only streams the 128 x H fp16 rows of h_{t-1} through a cp.async ring, issues the tcgen05.mma chain into
// a TMEM accumulator and runs the fused cell epilogue (gi prefetched while the MMAs run, fp32 master state updated in
// place, fp16 copy = next step's operand).  The H/32 CTAs that share an m-tile exchange h_t through L2 and a release /
// acquire counter; m-tile groups are independent, so the grid needs no cooperative launch (complete groups always make
// progress and free their SMs).
//
// Replaces 21 dependent launches per layer of the generic GEMM with the EPI_GRU epilogue (12.9 us each: weights
// re-fetched from L2, TMEM / barrier set-up and launch latency per step; profiles/r01_ncu_full_gru_step_fp16.txt).
// Weight packing is the one of EPI_GRU: row n of W_hh' = gate (n % 96) / 32 of hidden unit 32 (n / 96) + n % 32.
#include <cuda_fp16.h>
#include <stdint.h>

#include "se_internal.h"

namespace se {
namespace {

constexpr int BM = 128;
constexpr int BN = 96;            // [r | z | n] of 32 hidden units
constexpr int KB_HALVES = 64;     // one 128-byte swizzle atom row
constexpr int A_STAGE = BM * 128; // bytes
constexpr int W_KB = BN * 128;    // bytes of the resident weight slice per k-block
constexpr int STAGES = 6;  // 6 x 16 KB of h rows in flight next to the 96 KB resident weight slice (H = 512)
constexpr int kThreads = 9 * 32;  // 4 producer warps, 1 MMA warp, 4 epilogue warps

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
// K-major SWIZZLE_128B shared-memory matrix descriptor (as gemm_tc.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// ex2.approx.ftz / rcp.approx.ftz: one MUFU each.  __expf / __fdividef (and exp2f) add a range check and a rescale for
// denormal results around their MUFU -- about ten instructions per sigmoid instead of four.  Flushed denormals are exact
// zeros here: sigmoid -> 0 / 1, elu -> -1, tanh -> +-1 at the ends of their ranges.
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - 2.0f * rcp_ftz(1.0f + ex2_ftz(2.8853900817779268f * x)); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kThreads, 1) gru_tc_persist_kernel(GruTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* base_ptr = smem_raw + (base - raw);
    const int nkb = p.H / KB_HALVES;
    const uint32_t w_smem = base;                              // nkb x 12 KB resident weights
    const uint32_t a_smem = base + (uint32_t)nkb * W_KB;       // STAGES x 16 KB ring of h rows
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(base_ptr + nkb * W_KB + STAGES * A_STAGE);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * STAGES + 2);
    float* s_bias = reinterpret_cast<float*>(s_tmem + 2);      // [96] b_hh of this tile (packed order)
    const uint32_t bar0 = smem_u32(s_bar);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES), tempty_bar = bar0 + 8u * (2 * STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NT = p.H / 32;
    const int mt = blockIdx.x / NT, nt = blockIdx.x % NT;
    const int m0 = mt * BM, n0 = nt * BN, j0 = nt * 32;
    const __half* W = p.Whh;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 128);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < BN) s_bias[tid] = __ldg(p.bhh + n0 + tid);
    // resident weight slice: rows n0 .. n0+95, all of K; chunk j of row r at (r/8)*1024 + (r%8)*128 + ((j ^ (r%8)) << 4)
    for (int u = tid; u < nkb * BN * 8; u += kThreads) {
        const int j = u & 7, r = (u >> 3) % BN, kb = u / (8 * BN);
        const uint32_t dst = w_smem + (uint32_t)kb * W_KB + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
        cp_async16(dst, W + (long long)(n0 + r) * p.Kp + kb * KB_HALVES + 8 * j, 16u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);

    if (warp < 4) {
        // ============================ producers: h_{t-1} rows of this m-tile ============================
        const int j = tid & 7, g = tid >> 3;  // chunk j of rows g, g+16, ...
        const uint32_t dst_gj = (uint32_t)((g >> 3) * 1024 + (g & 7) * 128 + ((j ^ (g & 7)) << 4));
        uint32_t ps = 0, pphase = 0;
        const int* counter = p.counters + mt;
        for (int t = 0; t < p.T; ++t) {
            if (t > 0) {  // every CTA of this m-tile has published its 32 units of h_t (release: threadfence + atomic)
                if (lane == 0) {
                    uint64_t t0 = 0;
                    uint32_t spins = 0;
                    while (ld_acquire(counter) < NT * t) {
                        __nanosleep(64);
                        if ((++spins & 4095u) == 0) {
                            const uint64_t now = global_ns();
                            if (t0 == 0) t0 = now;
                            else if (now - t0 > 2000000000ull) __trap();
                        }
                    }
                }
                __syncwarp();
            }
            const __half* arow[BM / 16];
            uint32_t asz[BM / 16];
#pragma unroll
            for (int i = 0; i < BM / 16; ++i) {
                const int m = m0 + g + 16 * i;
                const bool ok = m < p.B;
                arow[i] = p.hseq + (long long)(ok ? m : 0) * p.hB + (long long)t * p.H + 8 * j;
                asz[i] = ok ? 16u : 0u;
            }
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(empty_bar(ps), pphase ^ 1u);
                const uint32_t stage = a_smem + (uint32_t)ps * A_STAGE + dst_gj;
#pragma unroll
                for (int i = 0; i < BM / 16; ++i) cp_async16(stage + (uint32_t)i * 2048u, arow[i] + kb * KB_HALVES, asz[i]);
                cp_async_mbar_arrive_noinc(full_bar(ps));
                if (++ps == STAGES) {
                    ps = 0;
                    pphase ^= 1u;
                }
            }
        }
    } else if (warp == 4) {
        // ============================ MMA issuer ============================
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);  // f16 x f16 -> f32
        if (lane == 0) {
            uint32_t ms = 0, mphase = 0;
            for (int t = 0; t < p.T; ++t) {
                mbar_wait(tempty_bar, (uint32_t)((t & 1) ^ 1));  // the epilogue of step t-1 has drained the accumulator
                tc_fence_after();
                uint32_t acc = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(full_bar(ms), mphase);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tc_fence_after();
                    const uint64_t adesc = make_desc(a_smem + ms * (uint32_t)A_STAGE);
                    const uint64_t bdesc = make_desc(w_smem + (uint32_t)kb * W_KB);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        tc_mma_f16(tmem_base, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc, acc);
                        acc = 1;
                    }
                    tc_commit(empty_bar(ms));
                    if (++ms == STAGES) {
                        ms = 0;
                        mphase ^= 1u;
                    }
                }
                tc_commit(tfull_bar);
            }
        }
        __syncwarp();
    } else {
        // ============================ epilogue: thread = stream row ============================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int m = m0 + q * 32 + lane;
        const bool ok = m < p.B;
        const long long b = ok ? m : 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
        float* h32 = p.h32 + b * p.H + j0;
        // fp32 master state of this row's 32 units: lives in registers for the whole chunk
        float4 ph[8];
#pragma unroll
        for (int v4 = 0; v4 < 8; ++v4)
            ph[v4] = ok ? *(reinterpret_cast<const float4*>(h32) + v4) : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t < p.T; ++t) {
            const float* gi = p.gi + b * p.giB + (long long)t * 3 * p.H + j0;
            __half* hout = p.hseq + b * p.hB + (long long)(t + 1) * p.H + j0;
            // the cell's other inputs do not depend on the MMAs: fetch all of them (4 x 128 bytes of this row) while the
            // operand loads and the MMA chain of this step run
            float4 pr[8], pz[8], pn[8];
#pragma unroll
            for (int v4 = 0; v4 < 8; ++v4) {
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                pr[v4] = ok ? __ldg(reinterpret_cast<const float4*>(gi) + v4) : z4;
                pz[v4] = ok ? __ldg(reinterpret_cast<const float4*>(gi + p.H) + v4) : z4;
                pn[v4] = ok ? __ldg(reinterpret_cast<const float4*>(gi + 2 * p.H) + v4) : z4;
            }
            mbar_wait(tfull_bar, (uint32_t)(t & 1));
            tc_fence_after();
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                const int u0 = 8 * c8;
                uint32_t v[24];
                tmem_ld8_nowait(tlane + u0, v);
                tmem_ld8_nowait(tlane + 32 + u0, v + 8);
                tmem_ld8_nowait(tlane + 64 + u0, v + 16);
                tmem_ld_wait();
                const float gr[8] = {pr[2 * c8].x, pr[2 * c8].y, pr[2 * c8].z, pr[2 * c8].w,
                                     pr[2 * c8 + 1].x, pr[2 * c8 + 1].y, pr[2 * c8 + 1].z, pr[2 * c8 + 1].w};
                const float gz[8] = {pz[2 * c8].x, pz[2 * c8].y, pz[2 * c8].z, pz[2 * c8].w,
                                     pz[2 * c8 + 1].x, pz[2 * c8 + 1].y, pz[2 * c8 + 1].z, pz[2 * c8 + 1].w};
                const float gn[8] = {pn[2 * c8].x, pn[2 * c8].y, pn[2 * c8].z, pn[2 * c8].w,
                                     pn[2 * c8 + 1].x, pn[2 * c8 + 1].y, pn[2 * c8 + 1].z, pn[2 * c8 + 1].w};
                const float hp[8] = {ph[2 * c8].x, ph[2 * c8].y, ph[2 * c8].z, ph[2 * c8].w,
                                     ph[2 * c8 + 1].x, ph[2 * c8 + 1].y, ph[2 * c8 + 1].z, ph[2 * c8 + 1].w};
                float hn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float rg = fast_sigmoid(gr[i] + __uint_as_float(v[i]) + s_bias[u0 + i]);
                    const float zg = fast_sigmoid(gz[i] + __uint_as_float(v[8 + i]) + s_bias[32 + u0 + i]);
                    const float ng = fast_tanh(gn[i] + rg * (__uint_as_float(v[16 + i]) + s_bias[64 + u0 + i]));
                    hn[i] = (1.0f - zg) * ng + zg * hp[i];
                }
                ph[2 * c8] = make_float4(hn[0], hn[1], hn[2], hn[3]);
                ph[2 * c8 + 1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
                if (ok) {
                    const __half2 h0 = __floats2half2_rn(hn[0], hn[1]), h1 = __floats2half2_rn(hn[2], hn[3]),
                                  h2 = __floats2half2_rn(hn[4], hn[5]), h3 = __floats2half2_rn(hn[6], hn[7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<const uint32_t*>(&h2);
                    pk.w = *reinterpret_cast<const uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(hout + u0) = pk;
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar);  // accumulator drained: the MMAs of step t+1 may overwrite it
            // publish h_t: the named barrier orders the 128 rows' stores before the single release at gpu scope
            // (cumulative), which the producers of the whole m-tile group acquire
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == 5 && lane == 0)
                asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.counters + mt) : "memory");
        }
        if (ok) {
#pragma unroll
            for (int v4 = 0; v4 < 8; ++v4) reinterpret_cast<float4*>(h32)[v4] = ph[v4];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

}  // namespace

size_t gru_tc_smem_bytes(int H) { return 1024 + (size_t)(H / KB_HALVES) * W_KB + STAGES * A_STAGE + 8 * (2 * STAGES + 2) + 16 + BN * 4; }

bool gru_tc_persist_supported(int H) { return H % 64 == 0 && H >= 64 && gru_tc_smem_bytes(H) <= 227 * 1024; }

int launch_gru_tc_persist(const GruTcParams& p, cudaStream_t st) {
    SE_REQUIRE(gru_tc_persist_supported(p.H), "gru_tc_persist: hidden size must be a multiple of 64 that fits shared memory");
    if (p.B <= 0) return 0;
    const size_t smem = gru_tc_smem_bytes(p.H);
    SE_DYN_SMEM(gru_tc_persist_kernel, smem);
    // The CTAs of an m-tile group wait for one another (release / acquire counter in L2), so every group of a launch must
    // be resident at once: one CTA per SM (shared memory), hence at most floor(SMs / (H / 32)) m-tiles per launch.  More
    // streams run as successive launches ("waves") over slices of the batch instead of relying on the dispatch order.
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    const int ntn = p.H / 32;
    const int tiles_per_wave = num_sms / ntn;
    SE_REQUIRE(tiles_per_wave >= 1, "gru_tc_persist: more n-tiles than SMs");
    const int mtiles = (p.B + BM - 1) / BM;
    SE_CUDA_OK(cudaMemsetAsync(p.counters, 0, sizeof(int) * mtiles, st));
    for (int m0 = 0; m0 < mtiles; m0 += tiles_per_wave) {
        GruTcParams w = p;
        const int b0 = m0 * BM;
        const int nt = mtiles - m0 < tiles_per_wave ? mtiles - m0 : tiles_per_wave;
        w.B = (p.B - b0) < nt * BM ? (p.B - b0) : nt * BM;
        w.gi = p.gi + (long long)b0 * p.giB;
        w.hseq = p.hseq + (long long)b0 * p.hB;
        w.h32 = p.h32 + (long long)b0 * p.H;
        w.counters = p.counters + m0;
        gru_tc_persist_kernel<<<nt * ntn, kThreads, smem, st>>>(w);
        SE_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace se
