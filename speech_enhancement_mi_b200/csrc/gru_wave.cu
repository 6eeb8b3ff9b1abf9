// Wavefront GRU for the streaming path (CRN_ELU.py:173, nn.GRU(num_layers=2); fp16 operand mode): ONE launch walks
// both layers through all T = 21 steps of a chunk, layer 1 one step behind layer 0.
//
// Layer 1's input at step t is h0_t alone, so its input projection does not have to wait for the whole layer-0
// sequence: the serial chain of the two recurrences (2 T dependent steps, plus the 768-column projection GEMM between
// them) becomes T + 1 steps.  CTA (layer, m-tile, n-tile) owns 128 streams x 32 hidden units for the whole chunk and
// keeps its weight slices RESIDENT in shared memory: layer 0 the 96 gate rows [r | z | n] of W_hh0, layer 1 those of
// W_ih1 AND W_hh1 (2 x 48 KB at H = 256).  Per step a layer-1 CTA streams the 128 x H fp16 rows of h0_t and of
// h1_{t-1} (TMA boxes of the two state histories), accumulates r and z over both (K = 2 H) and keeps the two halves
// of the n gate apart -- n = tanh(W_in h0 + b_in + r (W_hn h1 + b_hn)) -- in TMEM columns 64..95 / 96..127, then runs
// the fused cell epilogue (fp32 master state in registers for the whole chunk, fp16 copy = the operand of later steps
// and of the fc GEMM).  Layer-0 CTAs are the persistent recurrence of gru_tc_persist.cu with the same roles.  The
// H/32 CTAs of an m-tile group publish h_t through L2 and one release / acquire counter per (layer, m-tile); the
// layer-1 group of an m-tile also acquires the layer-0 counter.  All CTAs of a launch are resident at once (one per
// SM, 2 x H/32 per m-tile), so the dependency graph always makes progress; more streams run as successive launches.
//
// Roles: warp 0 = TMA producer (counter acquire + tile loads), warp 1 = MMA issuer (warp-uniform loop, one elected
// lane issues), warps 2..9 = epilogue (two warps per TMEM lane quarter, 16 hidden units each).
#include <cuda_fp16.h>
#include <stdint.h>

#include "se_internal.h"

#ifndef SE_GRU_PROFILE
#define SE_GRU_PROFILE 0  // -DSE_GRU_PROFILE=1: the roles account their wait cycles per layer (se_debug_gru_counters)
#endif

namespace se {
namespace {

constexpr int BM = 128;
constexpr int BN = 96;              // [r | z | n] of 32 hidden units
constexpr int KB_HALVES = 64;       // one 128-byte swizzle atom row
constexpr int A_STAGE = BM * 128;   // bytes of one k-block of state rows
constexpr int W_KB = BN * 128;      // bytes of one k-block of a resident weight slice
constexpr int STAGES = 6;
constexpr int kEpiWarps = 8;
constexpr int kFirstEpi = 2;
constexpr int kThreads = (kFirstEpi + kEpiWarps) * 32;
constexpr bool kProf = SE_GRU_PROFILE != 0;

// [layer][0] producer waiting for the published state (counter), [1] producer waiting for a free stage, [2] MMA warp
// waiting for operands, [3] MMA warp total, [4] epilogue lead warp waiting for the accumulator, [5] epilogue total,
// [6] MMA warp waiting for the drained accumulator, [7] steps
__device__ unsigned long long g_gru_prof[2][8];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
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
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
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
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_ftz(1.0f + ex2_ftz(-kLog2e * x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - 2.0f * rcp_ftz(1.0f + ex2_ftz(2.0f * kLog2e * x)); }
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// whole warp: lane 0 spins until *counter >= target (acquire), the warp barrier hands the ordering to the other lanes,
// and every lane fences towards the async proxy (whichever lane is elected issues the tile loads that follow)
__device__ __forceinline__ void wait_counter(const int* counter, int target, int lane) {
    if (lane == 0) {
        uint64_t t0 = 0;
        uint32_t spins = 0;
        while (ld_acquire(counter) < target) {
            if ((++spins & 4095u) == 0) {
                const uint64_t now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 2000000000ull) __trap();
            }
        }
    }
    __syncwarp();
    asm volatile("fence.proxy.async;" ::: "memory");
}

__global__ void __launch_bounds__(kThreads, 1) gru_wave_kernel(const __grid_constant__ GruWaveParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* base_ptr = smem_raw + (base - raw);
    const int nkb = p.H / KB_HALVES;
    const int wslots = p.layers * nkb;                            // resident weight k-blocks (12 KB each)
    const uint32_t w_smem = base;
    const uint32_t a_smem = base + (uint32_t)wslots * W_KB;       // STAGES x 16 KB ring of state rows
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(base_ptr + wslots * W_KB + STAGES * A_STAGE);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * STAGES + 2);
    float* s_bias = reinterpret_cast<float*>(s_tmem + 2);         // [128]: r, z, hidden-n, input-n (layer 1) biases
    const uint32_t bar0 = smem_u32(s_bar);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t tfull_bar = bar0 + 8u * (2 * STAGES), tempty_bar = bar0 + 8u * (2 * STAGES + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int NT = p.H / 32;
    const int per_layer = p.mtiles * NT;
    const int layer = (int)blockIdx.x >= per_layer ? 1 : 0;
    const int rem = (int)blockIdx.x - layer * per_layer;
    const int mt = rem / NT, nt = rem % NT;
    const int m0 = mt * BM, n0 = nt * BN, j0 = nt * 32;
    const int* cnt0 = p.counters + mt;
    const int* cnt1 = p.counters + p.cstride + mt;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, kEpiWarps * 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 128) {
        const int g = tid >> 5, u = tid & 31;  // g: 0 r, 1 z, 2 hidden n, 3 input n
        float v;
        if (layer == 0) v = g < 3 ? __ldg(p.bhh0 + n0 + 32 * g + u) : 0.f;
        else if (g < 2) v = __ldg(p.bih1 + n0 + 32 * g + u) + __ldg(p.bhh1 + n0 + 32 * g + u);
        else v = g == 2 ? __ldg(p.bhh1 + n0 + 64 + u) : __ldg(p.bih1 + n0 + 64 + u);
        s_bias[tid] = v;
    }
    {   // resident weight slices: rows n0 .. n0+95, all of K; chunk j of row r at (r/8)*1024 + (r%8)*128 + ((j ^ (r%8)) << 4)
        const int nslices = layer ? 2 : 1;
        for (int u = tid; u < nslices * nkb * BN * 8; u += kThreads) {
            const int j = u & 7, r = (u >> 3) % BN, kbs = u / (8 * BN);  // kbs: k-block over [first slice | second slice]
            const int sl = kbs / nkb, kb = kbs % nkb;
            const __half* W = layer == 0 ? p.Whh0 : (sl == 0 ? p.Wih1 : p.Whh1);
            const uint32_t dst = w_smem + (uint32_t)kbs * W_KB + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
            cp_async16(dst, W + (long long)(n0 + r) * p.Kp + kb * KB_HALVES + 8 * j);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(s_tmem);

    if (warp_u == 0) {
        // ============================ producer: acquire the published states, TMA the rows ============================
        uint32_t ps = 0, pphase = 0;
        long long w_cnt = 0, w_empty = 0;
        auto load_rows = [&](const TmaDesc* map, int slot) {
            for (int kb = 0; kb < nkb; ++kb) {
                const long long t0 = kProf ? clock64() : 0;
                mbar_wait(empty_bar(ps), pphase ^ 1u);
                if (kProf) w_empty += clock64() - t0;
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar(ps), (uint32_t)A_STAGE);
                    tma_load_3d(a_smem + ps * (uint32_t)A_STAGE, map, full_bar(ps), kb * KB_HALVES, slot, p.b0 + m0);
                }
                __syncwarp();
                if (++ps == STAGES) {
                    ps = 0;
                    pphase ^= 1u;
                }
            }
        };
        for (int t = 0; t < p.T; ++t) {
            const long long t0 = kProf ? clock64() : 0;
            if (layer == 0) {
                if (t > 0) wait_counter(cnt0, NT * t, lane);  // every CTA of this m-tile has published its 32 units of h0_{t-1}
                if (kProf) w_cnt += clock64() - t0;
                load_rows(&p.h0map, t);
            } else {
                wait_counter(cnt0, NT * (t + 1), lane);  // h0_t
                if (kProf) w_cnt += clock64() - t0;
                load_rows(&p.h0map, t + 1);
                const long long t1 = kProf ? clock64() : 0;
                if (t > 0) wait_counter(cnt1, NT * t, lane);  // h1_{t-1}
                if (kProf) w_cnt += clock64() - t1;
                load_rows(&p.h1map, t);
            }
        }
        if (kProf && lane == 0) {
            atomicAdd(&g_gru_prof[layer][0], (unsigned long long)w_cnt);
            atomicAdd(&g_gru_prof[layer][1], (unsigned long long)w_empty);
        }
    } else if (warp_u == 1) {
        // ============================ MMA issuer (warp-uniform, one elected lane issues) ============================
        constexpr uint32_t idesc96 = (1u << 4) | ((uint32_t)(96 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);  // f16 x f16 -> f32
        constexpr uint32_t idesc64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        constexpr uint32_t idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint32_t tacc = __shfl_sync(0xffffffffu, tmem_base, 0);
        uint32_t ms = 0, mphase = 0;
        long long w_full = 0, w_tempty = 0;
        const long long t_begin = kProf ? clock64() : 0;
        for (int t = 0; t < p.T; ++t) {
            const long long t0 = kProf ? clock64() : 0;
            mbar_wait(tempty_bar, (uint32_t)((t & 1) ^ 1));  // the epilogue of step t-1 has drained the accumulator
            if (kProf) w_tempty += clock64() - t0;
            tc_fence_after();
            for (int kb = 0; kb < nkb; ++kb) {  // layer 0: h0_{t-1} x W_hh0;  layer 1: h0_t x W_ih1  -> columns 0..95
                const long long t1 = kProf ? clock64() : 0;
                mbar_wait(full_bar(ms), mphase);
                if (kProf) w_full += clock64() - t1;
                tc_fence_after();
                const uint64_t adesc = make_desc(a_smem + ms * (uint32_t)A_STAGE);
                const uint64_t bdesc = make_desc(w_smem + (uint32_t)kb * W_KB);
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tc_mma_f16(tacc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc96, (kb | kk) ? 1u : 0u);
                    tc_commit(empty_bar(ms));
                }
                __syncwarp();
                if (++ms == STAGES) {
                    ms = 0;
                    mphase ^= 1u;
                }
            }
            if (layer) {
                for (int kb = 0; kb < nkb; ++kb) {  // h1_{t-1} x W_hh1: r, z accumulate on columns 0..63, n_h -> 96..127
                    const long long t1 = kProf ? clock64() : 0;
                    mbar_wait(full_bar(ms), mphase);
                    if (kProf) w_full += clock64() - t1;
                    tc_fence_after();
                    const uint64_t adesc = make_desc(a_smem + ms * (uint32_t)A_STAGE);
                    const uint64_t bdesc = make_desc(w_smem + (uint32_t)(nkb + kb) * W_KB);
                    const uint64_t bdesc_n = bdesc + (uint64_t)((64 / 8) * 1024 >> 4);  // rows 64..95 of the slice
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            tc_mma_f16(tacc, adesc + (uint64_t)(2 * kk), bdesc + (uint64_t)(2 * kk), idesc64, 1u);
                            tc_mma_f16(tacc + 96u, adesc + (uint64_t)(2 * kk), bdesc_n + (uint64_t)(2 * kk), idesc32,
                                       (kb | kk) ? 1u : 0u);
                        }
                        tc_commit(empty_bar(ms));
                    }
                    __syncwarp();
                    if (++ms == STAGES) {
                        ms = 0;
                        mphase ^= 1u;
                    }
                }
            }
            if (elect_one()) tc_commit(tfull_bar);
            __syncwarp();
        }
        if (kProf && lane == 0) {
            atomicAdd(&g_gru_prof[layer][2], (unsigned long long)w_full);
            atomicAdd(&g_gru_prof[layer][3], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&g_gru_prof[layer][6], (unsigned long long)w_tempty);
            atomicAdd(&g_gru_prof[layer][7], (unsigned long long)p.T);
        }
    } else {
        // ============================ epilogue: thread = stream row x 16 hidden units ============================
        const int ew = warp - kFirstEpi;
        const int q = warp & 3;           // TMEM lane quarter this warp may access
        const int half16 = ew >> 2;       // which 16 of the CTA's 32 hidden units
        const int u0 = 16 * half16;
        const int m = m0 + q * 32 + lane;
        const bool ok = m < p.B;
        const long long b = ok ? m : 0;
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)u0;
        float* h32 = (layer ? p.h32_1 : p.h32_0) + b * p.H + j0 + u0;
        __half* hseq = (layer ? p.hseq1 : p.hseq0) + b * p.hB + j0 + u0;
        int* my_counter = p.counters + layer * p.cstride + mt;
        float ph[16];  // fp32 master state of this row's units: lives in registers for the whole chunk
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
            const float4 x = ok ? *(reinterpret_cast<const float4*>(h32) + v4) : make_float4(0.f, 0.f, 0.f, 0.f);
            ph[4 * v4] = x.x;
            ph[4 * v4 + 1] = x.y;
            ph[4 * v4 + 2] = x.z;
            ph[4 * v4 + 3] = x.w;
        }
        long long w_tfull = 0;
        const long long t_begin = kProf ? clock64() : 0;
        for (int t = 0; t < p.T; ++t) {
            // layer 0: the input projections (incl. b_ih) do not depend on the MMAs -- fetch them while those run
            float gr[16], gz[16], gn[16];
            if (layer == 0) {
                const float* gi = p.gi0 + b * p.giB + (long long)t * 3 * p.H + j0 + u0;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 a = ok ? __ldg(reinterpret_cast<const float4*>(gi) + v4) : z4;
                    const float4 c = ok ? __ldg(reinterpret_cast<const float4*>(gi + p.H) + v4) : z4;
                    const float4 d = ok ? __ldg(reinterpret_cast<const float4*>(gi + 2 * p.H) + v4) : z4;
                    gr[4 * v4] = a.x, gr[4 * v4 + 1] = a.y, gr[4 * v4 + 2] = a.z, gr[4 * v4 + 3] = a.w;
                    gz[4 * v4] = c.x, gz[4 * v4 + 1] = c.y, gz[4 * v4 + 2] = c.z, gz[4 * v4 + 3] = c.w;
                    gn[4 * v4] = d.x, gn[4 * v4 + 1] = d.y, gn[4 * v4 + 2] = d.z, gn[4 * v4 + 3] = d.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) gr[i] = gz[i] = gn[i] = 0.f;
            }
            {
                const long long t0 = kProf ? clock64() : 0;
                if (lane == 0) mbar_wait(tfull_bar, (uint32_t)(t & 1));
                __syncwarp();
                if (kProf && (ew & 3) == 0) w_tfull += clock64() - t0;
            }
            tc_fence_after();
            float hn[16];
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
                uint32_t v[32];
                tmem_ld8_nowait(tlane + 8 * c8, v);             // r
                tmem_ld8_nowait(tlane + 32 + 8 * c8, v + 8);    // z
                tmem_ld8_nowait(tlane + 64 + 8 * c8, v + 16);   // layer 0: hidden n;  layer 1: input n
                tmem_ld8_nowait(tlane + 96 + 8 * c8, v + 24);   // layer 1: hidden n
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int u = 8 * c8 + i;
                    const float rg = fast_sigmoid(gr[u] + __uint_as_float(v[i]) + s_bias[u0 + u]);
                    const float zg = fast_sigmoid(gz[u] + __uint_as_float(v[8 + i]) + s_bias[32 + u0 + u]);
                    const float vh = layer ? __uint_as_float(v[24 + i]) : __uint_as_float(v[16 + i]);
                    const float vi = layer ? __uint_as_float(v[16 + i]) : 0.f;
                    const float ng = fast_tanh(gn[u] + vi + s_bias[96 + u0 + u] + rg * (vh + s_bias[64 + u0 + u]));
                    hn[u] = (1.0f - zg) * ng + zg * ph[u];
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar);  // accumulator drained: the MMAs of step t+1 may overwrite it
#pragma unroll
            for (int u = 0; u < 16; ++u) ph[u] = hn[u];
            if (ok) {
                __half* hout = hseq + (long long)(t + 1) * p.H;
#pragma unroll
                for (int c8 = 0; c8 < 2; ++c8) {
                    const __half2 h0 = __floats2half2_rn(hn[8 * c8], hn[8 * c8 + 1]), h1 = __floats2half2_rn(hn[8 * c8 + 2], hn[8 * c8 + 3]),
                                  h2 = __floats2half2_rn(hn[8 * c8 + 4], hn[8 * c8 + 5]), h3 = __floats2half2_rn(hn[8 * c8 + 6], hn[8 * c8 + 7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<const uint32_t*>(&h0);
                    pk.y = *reinterpret_cast<const uint32_t*>(&h1);
                    pk.z = *reinterpret_cast<const uint32_t*>(&h2);
                    pk.w = *reinterpret_cast<const uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(hout + 8 * c8) = pk;
                }
            }
            // publish h_t: the named barrier orders the 256 threads' stores before the single release at gpu scope
            // (cumulative), which the producers of the whole m-tile group (and of layer 1) acquire
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
            if (ew == 0 && lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(my_counter) : "memory");
        }
        if (ok) {
#pragma unroll
            for (int v4 = 0; v4 < 4; ++v4)
                reinterpret_cast<float4*>(h32)[v4] = make_float4(ph[4 * v4], ph[4 * v4 + 1], ph[4 * v4 + 2], ph[4 * v4 + 3]);
        }
        if (kProf && (ew & 3) == 0 && lane == 0 && ew == 0) {
            atomicAdd(&g_gru_prof[layer][4], (unsigned long long)w_tfull);
            atomicAdd(&g_gru_prof[layer][5], (unsigned long long)(clock64() - t_begin));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

size_t gru_wave_smem_bytes(int H, int layers) {
    return 1024 + (size_t)(layers * (H / KB_HALVES)) * W_KB + STAGES * A_STAGE + 8 * (2 * STAGES + 2) + 16 + 128 * 4;
}

}  // namespace

bool gru_wave_supported(int H, int layers) {
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return false;
    return H % 64 == 0 && H >= 64 && gru_wave_smem_bytes(H, layers) <= 227 * 1024 && layers * (H / 32) <= num_sms;
}

int make_gru_wave_maps(GruWaveParams* p, int maxB) {
    // state histories [maxB][T+1][H] fp16: box = 64 units x 1 slot x 128 streams (rows past maxB are zero-filled)
    if (make_tma_3d_f16(&p->h0map, p->hseq0, p->H, p->T + 1, maxB, p->H, p->hB, KB_HALVES, 1, BM)) return 1;
    if (p->layers < 2) return 0;
    return make_tma_3d_f16(&p->h1map, p->hseq1, p->H, p->T + 1, maxB, p->H, p->hB, KB_HALVES, 1, BM);
}

int launch_gru_wave(const GruWaveParams& p, cudaStream_t st) {
    SE_REQUIRE((p.layers == 1 || p.layers == 2) && gru_wave_supported(p.H, p.layers),
               "gru_wave: hidden size must be a multiple of 64 whose weight slices fit shared memory");
    if (p.B <= 0) return 0;
    const size_t smem = gru_wave_smem_bytes(p.H, p.layers);
    SE_DYN_SMEM(gru_wave_kernel, smem);
    // The CTAs of an m-tile wait for one another (and layer 1 for layer 0) through counters in L2, so all of a launch
    // must be resident at once: one CTA per SM, layers x H/32 per m-tile.  More streams run as successive launches.
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    const int ntn = p.H / 32;
    const int tiles_per_wave = num_sms / (p.layers * ntn);
    const int mtiles = (p.B + BM - 1) / BM;
    SE_REQUIRE(mtiles <= p.cstride, "gru_wave: counter array too small");
    SE_CUDA_OK(cudaMemsetAsync(p.counters, 0, sizeof(int) * p.layers * p.cstride, st));
    for (int mt0 = 0; mt0 < mtiles; mt0 += tiles_per_wave) {
        GruWaveParams w = p;
        const int b0 = mt0 * BM;
        const int nt = mtiles - mt0 < tiles_per_wave ? mtiles - mt0 : tiles_per_wave;
        w.B = (p.B - b0) < nt * BM ? (p.B - b0) : nt * BM;
        w.b0 = p.b0 + b0;
        w.mtiles = nt;
        w.gi0 = p.gi0 + (long long)b0 * p.giB;
        w.hseq0 = p.hseq0 + (long long)b0 * p.hB;
        w.h32_0 = p.h32_0 + (long long)b0 * p.H;
        if (p.layers == 2) {
            w.hseq1 = p.hseq1 + (long long)b0 * p.hB;
            w.h32_1 = p.h32_1 + (long long)b0 * p.H;
        }
        w.counters = p.counters + mt0;
        gru_wave_kernel<<<p.layers * nt * ntn, kThreads, smem, st>>>(w);
        SE_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int gru_profile_read(unsigned long long* out16, int reset) {
    SE_CUDA_OK(cudaDeviceSynchronize());
    if (out16) SE_CUDA_OK(cudaMemcpyFromSymbol(out16, g_gru_prof, 16 * sizeof(unsigned long long)));
    if (reset) {
        const unsigned long long z[16] = {0};
        SE_CUDA_OK(cudaMemcpyToSymbol(g_gru_prof, z, sizeof(z)));
    }
    return 0;
}

}  // namespace se
