// Persistent GRU recurrence for small batches (training micro-step and whole-file forward, CRN_ELU.py:173 through
// nn.GRU): one launch walks the time steps of a range of chunks of a layer with W_hh resident in registers.  Two forms:
//   * cluster-resident (default; round 2): a thread-block cluster of H / 32 CTAs per sequence, the state exchanged
//     through distributed shared memory with self-completing stores / bulk copies (gru_seq_*_cluster_kernel below);
//   * cooperative (round 1; SE_B200_GRU_CLUSTER=0 or where a cluster cannot be placed): each warp owns one hidden unit
//     and keeps its three rows of W_hh (forward) or its column of W_hh (backward) in registers, the hidden state goes
//     through L2 (a few KB) and the steps are separated by grid-wide barriers.
// The per-step GEMM launches these replace spent 44 us per step streaming 3 MB of weights out of L2 for a single row of
// output (profiles/r01_train_launches_*.csv).
//
// Batch layout ("chunk-major"): stream s = n * nb + i is chunk n of utterance i; the state entering chunk n is the
// state leaving chunk n-1 (CRN_ELU.py:173,183-185), detached for the backward (the gradient stops at chunk borders).
#include <cooperative_groups.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

#include "se_internal.h"

namespace cg = cooperative_groups;

namespace se {
namespace {

constexpr int kWarps = 4;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// KPL = H / 32 weights per lane and gate
template <int KPL>
__global__ void __launch_bounds__(kWarps * 32) gru_seq_fwd_kernel(GruSeqParams p) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    // gridDim.x = unit blocks x stream groups: group sg serves streams i = sg, sg + SG, ... of every chunk, so that larger
    // batches are spread over more warps instead of lengthening the latency-bound pass of one warp
    const int ublocks = (p.H + kWarps - 1) / kWarps;
    const int SG = gridDim.x / ublocks, sg = blockIdx.x / ublocks;
    const int j = (blockIdx.x % ublocks) * kWarps + (threadIdx.x >> 5);  // hidden unit of this warp
    const int H = p.H, T = p.T;
    const bool active = j < H;
    float wr[KPL], wz[KPL], wn[KPL];
    float br = 0.f, bz = 0.f, bn = 0.f;
    if (active) {
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
            const int k = lane + 32 * q;
            wr[q] = p.Whh[(long long)j * p.Kp + k];
            wz[q] = p.Whh[(long long)(H + j) * p.Kp + k];
            wn[q] = p.Whh[(long long)(2 * H + j) * p.Kp + k];
        }
        br = p.bhh[j];
        bz = p.bhh[H + j];
        bn = p.bhh[2 * H + j];
    }
    for (int n = p.n0; n < p.n0 + p.N; ++n) {
        for (int t = 0; t < T; ++t) {
            if (active) {
                for (int i = sg; i < p.nb; i += SG) {
                    const long long s = (long long)n * p.nb + i;
                    const float* hp = (t == 0 && n > 0) ? p.hseq + (s - p.nb) * p.hB + (long long)T * H
                                                        : p.hseq + s * p.hB + (long long)t * H;
                    const float* g = p.gi + s * p.giB + (long long)t * 3 * H;
                    float gr = 0.f, gz = 0.f, gn = 0.f, hj = 0.f;
                    if (lane == 0) {
                        gr = __ldg(g + j);
                        gz = __ldg(g + H + j);
                        gn = __ldg(g + 2 * H + j);
                        hj = __ldcg(hp + j);
                    }
                    float ar = 0.f, az = 0.f, an = 0.f;
#pragma unroll
                    for (int q = 0; q < KPL; ++q) {
                        const float h = __ldcg(hp + lane + 32 * q);
                        ar = fmaf(wr[q], h, ar);
                        az = fmaf(wz[q], h, az);
                        an = fmaf(wn[q], h, an);
                    }
                    ar = warp_sum(ar);
                    az = warp_sum(az);
                    an = warp_sum(an);
                    if (lane == 0) {
                        const float r = sigmoidf_(gr + ar + br);
                        const float z = sigmoidf_(gz + az + bz);
                        const float c = tanhf(gn + r * (an + bn));
                        p.hseq[s * p.hB + (long long)(t + 1) * H + j] = (1.0f - z) * c + z * hj;
                        if (t == 0 && n > 0) p.hseq[s * p.hB + j] = hj;  // slot 0 = state entering the chunk (backward reads it)
                    }
                }
            }
            grid.sync();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster-resident forward.  The chain above costs a grid-wide barrier plus an L2 round trip of h per step (~3 us x
// N*T steps: 2.5 of the 3 ms of a training forward).  Here one thread-block CLUSTER of H / 32 CTAs owns a sequence:
//   * 16 warps per CTA, a warp owns two hidden units: its six rows of W_hh stay in registers, h_{t-1} is read from the
//     CTA's own full copy in shared memory (double-buffered), the six dot products are summed over the warp by a
//     9-shuffle fold and left in shared memory;
//   * after ONE __syncthreads a single warp per utterance finishes the cell for the CTA's 32 units (lane = unit) and
//     sends the 32 new values into every CTA's copy of h -- the own one included -- through distributed shared memory
//     with 16-byte st.async stores that carry their own completion (complete_tx on the receiver's mbarrier);
//   * a CTA starts step t + 1 when its mbarrier has counted all H values of step t.  No grid barrier, no cluster barrier,
//     no fence.  A peer can only overwrite the buffer a warp is still reading after this CTA's values of the same step
//     have arrived there, and the cell warp sends them after every warp of the CTA has passed the __syncthreads behind
//     its reads.
// How it got there (profiles/r02_gru_seq_cluster_stalls.txt): barrier.cluster.arrive.release made every warp wait for its
// hseq store to reach L2 (2.05 us per step, no better than the grid barrier); 4-byte st.async's cost the receiver 512
// mbarrier updates per step; the cell in two lanes of every warp made the step issue-bound; a long single-warp cell
// (expf / tanhf / divisions, 16 bulk copies issued one by one) left the other 15 warps waiting.  The input projections of
// the next step arrive by cp.async behind the sends.  Utterances are spread over clusters (no dependence between
// them); a cluster serves its utterances inside the same step, one cell warp each.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kClWarps = 16;  // x 2 units = 32 units per CTA

// one lane of a converged warp; with the operands warp-uniform, ptxas issues the guarded instruction once from the
// uniform datapath instead of wrapping it into a per-lane BRA.U.ANY loop (as it does behind `if (lane == 0)`)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// the plain form, as for any TMA-filled buffer: data and completion arrive through the async proxy; the
// .acquire.cluster form only adds an L1 invalidation (CCTL.IVALL) per thread and step
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    }
}

// Sums of eight per-lane accumulators over the warp in 9 shuffles (instead of 8 x 5): each butterfly level halves the
// number of live values, the kept half chosen by the lane's bit.  Returns in EVERY lane the total of accumulator lane / 4.
__device__ __forceinline__ float fold8(float (&a)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float b[4], c[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? a[i] : a[i + 4];
        b[i] = (b4 ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? b[i] : b[i + 2];
        c[i] = (b3 ? b[i + 2] : b[i]) + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    float d = (b2 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b2 ? c[0] : c[1], 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}

// packed fp32 pairs for fma.rn.f32x2 (sm_100: two fp32 FMAs per issue slot)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_f(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi_f(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ void ffma2(unsigned long long& acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// MUFU forms (ex2 / rcp, ~2 ulp): a dozen instructions for the three activations of a cell instead of ~150 for expf,
// tanhf and two IEEE divisions -- the cell is the serial stretch of every step.  |error| < 3e-7 absolute.
__device__ __forceinline__ float fast_sigmoid(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return r;
}
__device__ __forceinline__ float fast_tanh(float x) { return 2.0f * fast_sigmoid(2.0f * x) - 1.0f; }

template <int KPL>  // H / 32 = CTAs per cluster
__global__ void __launch_bounds__(kClWarps * 32, 1) gru_seq_fwd_cluster_kernel(GruSeqParams p, int SG, int upc) {
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long mbar[2];
    constexpr int H = 32 * KPL;
    float* hb = sm;                          // [2][upc][H]   h of the step, every CTA holds all of it
    float* gst = hb + 2 * upc * H;           // [2][upc][3][32] input projections (r, z, n) of the CTA's 32 units
    float* dots = gst + 2 * upc * 96;        // [upc][3][32]  W_hh . h of the CTA's 32 units
    float* outst = dots + upc * 96;          // [2][upc][32]  the CTA's new values, re-read as 16-byte chunks by the sending lanes
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = blockIdx.x % KPL, sg = blockIdx.x / KPL;
    const int j0 = rank * 32 + warp * 2;  // first hidden unit of this warp's dot products
    const int T = p.T;
    int mine = 0;  // utterances of this cluster: sg, sg + SG, ...
    for (int u = 0; u < upc; ++u) mine += (sg + u * SG < p.nb) ? 1 : 0;
    const uint32_t tx_bytes = (uint32_t)mine * H * sizeof(float);
    // rows r(j0) r(j0+1) z(j0) z(j0+1) n(j0) n(j0+1) of W_hh.  H >= 128: the lane owns k = 128 q4 + 4 lane + {0..3}, held
    // as (k, k + 1) pairs, so that h arrives by 16-byte loads and a packed fma.rn.f32x2 does two k of a row at once (the
    // dot products were issue-bound: 16 LDS + 96 FFMA per warp and step become 4 LDS.128 + 48 FFMA2).  Smaller H: lane
    // owns k = lane + 32 q, scalar.
    constexpr bool PACK = KPL >= 4;
    constexpr int NW = PACK ? KPL / 2 : KPL;
    unsigned long long w2[6][PACK ? NW : 1];
    float w[6][PACK ? 1 : KPL];
#pragma unroll
    for (int gte = 0; gte < 3; ++gte)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float* row = p.Whh + (long long)(gte * H + j0 + u) * p.Kp;
            if (PACK) {
#pragma unroll
                for (int q = 0; q < NW; ++q) {
                    const int k = 128 * (q >> 1) + 4 * lane + 2 * (q & 1);
                    w2[2 * gte + u][q] = pack2(row[k], row[k + 1]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < KPL; ++q) w[2 * gte + u][q] = row[lane + 32 * q];
            }
        }
    // warp u finishes the cell of utterance u for all 32 units of the CTA (lane = unit): the activations run once per CTA
    // with full warps instead of in two lanes of every warp (which made the step issue-bound: 481 instructions per warp
    // and step)
    const int jc = rank * 32 + lane;
    float bh = 0.f, bz = 0.f, bn = 0.f;
    if (warp < mine) {
        bh = p.bhh[jc];
        bz = p.bhh[H + jc];
        bn = p.bhh[2 * H + jc];
    }
    // state entering chunk 0 (slot 0 of hseq, written by the caller)
    for (int u = 0; u < mine; ++u) {
        const int i = sg + u * SG;
        const float* h0 = p.n0 == 0 ? p.hseq + (long long)i * p.hB  // slot 0 of chunk 0, or the last state of chunk n0 - 1
                                    : p.hseq + ((long long)(p.n0 - 1) * p.nb + i) * p.hB + (long long)T * H;
        for (int k = threadIdx.x; k < H; k += blockDim.x) hb[u * H + k] = h0[k];
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto stage_gi = [&](int buf, int step) {  // cell warps: the projections of utterance `warp` at `step`
        const int n = p.n0 + step / T, t = step % T;
        const float* g = p.gi + ((long long)n * p.nb + sg + warp * SG) * p.giB + (long long)t * 3 * H + jc;
#pragma unroll
        for (int gte = 0; gte < 3; ++gte) {
            const uint32_t dst = smem_addr(gst + ((buf * upc + warp) * 3 + gte) * 32 + lane);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g + gte * H) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (warp < mine) stage_gi(0, 0);
    __syncthreads();
    cluster_arrive();  // every CTA of the cluster runs and has its mbarriers before anyone stores into it
    cluster_wait();
    const int steps = p.N * T;
    for (int step = 0; step < steps; ++step) {
        const int cur = step & 1;
        if (step > 0) mbar_wait(smem_addr(&mbar[cur]), ((step - 1) >> 1) & 1);
        for (int u = 0; u < mine; ++u) {
            const float* h = hb + (cur * upc + u) * H;
            float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (PACK) {
                unsigned long long acc[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};  // (even k, odd k) partial sums
#pragma unroll
                for (int q4 = 0; q4 < KPL / 4; ++q4) {
                    const ulonglong2 hv = *reinterpret_cast<const ulonglong2*>(h + 128 * q4 + 4 * lane);
#pragma unroll
                    for (int r = 0; r < 6; ++r) {
                        ffma2(acc[r], w2[r][2 * q4], hv.x);
                        ffma2(acc[r], w2[r][2 * q4 + 1], hv.y);
                    }
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) a[r] = lo_f(acc[r]) + hi_f(acc[r]);
            } else {
#pragma unroll
                for (int q = 0; q < KPL; ++q) {
                    const float hv = h[lane + 32 * q];
#pragma unroll
                    for (int r = 0; r < 6; ++r) a[r] = fmaf(w[r][q], hv, a[r]);
                }
            }
            const float tot = fold8(a, lane);  // accumulator lane / 4 = 2 * gate + unit
            if ((lane & 3) == 0 && lane < 24) dots[(u * 3 + (lane >> 3)) * 32 + warp * 2 + ((lane >> 2) & 1)] = tot;
        }
        __syncthreads();
        if (warp >= mine) continue;
        // ---- cell of utterance `warp` (lane = unit).  This single-warp stretch is the critical path of the step (every
        // other warp of the cluster waits for its result), so it is kept short: MUFU activations, 16-byte st.async sends
        // straight from the lanes, and everything that can wait (next step's projections, the hseq stores) after them
        const int u = warp;
        if (warp == 0 && lane == 0 && step + 1 < steps)  // the buffer filled during this step
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar[cur ^ 1])), "r"(tx_bytes)
                         : "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");  // the projections of this step (staged one step earlier)
        const float* gs = gst + (cur * upc + u) * 96;
        const float* d = dots + u * 96;
        const float hj = hb[(cur * upc + u) * H + jc];
        const float r = fast_sigmoid(gs[lane] + d[lane] + bh);
        const float z = fast_sigmoid(gs[32 + lane] + d[32 + lane] + bz);
        const float c = fast_tanh(gs[64 + lane] + r * (d[64 + lane] + bn));
        const float hnew = (1.0f - z) * c + z * hj;
        if (step + 1 < steps) {
            // lane L sends the 16-byte chunk L % 8 of the CTA's 32 new values to the CTAs L / 8 + 4 k (the own one
            // included); every store carries its completion to the receiver's mbarrier
            float* mine_out = outst + (cur * upc + u) * 32;
            mine_out[lane] = hnew;
            __syncwarp();
            const float4 chunk = *reinterpret_cast<const float4*>(mine_out + 4 * (lane & 7));
            const uint32_t dstl = smem_addr(hb + ((cur ^ 1) * upc + u) * H + rank * 32 + 4 * (lane & 7));
            const uint32_t lbar = smem_addr(&mbar[cur ^ 1]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int peer = (lane >> 3) + 4 * k;
                if (peer < KPL) {
                    uint32_t remote, rbar;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(dstl), "r"(peer));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(lbar), "r"(peer));
                    asm volatile(
                        "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                            remote),
                        "r"(__float_as_uint(chunk.x)), "r"(__float_as_uint(chunk.y)), "r"(__float_as_uint(chunk.z)),
                        "r"(__float_as_uint(chunk.w)), "r"(rbar)
                        : "memory");
                }
            }
            stage_gi(cur ^ 1, step + 1);
        }
        const int n = p.n0 + step / T, t = step % T;
        const long long s = (long long)n * p.nb + sg + u * SG;
        p.hseq[s * p.hB + (long long)(t + 1) * H + jc] = hnew;
        if (t == 0 && n > 0) p.hseq[s * p.hB + jc] = hj;  // slot 0 = state entering the chunk (backward reads it)
    }
    cluster_arrive();  // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

// Co-resident clusters of `kernel` (cluster size KPL, `smem` dynamic bytes) on the current device; 0 = cannot be placed.
// The occupancy query and the non-portable-size opt-in cost tens of milliseconds of host time: once per (device, kernel).
template <typename K>
int cluster_capacity(K kernel, int KPL, size_t smem, int* out) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cache;
    int dev = 0;
    SE_CUDA_OK(cudaGetDevice(&dev));
    const auto key = std::make_pair(reinterpret_cast<const void*>(kernel), dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return 0;
    }
    int ncl = 0;
    bool ok = true;
    if (KPL > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) ok = false;
    if (ok && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) ok = false;
    if (ok) {
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = KPL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cfg.blockDim = dim3(kClWarps * 32);
        cfg.gridDim = dim3(KPL);
        cfg.dynamicSmemBytes = smem;
        if (cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg) != cudaSuccess) ncl = 0;
    }
    cudaGetLastError();
    cache[key] = ncl;
    *out = ncl;
    return 0;
}

template <int KPL>
int launch_cluster(const GruSeqParams& p, cudaStream_t st, bool* launched) {
    *launched = false;
    auto kernel = gru_seq_fwd_cluster_kernel<KPL>;
    auto smem_of = [&](int upc) { return (size_t)(2 * upc * 32 * KPL + 2 * upc * 96 + upc * 96 + 2 * upc * 32) * sizeof(float); };
    int ncl = 0;
    if (cluster_capacity(kernel, KPL, smem_of(1), &ncl)) return 1;
    if (ncl < 1) return 0;  // this device cannot place the cluster: the caller takes the cooperative kernel
    int SG = p.nb < ncl ? p.nb : ncl;
    int upc = (p.nb + SG - 1) / SG;
    SG = (p.nb + upc - 1) / upc;
    if (smem_of(upc) > 200 * 1024 || upc > kClWarps) return 0;  // one cell warp per utterance of the cluster
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = KPL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(kClWarps * 32);
    cfg.stream = st;
    cfg.gridDim = dim3(KPL * SG);
    cfg.dynamicSmemBytes = smem_of(upc);
    SE_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p, SG, upc));
    *launched = true;
    return 0;
}

// backward through the T steps of every stream at once (no dependence between chunks: the carried state is detached)
template <int KPL>
__global__ void __launch_bounds__(kWarps * 32) gru_seq_bwd_kernel(GruSeqBwdParams p) {
    cg::grid_group grid = cg::this_grid();
    const int lane = threadIdx.x & 31;
    const int ublocks = (p.H + kWarps - 1) / kWarps;
    const int SG = gridDim.x / ublocks, sg = blockIdx.x / ublocks;  // stream groups, as in the forward kernel
    const int k = (blockIdx.x % ublocks) * kWarps + (threadIdx.x >> 5);  // hidden unit (input side of W_hh) of this warp
    const int H = p.H, T = p.T;
    const bool active = k < H;
    float wc[3 * KPL];  // column k of W_hh: rows n = lane + 32 q, q < 3H/32
    if (active) {
#pragma unroll
        for (int q = 0; q < 3 * KPL; ++q) wc[q] = p.Whh[(long long)(lane + 32 * q) * p.Kp + k];
    }
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gthreads = (long long)gridDim.x * blockDim.x;
    const long long BH = (long long)p.B * H;
    float* cur = p.dhrec;        // d loss / d h_t arriving through the recurrence ([B][H]); zero at the last step
    float* nxt = p.dhrec + BH;
    for (long long e = gtid; e < BH; e += gthreads) cur[e] = 0.f;
    grid.sync();
    for (int t = T - 1; t >= 0; --t) {
        // phase A: cell backward, elementwise over (stream, unit)
        for (long long e = gtid; e < BH; e += gthreads) {
            const int j = (int)(e % H);
            const long long s = e / H;
            const float* a = p.gi + s * p.gB + (long long)t * 3 * H;
            const float* h = p.gh + s * p.gB + (long long)t * 3 * H;
            const float r = sigmoidf_(a[j] + h[j]);
            const float z = sigmoidf_(a[H + j] + h[H + j]);
            const float hn = h[2 * H + j];
            const float c = tanhf(a[2 * H + j] + r * hn);
            const float hp = p.hseq[s * p.hB + (long long)t * H + j];
            const float dh = p.dH[s * p.hB + (long long)(t + 1) * H + j] + __ldcg(cur + e);
            const float dan = dh * (1.f - z) * (1.f - c * c);
            const float daz = dh * (hp - c) * z * (1.f - z);
            const float dar = dan * hn * r * (1.f - r);
            float* o = p.dgi + s * p.gB + (long long)t * 3 * H;
            float* q = p.dgh + s * p.gB + (long long)t * 3 * H;
            o[j] = dar;
            o[H + j] = daz;
            o[2 * H + j] = dan;
            q[j] = dar;
            q[H + j] = daz;
            q[2 * H + j] = dan * r;
            nxt[e] = dh * z;
        }
        if (t == 0) break;  // the state entering the chunk is detached (CRN_ELU.py:185)
        grid.sync();
        // phase B: nxt[s][k] += sum_n dgh[s][t][n] * W_hh[n][k]
        if (active) {
            for (int s = sg; s < p.B; s += SG) {
                const float* q = p.dgh + (long long)s * p.gB + (long long)t * 3 * H;
                float acc = 0.f;
#pragma unroll
                for (int u = 0; u < 3 * KPL; ++u) acc = fmaf(wc[u], __ldcg(q + lane + 32 * u), acc);
                acc = warp_sum(acc);
                if (lane == 0) nxt[(long long)s * H + k] = __ldcg(nxt + (long long)s * H + k) + acc;
            }
        }
        grid.sync();
        float* tmp = cur;
        cur = nxt;
        nxt = tmp;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster-resident backward.  The chunk-streams are independent sequences of T steps (the carried state is detached),
// so nothing needs the whole grid: a cluster of H / 32 CTAs holds W_hh by columns in registers (a warp owns two
// columns = two hidden units on the input side, same lane-strided order and butterfly as above: bit-identical) and walks
// UP sequences at a time from t = T - 1 down to 0.  Per step: the cell adjoint of the CTA's 32 units by one warp per
// sequence (lane = unit), the 3H values of d loss / d gh all-gathered into every CTA's shared memory by bulk copies
// with complete_tx (see the forward), then every warp's two dot products per sequence over them (one 9-shuffle fold
// for all of them), handed back to the cell warps through shared memory.  The cell inputs of the next step arrive by
// cp.async.
// The cooperative kernel above spent two barriers over ~1000 CTAs per step and walked a group's streams one after
// another: 0.44 ms per layer for 24 sequences, 2.3 ms for 192.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBwdUP = 4;  // sequences in flight per cluster

template <int KPL>
__global__ void __launch_bounds__(kClWarps * 32, 1) gru_seq_bwd_cluster_kernel(GruSeqBwdParams p, int NCL) {
    constexpr int H = 32 * KPL, H3 = 3 * H, UP = kBwdUP;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) unsigned long long mbar[2];
    float* gb = sm;                    // [2][UP][KPL][3][32] d loss / d gh of the step, every CTA holds all of it
    float* stg = gb + 2 * UP * H3;     // [2][UP][8][32] cell inputs a_r a_z a_n h_r h_z h_n h_prev dH of the CTA's 32 units
    float* accs = stg + 2 * UP * 256;  // [UP][32] W_hh^T . d gh of the CTA's 32 units
    float* outst = accs + UP * 32;     // [2][UP][3][32] the CTA's part of d gh, source of the bulk copies
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = blockIdx.x % KPL, cl = blockIdx.x / KPL;
    const int j0 = rank * 32 + warp * 2;  // columns of this warp's dot products
    const int jc = rank * 32 + lane;      // unit of this lane in the cell phase (warps 0 .. UP-1, warp = sequence)
    const int T = p.T;
    const int nseq = cl < p.B ? (p.B - cl + NCL - 1) / NCL : 0;  // sequences cl, cl + NCL, ...
    const int ngroups = (nseq + UP - 1) / UP;
    // columns j0, j0 + 1 of W_hh.  The gathered vector lies in shared memory as [rank][gate][32] blocks; KPL even: the lane
    // owns the element pair 2 (lane % 16), + 1 of the blocks 2 bb + lane / 16 (8-byte loads, packed fma.rn.f32x2 over the
    // pair: 24 LDS.64 + 48 FFMA2 per sequence instead of 48 LDS + 96 FFMA); KPL = 1: rows n = lane + 32 q, scalar.
    constexpr bool PACK = (KPL % 2) == 0;
    constexpr int NP = 3 * KPL / 2;
    unsigned long long wc2[2][PACK ? NP : 1];
    float wc[2][PACK ? 1 : 3 * KPL];
    const int half = lane >> 4, i0 = 2 * (lane & 15);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (PACK) {
#pragma unroll
            for (int bb = 0; bb < NP; ++bb) {
                const int b = 2 * bb + half;  // block = rank * 3 + gate
                const long long n = (long long)(b % 3) * H + (b / 3) * 32 + i0;
                wc2[c][bb] = pack2(p.Whh[n * p.Kp + j0 + c], p.Whh[(n + 1) * p.Kp + j0 + c]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 3 * KPL; ++q) wc[c][q] = p.Whh[(long long)(lane + 32 * q) * p.Kp + j0 + c];
        }
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // cell inputs of step `it` (group it / T, t = T - 1 - it % T) of sequence `warp` of the group into staging buffer it & 1
    auto stage = [&](int it) {
        const int grp = it / T, t = T - 1 - (it - grp * T);
        const int q = grp * UP + warp;
        if (q < nseq) {
            const long long s = cl + (long long)q * NCL;
            const float* a = p.gi + s * p.gB + (long long)t * H3 + jc;
            const float* h = p.gh + s * p.gB + (long long)t * H3 + jc;
            const uint32_t dst = smem_addr(stg + (((it & 1) * UP + warp) * 8) * 32 + lane);
#pragma unroll
            for (int gte = 0; gte < 3; ++gte) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + gte * 128), "l"(a + gte * H) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (3 + gte) * 128), "l"(h + gte * H) : "memory");
            }
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 6 * 128), "l"(p.hseq + s * p.hB + (long long)t * H + jc)
                         : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 7 * 128),
                         "l"(p.dH + s * p.hB + (long long)(t + 1) * H + jc)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int iters = ngroups * T;
    if (iters > 0 && warp < UP) stage(0);
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    float keep = 0.f;  // cell warps: dh * z of the lane's unit
    int x = 0;         // exchanges so far (steps with t > 0)
    for (int it = 0; it < iters; ++it) {
        const int grp = it / T, t = T - 1 - (it - grp * T);
        int act = nseq - grp * UP;
        if (act > UP) act = UP;
        if (warp < act) {  // cell adjoint of sequence `warp` of the group, lane = unit
            // d loss / d h_t arriving through the recurrence: dh_{t+1} z + W_hh^T d gh_{t+1} (the dot products of the
            // previous step, behind its __syncthreads)
            const float rec = t == T - 1 ? 0.f : keep + accs[warp * 32 + lane];
            if (warp == 0 && lane == 0 && t > 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar[x & 1])),
                             "r"((uint32_t)(act * H3 * sizeof(float)))
                             : "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");  // the cell inputs of this step (staged one step earlier)
            const long long s = cl + (long long)(grp * UP + warp) * NCL;
            const float* in = stg + (((it & 1) * UP + warp) * 8) * 32 + lane;
            const float r = fast_sigmoid(in[0] + in[3 * 32]);
            const float z = fast_sigmoid(in[32] + in[4 * 32]);
            const float hn = in[5 * 32];
            const float c = fast_tanh(in[2 * 32] + r * hn);
            const float hp = in[6 * 32];
            const float dh = in[7 * 32] + rec;
            const float dan = dh * (1.f - z) * (1.f - c * c);
            const float daz = dh * (hp - c) * z * (1.f - z);
            const float dar = dan * hn * r * (1.f - r);
            keep = dh * z;
            if (t > 0) {  // all-gather: one 384-byte bulk copy per CTA of the cluster (the own one included); before the
                          // global stores, which the fence would otherwise wait for
                float* own = outst + ((x & 1) * UP + warp) * 96 + lane;
                own[0] = dar;
                own[32] = daz;
                own[64] = dan * r;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                const uint32_t src = smem_addr(outst + ((x & 1) * UP + warp) * 96);
                const uint32_t dstl = smem_addr(gb + ((x & 1) * UP + warp) * H3 + rank * 96);
                const uint32_t lbar = smem_addr(&mbar[x & 1]);
#pragma unroll
                for (int peer = 0; peer < KPL; ++peer) {
                    uint32_t remote, rbar;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(dstl), "r"(peer));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(lbar), "r"(peer));
                    if (elect_one())
                        asm volatile(
                            "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 384, [%2];" ::"r"(
                                remote),
                            "r"(src), "r"(rbar)
                            : "memory");
                }
            }
            if (it + 1 < iters) stage(it + 1);  // after the sends: off the critical path of the step
            float* o = p.dgi + s * p.gB + (long long)t * H3;
            float* q = p.dgh + s * p.gB + (long long)t * H3;
            o[jc] = dar;
            o[H + jc] = daz;
            o[2 * H + jc] = dan;
            q[jc] = dar;
            q[H + jc] = daz;
            q[2 * H + jc] = dan * r;
        }
        if (t == 0) continue;  // the state entering the chunk is detached (CRN_ELU.py:185)
        mbar_wait(smem_addr(&mbar[x & 1]), (x >> 1) & 1);
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // (sequence u, column c) at 2 u + c
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            if (u >= act) continue;
            const float* g = gb + ((x & 1) * UP + u) * H3;
            if (PACK) {
                unsigned long long s0 = 0ull, s1 = 0ull;
#pragma unroll
                for (int bb = 0; bb < NP; ++bb) {
                    const unsigned long long gv = *reinterpret_cast<const unsigned long long*>(g + (2 * bb + half) * 32 + i0);
                    ffma2(s0, wc2[0][bb], gv);
                    ffma2(s1, wc2[1][bb], gv);
                }
                a[2 * u] = lo_f(s0) + hi_f(s0);
                a[2 * u + 1] = lo_f(s1) + hi_f(s1);
            } else {
#pragma unroll
                for (int q = 0; q < 3 * KPL; ++q) {
                    const float gv = g[((q % KPL) * 3 + q / KPL) * 32 + lane];  // row n = lane + 32 q = gate (q / KPL), rank (q % KPL)
                    a[2 * u] = fmaf(wc[0][q], gv, a[2 * u]);
                    a[2 * u + 1] = fmaf(wc[1][q], gv, a[2 * u + 1]);
                }
            }
        }
        const float tot = fold8(a, lane);
        if ((lane & 3) == 0 && (lane >> 3) < act) accs[(lane >> 3) * 32 + warp * 2 + ((lane >> 2) & 1)] = tot;
        __syncthreads();
        ++x;
    }
    cluster_arrive();  // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

template <int KPL>
int launch_cluster_bwd(const GruSeqBwdParams& p, cudaStream_t st, bool* launched) {
    *launched = false;
    auto kernel = gru_seq_bwd_cluster_kernel<KPL>;
    const size_t smem = (size_t)(2 * kBwdUP * 3 * 32 * KPL + 2 * kBwdUP * 256 + kBwdUP * 32 + 2 * kBwdUP * 96) * sizeof(float);
    int ncl = 0;
    if (cluster_capacity(kernel, KPL, smem, &ncl)) return 1;
    if (ncl < 1) return 0;
    const int NCL = p.B < ncl ? p.B : ncl;  // one wave of clusters, sequences cl, cl + NCL, ... each
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = KPL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(kClWarps * 32);
    cfg.stream = st;
    cfg.gridDim = dim3(KPL * NCL);
    cfg.dynamicSmemBytes = smem;
    SE_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p, NCL));
    *launched = true;
    return 0;
}

template <typename P, typename K>
int launch_coop(K kernel, const P& p, int H, int want_groups, cudaStream_t st, const char* what) {
    const int ublocks = (H + kWarps - 1) / kWarps;
    int dev = 0, sms = 0, per_sm = 0;
    SE_CUDA_OK(cudaGetDevice(&dev));
    SE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarps * 32, 0));
    SE_REQUIRE(ublocks <= sms * per_sm, std::string(what) + ": hidden size too large for one co-resident grid");
    int groups = want_groups < 1 ? 1 : want_groups;
    if (groups > sms * per_sm / ublocks) groups = sms * per_sm / ublocks;  // every CTA must be resident (grid barriers)
    const int grid = ublocks * groups;
    P copy = p;
    void* args[] = {&copy};
    SE_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(grid), dim3(kWarps * 32), args, 0, st));
    return 0;
}

}  // namespace

#define SE_GRU_DISPATCH(KERNEL, P)                                                        \
    switch (p.H / 32) {                                                                   \
        case 1: return launch_coop(KERNEL<1>, p, p.H, groups, st, #KERNEL);               \
        case 2: return launch_coop(KERNEL<2>, p, p.H, groups, st, #KERNEL);               \
        case 4: return launch_coop(KERNEL<4>, p, p.H, groups, st, #KERNEL);               \
        case 8: return launch_coop(KERNEL<8>, p, p.H, groups, st, #KERNEL);               \
        case 16: return launch_coop(KERNEL<16>, p, p.H, groups, st, #KERNEL);             \
    }

bool gru_seq_supported(int H) { return H == 32 || H == 64 || H == 128 || H == 256 || H == 512; }

int launch_gru_seq_fwd(const GruSeqParams& p, cudaStream_t st) {
    SE_REQUIRE(gru_seq_supported(p.H), "gru_seq: hidden size must be 32, 64, 128, 256 or 512");
    if (p.nb <= 0 || p.N <= 0) return 0;
    const char* e = getenv("SE_B200_GRU_CLUSTER");  // 0: cooperative kernel with grid barriers (round-1 form)
    const bool use_cluster = e == nullptr || atoi(e) != 0;
    if (use_cluster) {
        bool launched = false;
        int rc = 0;
        switch (p.H / 32) {
            case 1: rc = launch_cluster<1>(p, st, &launched); break;
            case 2: rc = launch_cluster<2>(p, st, &launched); break;
            case 4: rc = launch_cluster<4>(p, st, &launched); break;
            case 8: rc = launch_cluster<8>(p, st, &launched); break;
            case 16: rc = launch_cluster<16>(p, st, &launched); break;
        }
        if (rc) return rc;
        if (launched) return 0;
    }
    const int groups = p.nb < 4 ? p.nb : 4;
    SE_GRU_DISPATCH(gru_seq_fwd_kernel, GruSeqParams)
    return 2;
}

int launch_gru_seq_bwd(const GruSeqBwdParams& p, cudaStream_t st) {
    SE_REQUIRE(gru_seq_supported(p.H), "gru_seq: hidden size must be 32, 64, 128, 256 or 512");
    if (p.B <= 0) return 0;
    const char* e = getenv("SE_B200_GRU_CLUSTER");
    const bool use_cluster = e == nullptr || atoi(e) != 0;
    if (use_cluster) {
        bool launched = false;
        int rc = 0;
        switch (p.H / 32) {
            case 1: rc = launch_cluster_bwd<1>(p, st, &launched); break;
            case 2: rc = launch_cluster_bwd<2>(p, st, &launched); break;
            case 4: rc = launch_cluster_bwd<4>(p, st, &launched); break;
            case 8: rc = launch_cluster_bwd<8>(p, st, &launched); break;
            case 16: rc = launch_cluster_bwd<16>(p, st, &launched); break;
        }
        if (rc) return rc;
        if (launched) return 0;
    }
    const int groups = p.B < 8 ? p.B : 8;
    SE_GRU_DISPATCH(gru_seq_bwd_kernel, GruSeqBwdParams)
    return 2;
}

}  // namespace se
