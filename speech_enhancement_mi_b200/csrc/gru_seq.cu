// Persistent GRU recurrence for small batches (training micro-step, CRN_ELU.py:173 through nn.GRU): ONE cooperative
// launch walks all time steps of all chunks of a layer.  Each warp owns one hidden unit and keeps its three rows of
// W_hh (forward) or its column of W_hh (backward) in registers for the whole sequence; the hidden state goes through
// L2 (a few KB) and the steps are separated by grid-wide barriers.  The per-step GEMM launches this replaces spent
// 44 us per step streaming 3 MB of weights out of L2 for a single row of output (profiles/r01_train_launches_*.csv).
//
// Batch layout ("chunk-major"): stream s = n * nb + i is chunk n of utterance i; the state entering chunk n is the
// state leaving chunk n-1 (CRN_ELU.py:173,183-185), detached for the backward (the gradient stops at chunk borders).
#include <cooperative_groups.h>
#include <stdlib.h>

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
    for (int n = 0; n < p.N; ++n) {
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
// 16 warps per CTA, a warp owns two hidden units (its six rows of W_hh in registers, same lane-strided order and the
// same butterfly as above, so the results are bit-identical), every CTA keeps a full copy of h in shared memory
// (double-buffered), and a step ends with the CTA's 32 new values (written in place into its own copy) going to every
// peer's copy as ONE 128-byte bulk copy through distributed shared memory that carries its own completion
// (cp.async.bulk ... complete_tx on the receiver's mbarrier), so a step needs no cluster barrier and no fence:
// barrier.cluster.arrive.release made every warp wait for its hseq store to reach L2 (measured 2.05 us per step, no
// better than the grid barrier), and 4-byte st.async's cost the receiver 512 mbarrier updates per step (1.7 us).  A
// CTA proceeds to step t + 1 when its mbarrier has counted the H - 32 foreign values of step t; a peer can only
// overwrite the buffer a warp is still reading after this CTA's output of the same step has arrived there, i.e. after
// the __syncthreads that follows every warp's read.  The input projections of the next step arrive by
// cp.async while the current one computes.  Utterances are spread over clusters (no dependence between them); a
// cluster serves its utterances inside the same step.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kClWarps = 16;  // x 2 units = 32 units per CTA

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait_cluster(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
    }
}

template <int KPL>  // H / 32 = CTAs per cluster
__global__ void __launch_bounds__(kClWarps * 32, 1) gru_seq_fwd_cluster_kernel(GruSeqParams p, int SG, int upc) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) unsigned long long mbar[2];
    constexpr int H = 32 * KPL;
    float* hb = sm;                       // [2][upc][H]
    float* gst = sm + 2 * upc * H;        // [2][upc][kClWarps][8]: r0 r1 z0 z1 n0 n1 of the warp's two units
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = blockIdx.x % KPL, sg = blockIdx.x / KPL;
    const int j0 = rank * 32 + warp * 2;  // first hidden unit of this warp
    const int T = p.T;
    int mine = 0;  // utterances of this cluster
    for (int u = 0; u < upc; ++u) mine += (sg + u * SG < p.nb) ? 1 : 0;
    const uint32_t tx_bytes = (uint32_t)mine * (H - 32) * sizeof(float);  // the own 32 values are written in place
    float w[6][KPL];  // rows r(j0) r(j0+1) z(j0) z(j0+1) n(j0) n(j0+1)
    float bh = 0.f, bz = 0.f, bn = 0.f;  // lanes 0, 1: b_hh of the lane's unit
#pragma unroll
    for (int gte = 0; gte < 3; ++gte)
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int q = 0; q < KPL; ++q)
                w[2 * gte + u][q] = p.Whh[(long long)(gte * H + j0 + u) * p.Kp + lane + 32 * q];
    if (lane < 2) {
        bh = p.bhh[j0 + lane];
        bz = p.bhh[H + j0 + lane];
        bn = p.bhh[2 * H + j0 + lane];
    }
    // state entering chunk 0 (slot 0 of hseq, written by the caller)
    for (int u = 0; u < upc; ++u) {
        const int i = sg + u * SG;
        for (int k = threadIdx.x; k < H; k += blockDim.x)
            hb[u * H + k] = i < p.nb ? p.hseq[(long long)i * p.hB + k] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    auto stage_gi = [&](int buf, int n, int t) {
        if (lane < 6) {
            for (int u = 0; u < upc; ++u) {
                const int i = sg + u * SG;
                if (i >= p.nb) break;
                const float* g = p.gi + ((long long)n * p.nb + i) * p.giB + (long long)t * 3 * H + (lane >> 1) * H + j0 + (lane & 1);
                const uint32_t dst = smem_addr(gst + ((buf * upc + u) * kClWarps + warp) * 8 + lane);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(g) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_gi(0, 0, 0);
    __syncthreads();
    cluster_arrive();  // every CTA of the cluster runs and has its mbarriers before anyone stores into it
    cluster_wait();
    const int steps = p.N * T;
    for (int step = 0; step < steps; ++step) {
        const int n = step / T, t = step - n * T;
        const int cur = step & 1;
        if (step > 0) mbar_wait_cluster(smem_addr(&mbar[cur]), ((step - 1) >> 1) & 1);
        if (threadIdx.x == 0 && step + 1 < steps)  // the buffer filled during this step
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar[cur ^ 1])), "r"(tx_bytes)
                         : "memory");
        // the projections of THIS step were committed one group earlier
        if (step + 1 < steps) {
            stage_gi(cur ^ 1, (step + 1) / T, (step + 1) % T);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        for (int u = 0; u < upc; ++u) {
            const int i = sg + u * SG;
            if (i >= p.nb) break;
            const long long s = (long long)n * p.nb + i;
            const float* h = hb + (cur * upc + u) * H;
            float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < KPL; ++q) {
                const float hv = h[lane + 32 * q];
#pragma unroll
                for (int r = 0; r < 6; ++r) a[r] = fmaf(w[r][q], hv, a[r]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int r = 0; r < 6; ++r) a[r] += __shfl_xor_sync(0xffffffffu, a[r], off);
            float hnew = 0.f;
            if (lane < 2) {
                const float* gs = gst + ((cur * upc + u) * kClWarps + warp) * 8;
                const float gr = gs[lane], gz = gs[2 + lane], gn = gs[4 + lane];
                const float ar = lane ? a[1] : a[0], az = lane ? a[3] : a[2], an = lane ? a[5] : a[4];
                const float hj = h[j0 + lane];
                const float r = sigmoidf_(gr + ar + bh);
                const float z = sigmoidf_(gz + az + bz);
                const float c = tanhf(gn + r * (an + bn));
                hnew = (1.0f - z) * c + z * hj;
                p.hseq[s * p.hB + (long long)(t + 1) * H + j0 + lane] = hnew;
                if (t == 0 && n > 0) p.hseq[s * p.hB + j0 + lane] = hj;  // slot 0 = state entering the chunk (backward reads it)
                hb[((cur ^ 1) * upc + u) * H + j0 + lane] = hnew;
            }
        }
        if (step + 1 == steps) break;
        // the CTA's 32 new values are in place in its own copy; one 128-byte bulk copy per peer carries them (and their
        // completion) into the others' copies -- 4-byte st.async's cost the receiver one mbarrier update each (512 per
        // step: 1.7 us per step)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (warp == 0 && lane < KPL && lane != rank) {
            for (int u = 0; u < mine; ++u) {
                const uint32_t src = smem_addr(hb + ((cur ^ 1) * upc + u) * H + rank * 32);
                uint32_t remote, rbar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(src), "r"(lane));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_addr(&mbar[cur ^ 1])), "r"(lane));
                asm volatile(
                    "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];" ::"r"(remote),
                    "r"(src), "r"(rbar)
                    : "memory");
            }
        }
    }
    cluster_arrive();  // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

template <int KPL>
int launch_cluster(const GruSeqParams& p, cudaStream_t st, bool* launched) {
    *launched = false;
    auto kernel = gru_seq_fwd_cluster_kernel<KPL>;
    if (KPL > 8) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
    }
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
    auto smem_of = [&](int upc) { return (size_t)(2 * upc * 32 * KPL + 2 * upc * kClWarps * 8) * sizeof(float); };
    cfg.gridDim = dim3(KPL);
    cfg.dynamicSmemBytes = smem_of(1);
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg) != cudaSuccess || ncl < 1) {
        cudaGetLastError();
        return 0;  // this device cannot place the cluster: the caller takes the cooperative kernel
    }
    int SG = p.nb < ncl ? p.nb : ncl;
    int upc = (p.nb + SG - 1) / SG;
    SG = (p.nb + upc - 1) / upc;
    if (smem_of(upc) > 200 * 1024) return 0;
    SE_DYN_SMEM(kernel, smem_of(upc));
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
// UP sequences at a time from t = T - 1 down to 0.  Per step: the cell adjoint of the warp's two units (lanes 0, 1; the
// recurrent part of d loss / d h stays in a register of those lanes, it never leaves the warp), the 3H values of
// d loss / d gh all-gathered into every CTA's shared memory by bulk copies with complete_tx (see the forward),
// then the warp's two dot products over them.  The cell inputs of the next step arrive by cp.async.
// The cooperative kernel above spent two barriers over ~1000 CTAs per step and walked a group's streams one after
// another: 0.44 ms per layer for 24 sequences, 2.3 ms for 192.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBwdUP = 4;  // sequences in flight per cluster

template <int KPL>
__global__ void __launch_bounds__(kClWarps * 32, 1) gru_seq_bwd_cluster_kernel(GruSeqBwdParams p, int NCL) {
    constexpr int H = 32 * KPL, H3 = 3 * H, UP = kBwdUP;
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) unsigned long long mbar[2];
    float* gb = sm;                          // [2][UP][3H] d loss / d gh of the step
    float* stg = sm + 2 * UP * H3;           // [2][UP][kClWarps][16]: per unit a_r a_z a_n h_r h_z h_n h_prev dH
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rank = blockIdx.x % KPL, cl = blockIdx.x / KPL;
    const int j0 = rank * 32 + warp * 2;
    const int T = p.T;
    const int nseq = cl < p.B ? (p.B - cl + NCL - 1) / NCL : 0;  // sequences cl, cl + NCL, ...
    const int ngroups = (nseq + UP - 1) / UP;
    float wc[2][3 * KPL];  // columns j0, j0 + 1 of W_hh: rows n = lane + 32 q
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < 3 * KPL; ++q) wc[c][q] = p.Whh[(long long)(lane + 32 * q) * p.Kp + j0 + c];
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&mbar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // cell inputs of step `it` (group it / T, t = T - 1 - it % T) into staging buffer it & 1
    auto stage = [&](int it) {
        const int grp = it / T, t = T - 1 - (it - grp * T);
        if (lane < 16) {
            const int unit = lane >> 3, v = lane & 7, j = j0 + unit;
#pragma unroll
            for (int u = 0; u < UP; ++u) {
                const int q = grp * UP + u;
                if (q >= nseq) break;
                const long long s = cl + (long long)q * NCL;
                const float* src;
                if (v < 3)
                    src = p.gi + s * p.gB + (long long)t * H3 + v * H + j;
                else if (v < 6)
                    src = p.gh + s * p.gB + (long long)t * H3 + (v - 3) * H + j;
                else if (v == 6)
                    src = p.hseq + s * p.hB + (long long)t * H + j;
                else
                    src = p.dH + s * p.hB + (long long)(t + 1) * H + j;
                const uint32_t dst = smem_addr(stg + ((((it & 1) * UP + u) * kClWarps + warp) * 16 + lane));
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int iters = ngroups * T;
    if (iters > 0) stage(0);
    __syncthreads();
    cluster_arrive();
    cluster_wait();
    float rec[UP];  // lanes 0, 1: d loss / d h_t of the lane's unit arriving through the recurrence
    int x = 0;      // exchanges so far (steps with t > 0)
    for (int it = 0; it < iters; ++it) {
        const int grp = it / T, t = T - 1 - (it - grp * T);
        int act = nseq - grp * UP;
        if (act > UP) act = UP;
        if (t == T - 1) {
#pragma unroll
            for (int u = 0; u < UP; ++u) rec[u] = 0.f;
        }
        if (threadIdx.x == 0 && t > 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&mbar[x & 1])),
                         "r"((uint32_t)(act * (H3 - 96) * sizeof(float)))
                         : "memory");
        if (it + 1 < iters) {
            stage(it + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        float keep[UP];  // dh * z
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            keep[u] = 0.f;
            if (u >= act) continue;
            const long long s = cl + (long long)(grp * UP + u) * NCL;
            float dar = 0.f, daz = 0.f, danr = 0.f;
            if (lane < 2) {
                const float* in = stg + ((((it & 1) * UP + u) * kClWarps + warp) * 16 + 8 * lane);
                const float r = sigmoidf_(in[0] + in[3]);
                const float z = sigmoidf_(in[1] + in[4]);
                const float hn = in[5];
                const float c = tanhf(in[2] + r * hn);
                const float hp = in[6];
                const float dh = in[7] + rec[u];
                const float dan = dh * (1.f - z) * (1.f - c * c);
                daz = dh * (hp - c) * z * (1.f - z);
                dar = dan * hn * r * (1.f - r);
                danr = dan * r;
                keep[u] = dh * z;
                const int j = j0 + lane;
                float* o = p.dgi + s * p.gB + (long long)t * H3;
                float* q = p.dgh + s * p.gB + (long long)t * H3;
                o[j] = dar;
                o[H + j] = daz;
                o[2 * H + j] = dan;
                q[j] = dar;
                q[H + j] = daz;
                q[2 * H + j] = danr;
            }
            if (t > 0 && lane < 2) {  // own segment of the gathered vector, layout [rank][gate][32]
                float* own = gb + ((x & 1) * UP + u) * H3 + rank * 96 + warp * 2 + lane;
                own[0] = dar;
                own[32] = daz;
                own[64] = danr;
            }
        }
        if (t == 0) continue;  // the state entering the chunk is detached (CRN_ELU.py:185)
        // all-gather: one 384-byte bulk copy per peer and sequence carries the CTA's 96 values and their completion
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (warp == 0 && lane < KPL && lane != rank) {
            for (int u = 0; u < act; ++u) {
                const uint32_t src = smem_addr(gb + ((x & 1) * UP + u) * H3 + rank * 96);
                uint32_t remote, rbar;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(src), "r"(lane));
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_addr(&mbar[x & 1])), "r"(lane));
                asm volatile(
                    "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], 384, [%2];" ::"r"(remote),
                    "r"(src), "r"(rbar)
                    : "memory");
            }
        }
        mbar_wait_cluster(smem_addr(&mbar[x & 1]), (x >> 1) & 1);
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            if (u >= act) continue;
            const float* g = gb + ((x & 1) * UP + u) * H3;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int q = 0; q < 3 * KPL; ++q) {
                const float gv = g[((q % KPL) * 3 + q / KPL) * 32 + lane];  // row n = lane + 32 q = gate (q / KPL), rank (q % KPL)
                a0 = fmaf(wc[0][q], gv, a0);
                a1 = fmaf(wc[1][q], gv, a1);
            }
            a0 = warp_sum(a0);
            a1 = warp_sum(a1);
            rec[u] = keep[u] + (lane ? a1 : a0);
        }
        ++x;
    }
    cluster_arrive();  // nobody leaves while a peer may still address its shared memory
    cluster_wait();
}

template <int KPL>
int launch_cluster_bwd(const GruSeqBwdParams& p, cudaStream_t st, bool* launched) {
    *launched = false;
    auto kernel = gru_seq_bwd_cluster_kernel<KPL>;
    if (KPL > 8) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
    }
    const size_t smem = (size_t)(2 * kBwdUP * 3 * 32 * KPL + 2 * kBwdUP * kClWarps * 16) * sizeof(float);
    SE_DYN_SMEM(kernel, smem);
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
    cfg.gridDim = dim3(KPL);
    cfg.dynamicSmemBytes = smem;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg) != cudaSuccess || ncl < 1) {
        cudaGetLastError();
        return 0;
    }
    // one wave of clusters; fewer when that keeps kBwdUP sequences in flight per cluster would leave clusters idle
    int NCL = p.B < ncl ? p.B : ncl;
    cfg.gridDim = dim3(KPL * NCL);
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
