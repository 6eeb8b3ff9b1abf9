// Persistent GRU recurrence for small batches (training micro-step, CRN_ELU.py:173 through nn.GRU): ONE cooperative
// launch walks all time steps of all chunks of a layer.  Each warp owns one hidden unit and keeps its three rows of
// W_hh (forward) or its column of W_hh (backward) in registers for the whole sequence; the hidden state goes through
// L2 (a few KB) and the steps are separated by grid-wide barriers.  The per-step GEMM launches this replaces spent
// 44 us per step streaming 3 MB of weights out of L2 for a single row of output (profiles/r01_train_launches_*.csv).
//
// Batch layout ("chunk-major"): stream s = n * nb + i is chunk n of utterance i; the state entering chunk n is the
// state leaving chunk n-1 (CRN_ELU.py:173,183-185), detached for the backward (the gradient stops at chunk borders).
#include <cooperative_groups.h>

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
    const int groups = p.nb < 4 ? p.nb : 4;
    SE_GRU_DISPATCH(gru_seq_fwd_kernel, GruSeqParams)
    return 2;
}

int launch_gru_seq_bwd(const GruSeqBwdParams& p, cudaStream_t st) {
    SE_REQUIRE(gru_seq_supported(p.H), "gru_seq: hidden size must be 32, 64, 128, 256 or 512");
    if (p.B <= 0) return 0;
    const int groups = p.B < 8 ? p.B : 8;
    SE_GRU_DISPATCH(gru_seq_bwd_kernel, GruSeqBwdParams)
    return 2;
}

}  // namespace se
