// Backward kernels of the CRN_ELU training micro-step (train.py:195-198: realtime_process -> compute_loss ->
// backward).  The reference gets these from PyTorch autograd; here every adjoint is written out:
//   * dense contractions (causal / transposed convolutions, 1x1 pairs, GRU projections, Linear): weight gradient
//     dW = G^T . im2col(A) and data gradient dA = col2im(G . W), both driven by the SAME GemmParams (gather table,
//     strides, packed weights) as the forward GEMM, so no second description of a layer exists;
//   * GlobalLayerNorm (CRN_ELU.py:37-56), gate (:240), ELU, gated skip blend (:297-306), GRU cell (:173).
// The carried conv buffers and the GRU state are detached in the reference (CRN_ELU.py:185,243): the gradient of a
// chunk never leaves the chunk, so the backward runs batched over all chunks of all utterances at once.
// The two contractions run on mma.sync.m16n8k8 tf32 (3xTF32 head / tail operands beside the exact fp32 forward, one pass
// beside the tf32 forward; the fp32 CUDA-core forms stay selectable); everything else is fp32 on CUDA cores.  The
// training configuration of the reference is one utterance piece per step (config.yaml:92), i.e. a few dozen
// chunk-streams -- these kernels are sized for that, not for the 1024-stream inference path.
#include "se_internal.h"

namespace se {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ long long row_off(int m, int rows_per_stream, int Fo, long long sB, long long sT,
                                             long long sF, int* f_out = nullptr) {
    const int b = m / rows_per_stream;
    const int r = m - b * rows_per_stream;
    const int t = r / Fo;
    const int f = r - t * Fo;
    if (f_out) *f_out = f;
    return b * sB + t * sT + f * sF;
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient: 64 (n) x 64 (k) tile per CTA, reduction over a slice of the rows, atomics into the arena twin
// ---------------------------------------------------------------------------------------------------------------
constexpr int WG_BM = 16;
__global__ void __launch_bounds__(kThreads) wgrad_kernel(GemmParams p, const float* __restrict__ G, StridedRows g,
                                                         float* __restrict__ dW, float* __restrict__ dbias,
                                                         int rows_per_cta) {
    __shared__ __align__(16) float Gs[WG_BM][64 + 4];
    __shared__ __align__(16) float As[WG_BM][64 + 4];
    const int tid = threadIdx.x;
    const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int m_begin = blockIdx.z * rows_per_cta;
    const int m_end = min(p.M, m_begin + rows_per_cta);
    const int tx = tid & 15, ty = tid >> 4;
    const int rps = p.Tn * p.Fo;
    const float* A = reinterpret_cast<const float*>(p.A);
    float acc[4][4] = {};
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m0 = m_begin; m0 < m_end; m0 += WG_BM) {
        {  // G tile: 16 rows x 64 columns, 4 elements per thread (one row, 4 consecutive columns)
            const int r = tid >> 4, c4 = (tid & 15) * 4;
            const int m = m0 + r;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < m_end) {
                int f;
                const long long off = row_off(m, rps, p.Fo, g.sB, g.sT, g.sF, &f);
                const int nlim = (p.odd_tail && f == p.Fo - 1) ? p.N / 2 : p.N;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + c4 + j < nlim) v[j] = G[off + n0 + c4 + j];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) Gs[r][c4 + j] = v[j];
        }
        {  // A tile: 16 rows x 64 k, one gathered float4 per thread
            const int r = tid >> 4, ku = tid & 15;
            const int m = m0 + r, k = k0 + 4 * ku;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < m_end && k < p.K)
                v = *reinterpret_cast<const float4*>(A + row_off(m, rps, p.Fo, p.sB, p.sT, p.sF) + __ldg(p.koff + (k >> 2)));
            *reinterpret_cast<float4*>(&As[r][4 * ku]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WG_BM; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&As[r][tx * 4]);
            const float4 gg = *reinterpret_cast<const float4*>(&Gs[r][ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float gv[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                bsum[i] += gv[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + ty * 4 + i;
        if (n >= p.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < p.K && acc[i][j] != 0.f) atomicAdd(dW + (long long)n * p.K + k, acc[i][j]);
        }
        if (dbias != nullptr && blockIdx.x == 0 && tx == 0 && bsum[i] != 0.f) atomicAdd(dbias + n, bsum[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// data gradient: 64 (rows) x 64 (k) tile per CTA, reduction over the N columns, scatter-add through the gather table
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) dgrad_kernel(GemmParams p, const float* __restrict__ G, StridedRows g,
                                                         float* __restrict__ dA) {
    __shared__ __align__(16) float Gs[16][64 + 4];  // [n][m]
    __shared__ __align__(16) float Ws[16][64 + 4];  // [n][k]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int tx = tid & 15, ty = tid >> 4;
    const int rps = p.Tn * p.Fo;
    const float* W = reinterpret_cast<const float*>(p.W);
    // this thread loads G for row (tid >> 2), columns 4*(tid & 3) .. +3 of every 16-column step
    const int lr = tid >> 2, lc = (tid & 3) * 4;
    long long goff = -1;
    int nlim_l = 0;
    if (m0 + lr < p.M) {
        int f;
        goff = row_off(m0 + lr, rps, p.Fo, g.sB, g.sT, g.sF, &f);
        nlim_l = (p.odd_tail && f == p.Fo - 1) ? p.N / 2 : p.N;
    }
    float acc[4][4] = {};
    for (int n0 = 0; n0 < p.N; n0 += 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + lc + j;
            Gs[lc + j][lr] = (goff >= 0 && n < nlim_l) ? G[goff + n] : 0.f;
        }
        {
            const int n = n0 + (tid >> 4), k = k0 + 4 * (tid & 15);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < p.N && k < p.K) v = *reinterpret_cast<const float4*>(W + (long long)n * p.K + k);
            *reinterpret_cast<float4*>(&Ws[tid >> 4][4 * (tid & 15)]) = v;
        }
        __syncthreads();
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float4 gg = *reinterpret_cast<const float4*>(&Gs[n][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[n][tx * 4]);
            const float gv[4] = {gg.x, gg.y, gg.z, gg.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gv[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int k = k0 + tx * 4;
    if (k >= p.K) return;
    const int ko = __ldg(p.koff + (k >> 2));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
        float* dst = dA + row_off(m, rps, p.Fo, p.sB, p.sT, p.sF) + ko;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (acc[i][j] != 0.f) atomicAdd(dst + j, acc[i][j]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core forms of the two contractions above: warp-level mma.sync.m16n8k8 (tf32 operands, fp32 accumulation).
// SPLIT = true is the error-compensated "3xTF32" product: every operand is stored in shared memory as a tf32 head and a
// tf32 tail (x = hi + lo up to 2^-21 |x|) and a tile step issues lo*hi + hi*lo + hi*hi, which keeps the gradients at fp32
// accuracy (the fp32 parity mode); SPLIT = false is the single-pass tf32 product used when the forward runs in tf32
// too.  Same tiles, gather table and atomics as the CUDA-core kernels, so the two forms are interchangeable per launch
// (SE_B200_BWD_MMA=0 selects the CUDA-core kernels).
// Fragment roles (gq = lane / 4, tq = lane % 4): A (16 x 8): rows gq, gq + 8, columns tq, tq + 4; B (8 x 8): row tq,
// tq + 4, column gq; C (16 x 8): rows gq, gq + 8, columns 2 tq, 2 tq + 1.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TM_BM = 32;   // rows (wgrad) / columns n (dgrad) per shared-memory stage
constexpr int TM_LD8 = 72;  // row pitch = 8 mod 32 banks: the (tq, gq) fragment reads of a [k][m|n] tile are conflict free
constexpr int TM_LD4 = 36;  // row pitch = 4 mod 32 banks: the (gq, tq) fragment reads of a [m][k] tile are conflict free

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma1688(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four values -> tf32 heads (and tails) as one 16-byte shared-memory store each
template <bool SPLIT>
__device__ __forceinline__ void store_split4(uint32_t* hi, uint32_t* lo, const float (&v)[4]) {
    uint4 h, l;
    h.x = to_tf32(v[0]);
    h.y = to_tf32(v[1]);
    h.z = to_tf32(v[2]);
    h.w = to_tf32(v[3]);
    *reinterpret_cast<uint4*>(hi) = h;
    if (SPLIT) {
        l.x = to_tf32(v[0] - __uint_as_float(h.x));
        l.y = to_tf32(v[1] - __uint_as_float(h.y));
        l.z = to_tf32(v[2] - __uint_as_float(h.z));
        l.w = to_tf32(v[3] - __uint_as_float(h.w));
        *reinterpret_cast<uint4*>(lo) = l;
    }
}

// C fragment {(gq, 2tq), (gq, 2tq+1), (gq+8, 2tq), (gq+8, 2tq+1)} of a lane pair -> four consecutive columns of one row:
// even tq keeps row gq, odd tq row gq + 8
__device__ __forceinline__ float4 pair_rows(const float (&c)[4], int tq) {
    const bool odd = tq & 1;
    const float s0 = odd ? c[0] : c[2], s1 = odd ? c[1] : c[3];  // what the partner needs
    const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
    return odd ? make_float4(r0, r1, c[2], c[3]) : make_float4(c[0], c[1], r0, r1);
}

// weight gradient.  Two tile shapes: 64 (n) x 64 (k) per CTA -- warp w owns output rows n0 + 16 (w % 4) and columns
// k0 + 32 (w / 4) .. + 31 (four n8 tiles) -- and, for the few-channel layers (N <= 16: pre-convolutions, first encoder and
// last decoder level, where a 64-row tile multiplies zeros), 16 (n) x 256 (k): every warp owns 32 columns of the same 16
// rows.  The reduction runs over this CTA's slice of the GEMM rows in stages of BM; the row offsets of a stage (two
// integer divisions each) are computed once per row by the first BM threads, one stage ahead, and passed through
// shared memory; the rows themselves travel global -> registers -> shared memory with the loads of stage s + 1 issued
// before the MMAs of stage s.
template <bool SPLIT, bool N16>
__global__ void __launch_bounds__(kThreads) wgrad_mma_kernel(GemmParams p, const float* __restrict__ G, StridedRows g,
                                                             float* __restrict__ dW, float* __restrict__ dbias,
                                                             int rows_per_cta) {
    constexpr int NS = SPLIT ? 2 : 1;
    constexpr int BN = N16 ? 16 : 64, BK = N16 ? 256 : 64, BM = (N16 && SPLIT) ? 16 : TM_BM;
    constexpr int LDG = BN + 8, LDA = BK + 8;  // = 8 mod 32 banks (24 works as well: tq * 24 + gq are distinct banks)
    constexpr int UA = BK / 4, RA = kThreads / UA, PA = BM / RA;               // A: units per row, rows per pass, passes
    constexpr int UG = BN / 4, RG = kThreads / UG, PG = BM > RG ? BM / RG : 1;  // G likewise
    __shared__ __align__(16) uint32_t Gs[NS][BM][LDG];  // [m][n]
    __shared__ __align__(16) uint32_t As[NS][BM][LDA];  // [m][k]
    __shared__ long long s_aoff[2][BM], s_goff[2][BM];
    __shared__ int s_nlim[2][BM];
    __shared__ float s_bias[BN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int k0 = blockIdx.x * BK, n0 = blockIdx.y * BN;
    const int m_begin = blockIdx.z * rows_per_cta;
    const int m_end = min(p.M, m_begin + rows_per_cta);
    const int nt = N16 ? 0 : (warp & 3) * 16, kh = N16 ? warp * 32 : (warp >> 2) * 32;
    const int rps = p.Tn * p.Fo;
    const float* A = reinterpret_cast<const float*>(p.A);
    const bool want_bias = dbias != nullptr && blockIdx.x == 0;
    float acc[4][4] = {};
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    if (tid < BN) s_bias[tid] = 0.f;
    const int au = tid % UA, ar = tid / UA, gu = tid % UG, gr = tid / UG;  // loader roles
    const bool k_ok = k0 + 4 * au < p.K;
    const int koff_l = k_ok ? __ldg(p.koff + ((k0 + 4 * au) >> 2)) : 0;
    auto row_offsets = [&](int m0, int buf) {  // threads < BM: one row each
        if (tid < BM) {
            const int m = m0 + tid;
            int f = 0;
            long long go = 0, ao = 0;
            int nl = 0;
            if (m < m_end) {
                go = row_off(m, rps, p.Fo, g.sB, g.sT, g.sF, &f);
                ao = row_off(m, rps, p.Fo, p.sB, p.sT, p.sF);
                nl = (p.odd_tail && f == p.Fo - 1) ? p.N / 2 : p.N;
            }
            s_goff[buf][tid] = go;
            s_aoff[buf][tid] = ao;
            s_nlim[buf][tid] = nl;  // 0 beyond the slice: nothing is loaded
        }
    };
    float v[PG][4];
    float4 a[PA];
    auto load_stage = [&](int buf) {
#pragma unroll
        for (int pass = 0; pass < PG; ++pass) {
            const int r = gr + RG * pass;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[pass][j] = 0.f;
            if (r < BM) {
                const int nlim = s_nlim[buf][r];
                const long long off = s_goff[buf][r];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n0 + 4 * gu + j < nlim) v[pass][j] = G[off + n0 + 4 * gu + j];
            }
        }
#pragma unroll
        for (int pass = 0; pass < PA; ++pass) {
            const int r = ar + RA * pass;
            a[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k_ok && s_nlim[buf][r] > 0) a[pass] = *reinterpret_cast<const float4*>(A + s_aoff[buf][r] + koff_l);
        }
    };
    row_offsets(m_begin, 0);
    __syncthreads();
    if (m_begin < m_end) load_stage(0);
    int buf = 0;
    for (int m0 = m_begin; m0 < m_end; m0 += BM, buf ^= 1) {
#pragma unroll
        for (int pass = 0; pass < PG; ++pass) {
            const int r = gr + RG * pass;
            if (r < BM) {
#pragma unroll
                for (int j = 0; j < 4; ++j) bsum[j] += v[pass][j];
                store_split4<SPLIT>(&Gs[0][r][4 * gu], &Gs[NS - 1][r][4 * gu], v[pass]);
            }
        }
#pragma unroll
        for (int pass = 0; pass < PA; ++pass) {
            const int r = ar + RA * pass;
            const float av[4] = {a[pass].x, a[pass].y, a[pass].z, a[pass].w};
            store_split4<SPLIT>(&As[0][r][4 * au], &As[NS - 1][r][4 * au], av);
        }
        if (m0 + BM < m_end) row_offsets(m0 + BM, buf ^ 1);
        __syncthreads();
        if (m0 + BM < m_end) load_stage(buf ^ 1);
#pragma unroll
        for (int ms = 0; ms < BM; ms += 8) {
            uint32_t ah[4], al[4];
            ah[0] = Gs[0][ms + tq][nt + gq];
            ah[1] = Gs[0][ms + tq][nt + gq + 8];
            ah[2] = Gs[0][ms + tq + 4][nt + gq];
            ah[3] = Gs[0][ms + tq + 4][nt + gq + 8];
            if (SPLIT) {
                al[0] = Gs[NS - 1][ms + tq][nt + gq];
                al[1] = Gs[NS - 1][ms + tq][nt + gq + 8];
                al[2] = Gs[NS - 1][ms + tq + 4][nt + gq];
                al[3] = Gs[NS - 1][ms + tq + 4][nt + gq + 8];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = kh + 8 * j + gq;
                const uint32_t bh0 = As[0][ms + tq][kc], bh1 = As[0][ms + tq + 4][kc];
                if (SPLIT) {
                    const uint32_t bl0 = As[NS - 1][ms + tq][kc], bl1 = As[NS - 1][ms + tq + 4][kc];
                    mma1688(acc[j], al, bh0, bh1);
                    mma1688(acc[j], ah, bl0, bl1);
                }
                mma1688(acc[j], ah, bh0, bh1);
            }
        }
        __syncthreads();
    }
    // a lane pair (tq even / odd) holds columns 4 (tq / 2) .. + 3 of rows gq and gq + 8: one exchange gives the even
    // lane the four values of row gq and the odd lane those of row gq + 8 -> ONE 16-byte reduction per row and unit
    // instead of four (the L2 reduction rate, not the arithmetic, bounds the scatter)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v4 = pair_rows(acc[j], tq);
        const int k = k0 + kh + 8 * j + 4 * (tq >> 1);
        const int n = n0 + nt + gq + 8 * (tq & 1);
        if (k < p.K && n < p.N && (v4.x != 0.f || v4.y != 0.f || v4.z != 0.f || v4.w != 0.f))
            atomicAdd(reinterpret_cast<float4*>(dW + (long long)n * p.K + k), v4);
    }
    if (want_bias) {  // the raw fp32 values, not their tf32 heads
        if (gr < BM) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (bsum[j] != 0.f) atomicAdd(&s_bias[4 * gu + j], bsum[j]);
        }
        __syncthreads();
        if (tid < BN && n0 + tid < p.N && s_bias[tid] != 0.f) atomicAdd(dbias + n0 + tid, s_bias[tid]);
    }
}

// data gradient: 64 (rows) x 64 (k) tile per CTA; warp w owns rows m0 + 16 (w % 4) and columns k0 + 32 (w / 4) .. + 31;
// the reduction runs over the N columns of G in stages of TM_BM
template <bool SPLIT>
__global__ void __launch_bounds__(kThreads) dgrad_mma_kernel(GemmParams p, const float* __restrict__ G, StridedRows g,
                                                             float* __restrict__ dA, int n_per_cta) {
    constexpr int NS = SPLIT ? 2 : 1;
    __shared__ __align__(16) uint32_t Gs[NS][64][TM_LD4];     // [m][n]
    __shared__ __align__(16) uint32_t Ws[NS][TM_BM][TM_LD8];  // [n][k]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int m0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int mt = (warp & 3) * 16, kh = (warp >> 2) * 32;
    const int rps = p.Tn * p.Fo;
    const float* W = reinterpret_cast<const float*>(p.W);
    // loader role for G: row tid / 4, columns 8 (tid % 4) .. + 7 of every stage
    const int lr = tid >> 2, lc = (tid & 3) * 8;
    long long goff = -1;
    int nlim_l = 0;
    if (m0 + lr < p.M) {
        int f;
        goff = row_off(m0 + lr, rps, p.Fo, g.sB, g.sT, g.sF, &f);
        nlim_l = (p.odd_tail && f == p.Fo - 1) ? p.N / 2 : p.N;
    }
    float acc[4][4] = {};
    // blockIdx.z: slice of the reduction (few-row GEMMs such as the GRU projections would otherwise run on 64 CTAs)
    const int n_begin = blockIdx.z * n_per_cta, n_end = min(p.N, n_begin + n_per_cta);
    float gv[2][4];
    float4 wv4[TM_BM / 16];
    auto load_stage = [&](int n0) {  // next stage into registers, in flight during the MMAs of the current one
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + lc + 4 * half + j;
                gv[half][j] = (goff >= 0 && n < nlim_l && n < n_end) ? G[goff + n] : 0.f;
            }
#pragma unroll
        for (int pass = 0; pass < TM_BM / 16; ++pass) {
            const int n = n0 + (tid >> 4) + 16 * pass, k = k0 + 4 * (tid & 15);
            wv4[pass] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < n_end && k < p.K) wv4[pass] = *reinterpret_cast<const float4*>(W + (long long)n * p.K + k);
        }
    };
    if (n_begin < n_end) load_stage(n_begin);
    for (int n0 = n_begin; n0 < n_end; n0 += TM_BM) {
#pragma unroll
        for (int half = 0; half < 2; ++half)
            store_split4<SPLIT>(&Gs[0][lr][lc + 4 * half], &Gs[NS - 1][lr][lc + 4 * half], gv[half]);
#pragma unroll
        for (int pass = 0; pass < TM_BM / 16; ++pass) {
            const int nr = (tid >> 4) + 16 * pass;
            const float wv[4] = {wv4[pass].x, wv4[pass].y, wv4[pass].z, wv4[pass].w};
            store_split4<SPLIT>(&Ws[0][nr][4 * (tid & 15)], &Ws[NS - 1][nr][4 * (tid & 15)], wv);
        }
        __syncthreads();
        if (n0 + TM_BM < n_end) load_stage(n0 + TM_BM);
#pragma unroll
        for (int ns = 0; ns < TM_BM; ns += 8) {
            uint32_t ah[4], al[4];
            ah[0] = Gs[0][mt + gq][ns + tq];
            ah[1] = Gs[0][mt + gq + 8][ns + tq];
            ah[2] = Gs[0][mt + gq][ns + tq + 4];
            ah[3] = Gs[0][mt + gq + 8][ns + tq + 4];
            if (SPLIT) {
                al[0] = Gs[NS - 1][mt + gq][ns + tq];
                al[1] = Gs[NS - 1][mt + gq + 8][ns + tq];
                al[2] = Gs[NS - 1][mt + gq][ns + tq + 4];
                al[3] = Gs[NS - 1][mt + gq + 8][ns + tq + 4];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = kh + 8 * j + gq;
                const uint32_t bh0 = Ws[0][ns + tq][kc], bh1 = Ws[0][ns + tq + 4][kc];
                if (SPLIT) {
                    const uint32_t bl0 = Ws[NS - 1][ns + tq][kc], bl1 = Ws[NS - 1][ns + tq + 4][kc];
                    mma1688(acc[j], al, bh0, bh1);
                    mma1688(acc[j], ah, bl0, bl1);
                }
                mma1688(acc[j], ah, bh0, bh1);
            }
        }
        __syncthreads();
    }
    const int m = m0 + mt + gq + 8 * (tq & 1);
    const long long roff = m < p.M ? row_off(m, rps, p.Fo, p.sB, p.sT, p.sF) : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v = pair_rows(acc[j], tq);  // row m, the whole 16-byte unit k .. k + 3 of the gather table
        const int k = k0 + kh + 8 * j + 4 * (tq >> 1);
        if (k < p.K && m < p.M && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f))
            atomicAdd(reinterpret_cast<float4*>(dA + roff + __ldg(p.koff + (k >> 2))), v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// GlobalLayerNorm backward.  out = (y - mu) / D * w + b,  D = sqrt(var + eps) + eps (teacher) | sqrt(var) + eps (student)
//   g = dout * w;  dy = (g - mean(g)) / D - (y - mu) * sum(g (y - mu)) / (N D^2 s),  s = dD/dvar^-1 / 2 = sqrt(var [+ eps])
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gln_consts(const double* stats, int b, double count, int student, float& mean,
                                           float& inv, float& s) {
    const double mu = stats[2 * b] / count;
    double var = stats[2 * b + 1] / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float varf = (float)var;
    s = student ? sqrtf(varf) : sqrtf(varf + 1e-8f);
    inv = 1.0f / (s + 1e-8f);
    mean = (float)mu;
}

__global__ void __launch_bounds__(kThreads) gln_bwd_reduce_kernel(GlnBwdParams p) {
    __shared__ float sdw[kThreads], sdb[kThreads];
    __shared__ double red[2][kThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int Ce = p.per_feature ? p.F * p.C : p.C;  // period of the affine index
    const int npos = p.per_feature ? p.T : p.T * p.F;
    const int lanes = Ce < kThreads ? Ce : kThreads;
    const int groups = kThreads / lanes;
    const int lane = tid % lanes, group = tid / lanes;
    float mean, inv, s;
    gln_consts(p.stats, b, p.count, p.student, mean, inv, s);
    if (tid < lanes) {
        sdw[tid] = 0.f;
        sdb[tid] = 0.f;
    }
    __syncthreads();
    double sg = 0.0, sgy = 0.0;
    if (group < groups) {
        for (int k = lane; k < Ce; k += lanes) {
            const float w = p.w[k];
            const int fk = p.per_feature ? k / p.C : 0, ck = p.per_feature ? k - fk * p.C : k;
            float dwk = 0.f, dbk = 0.f;
            for (int pos = group + groups * blockIdx.y; pos < npos; pos += groups * gridDim.y) {
                int t, f;
                if (p.per_feature) {
                    t = pos;
                    f = fk;
                } else {
                    t = pos / p.F;
                    f = pos - t * p.F;
                }
                const float go = p.g[b * p.gB + t * p.gT + f * p.gF + ck];
                const float yc = p.y[((long long)b * npos + pos) * Ce + k] - mean;
                dwk = fmaf(go, yc * inv, dwk);
                dbk += go;
                const float gw = go * w;
                sg += gw;
                sgy += (double)gw * yc;
            }
            if (Ce <= kThreads) {
                atomicAdd(&sdw[lane], dwk);
                atomicAdd(&sdb[lane], dbk);
            } else {
                atomicAdd(p.dw + k, dwk);
                atomicAdd(p.db + k, dbk);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sg += __shfl_xor_sync(0xffffffffu, sg, off);
        sgy += __shfl_xor_sync(0xffffffffu, sgy, off);
    }
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = sg;
        red[1][tid >> 5] = sgy;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0, c = 0;
        for (int w = 0; w < kThreads / 32; ++w) {
            a += red[0][w];
            c += red[1][w];
        }
        atomicAdd(p.red + 2 * b, a);  // zeroed by the launcher; gridDim.y CTAs share a stream
        atomicAdd(p.red + 2 * b + 1, c);
    }
    if (Ce <= kThreads && tid < lanes) {
        atomicAdd(p.dw + tid, sdw[tid]);
        atomicAdd(p.db + tid, sdb[tid]);
    }
}

__global__ void __launch_bounds__(kThreads) gln_bwd_apply_kernel(GlnBwdParams p) {
    const int b = blockIdx.y;
    const int Ce = p.per_feature ? p.F * p.C : p.C;
    const int n = p.T * p.F * p.C;
    float mean, inv, s;
    gln_consts(p.stats, b, p.count, p.student, mean, inv, s);
    const float mg = (float)(p.red[2 * b] / p.count);
    const float coef = (s > 0.f) ? (float)(p.red[2 * b + 1] / p.count) * inv * inv / s : 0.f;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int c = e % p.C;
        const int pf = e / p.C;  // t*F + f
        const int t = pf / p.F, f = pf - t * p.F;
        const int k = p.per_feature ? f * p.C + c : c;
        (void)Ce;
        const float y = p.y[(long long)b * n + e];
        const float gw = p.g[b * p.gB + t * p.gT + f * p.gF + c] * p.w[k];
        float d = (gw - mg) * inv - (y - mean) * coef;
        if (p.elu) d *= (y > 0.f ? 1.f : y + 1.f);
        p.dy[((long long)b * p.T * p.F + pf) * p.oC + c * p.ostep + p.ooff] = d;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// gated skip blend backward (CRN_ELU.py:297-306): out = m * rr + (1 - m) * o,  m = sigmoid(GLN_r(rm)), o = GLN(y)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) blend_bwd_kernel(BlendBwdParams p) {
    const int Fm = p.Fs > p.Fy ? p.Fs : p.Fy;
    const long long per = (long long)p.T * Fm * p.C;
    const long long total = per * p.B;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        long long r = i - b * per;
        const int c = (int)(r % p.C);
        r /= p.C;
        const int f = (int)(r % Fm), t = (int)(r / Fm);
        float mean, inv, s, mr, ir, sr;
        gln_consts(p.stats, b, p.count, p.student, mean, inv, s);
        gln_consts(p.stats_r, b, p.count_r, p.student, mr, ir, sr);
        float go = 0.f, m = 0.f;
        if (f < p.Fs) {
            go = p.g[b * p.gB + t * p.gT + f * p.gF + c];
            const long long ri = (((long long)b * p.T + t) * p.Fs + f) * p.C + c;
            m = sigmoidf_((p.rm[ri] - mr) * ir * p.wr[c] + p.br[c]);
            const float rr = p.rr[ri];
            float o = 0.f;
            if (f < p.Fy) o = (p.y[(((long long)b * p.T + t) * p.Fy + f) * p.C + c] - mean) * inv * p.w[c] + p.b[c];
            p.g_r[ri] = go * (rr - o) * m * (1.f - m);
            p.G2[ri * 2 + 1] = go * m * (rr > 0.f ? 1.f : rr + 1.f);
        }
        if (f < p.Fy) p.g_o[(((long long)b * p.T + t) * p.Fy + f) * p.C + c] = go * (1.f - m);
    }
}

__global__ void __launch_bounds__(kThreads) gate_bwd_kernel(float* __restrict__ uv, const float* __restrict__ dy,
                                                            long long n) {  // n = rows * C
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float u = uv[2 * i], v = uv[2 * i + 1], d = dy[i];
        const float sg = sigmoidf_(v);
        uv[2 * i] = d * sg;
        uv[2 * i + 1] = d * u * sg * (1.f - sg);
    }
}

__global__ void __launch_bounds__(kThreads) elu_bwd_kernel(float* __restrict__ de, const float* __restrict__ e,
                                                           long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = e[i];
        de[i] *= (v > 0.f ? 1.f : v + 1.f);
    }
}

__global__ void __launch_bounds__(kThreads) add_strided_kernel(float* dst, StridedRows d, const float* src,
                                                               StridedRows s, int B, int T, int F, int C) {
    const long long total = (long long)B * T * F * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long r = i / C;
        const int f = (int)(r % F);
        r /= F;
        const int t = (int)(r % T), b = (int)(r / T);
        dst[b * d.sB + t * d.sT + f * d.sF + c] += src[b * s.sB + t * s.sT + f * s.sF + c];
    }
}

// destination-contiguous order when the destination is channels-last, i.e. c fastest (the injection direction); the
// export direction (destination [C][F][T]) walks t fastest: the flag picks the index decomposition
__global__ void __launch_bounds__(kThreads) permute4_kernel(float* __restrict__ dst, Strides4 d,
                                                            const float* __restrict__ src, Strides4 s, int B, int T,
                                                            int F, int C, int add, int t_fastest) {
    const long long total = (long long)B * T * F * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int b, t, f, c;
        if (t_fastest) {
            t = (int)(i % T);
            long long r = i / T;
            f = (int)(r % F);
            r /= F;
            c = (int)(r % C);
            b = (int)(r / C);
        } else {
            c = (int)(i % C);
            long long r = i / C;
            f = (int)(r % F);
            r /= F;
            t = (int)(r % T);
            b = (int)(r / T);
        }
        const float v = src[b * s.sB + t * s.sT + f * s.sF + c * s.sC];
        float* o = dst + b * d.sB + t * d.sT + f * d.sF + c * d.sC;
        *o = add ? *o + v : v;
    }
}

__global__ void __launch_bounds__(kThreads) gru_bwd_pw_kernel(const float* __restrict__ gi, long long giB,
                                                              const float* __restrict__ gh, long long ghB,
                                                              const float* __restrict__ hprev, long long hB,
                                                              const float* __restrict__ dH, long long dHB,
                                                              float* __restrict__ dhrec, float* __restrict__ dgi,
                                                              float* __restrict__ dgh, long long dgB, int B, int H) {
    const long long total = (long long)B * H;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % H), b = (int)(i / H);
        const float* a = gi + b * giB;
        const float* h = gh + b * ghB;
        const float r = sigmoidf_(a[j] + h[j]);
        const float z = sigmoidf_(a[H + j] + h[H + j]);
        const float hn = h[2 * H + j];
        const float n = tanhf(a[2 * H + j] + r * hn);
        const float hp = hprev[b * hB + j];
        const float dh = dH[b * dHB + j] + dhrec[(long long)b * H + j];
        const float dan = dh * (1.f - z) * (1.f - n * n);
        const float daz = dh * (hp - n) * z * (1.f - z);
        const float dar = dan * hn * r * (1.f - r);
        float* o = dgi + b * dgB;
        float* q = dgh + b * dgB;
        o[j] = dar;
        o[H + j] = daz;
        o[2 * H + j] = dan;
        q[j] = dar;
        q[H + j] = daz;
        q[2 * H + j] = dan * r;
        dhrec[(long long)b * H + j] = dh * z;
    }
}

__global__ void __launch_bounds__(kThreads) copy_rows_kernel(float* dst, long long dB, const float* src, long long sB,
                                                             int count) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        dst[b * dB + i] = src ? src[b * sB + i] : 0.f;
}

__global__ void __launch_bounds__(kThreads) arena_gather_kernel(float* arena, const int* map, const float* theta,
                                                                long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = map[i];
        arena[i] = s > 0 ? theta[s - 1] : 0.f;
    }
}
__global__ void __launch_bounds__(kThreads) arena_scatter_add_kernel(const float* garena, const int* map, float* grad,
                                                                     long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = map[i];
        const float v = garena[i];
        if (s > 0 && v != 0.f) atomicAdd(grad + (s - 1), v);
    }
}

inline int grid_for(long long n) {
    long long g = (n + kThreads - 1) / kThreads;
    const long long cap = 148LL * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

int launch_wgrad(const GemmParams& p, const float* G, StridedRows g, float* dW, float* dbias, cudaStream_t st, int mode) {
    if (p.M <= 0) return 0;
    SE_REQUIRE(!p.a_half, "wgrad: fp32 operands only");
    const bool n16 = mode != BWD_CUDA_CORES && p.N <= 16;  // few-channel layers: 16 x 256 tiles
    const int bm = mode == BWD_CUDA_CORES ? WG_BM : (n16 && mode == BWD_3XTF32) ? 16 : TM_BM;
    const int bk = n16 ? 256 : 64, bn = n16 ? 16 : 64;
    const int kt = (p.K + bk - 1) / bk, nt = (p.N + bn - 1) / bn;
    int splits = (148 * (mode == BWD_CUDA_CORES ? 4 : 6) + kt * nt - 1) / (kt * nt);
    const int max_splits = (p.M + 4 * bm - 1) / (4 * bm);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int rows = (p.M + splits - 1) / splits;
    rows = (rows + bm - 1) / bm * bm;
    splits = (p.M + rows - 1) / rows;
    const dim3 grid(kt, nt, splits);
    if (mode == BWD_CUDA_CORES)
        wgrad_kernel<<<grid, kThreads, 0, st>>>(p, G, g, dW, dbias, rows);
    else if (mode == BWD_TF32 && n16)
        wgrad_mma_kernel<false, true><<<grid, kThreads, 0, st>>>(p, G, g, dW, dbias, rows);
    else if (mode == BWD_TF32)
        wgrad_mma_kernel<false, false><<<grid, kThreads, 0, st>>>(p, G, g, dW, dbias, rows);
    else if (n16)
        wgrad_mma_kernel<true, true><<<grid, kThreads, 0, st>>>(p, G, g, dW, dbias, rows);
    else
        wgrad_mma_kernel<true, false><<<grid, kThreads, 0, st>>>(p, G, g, dW, dbias, rows);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_dgrad(const GemmParams& p, const float* G, StridedRows g, float* dA, cudaStream_t st, int mode) {
    if (p.M <= 0) return 0;
    SE_REQUIRE(!p.a_half, "dgrad: fp32 operands only");
    dim3 grid((p.M + 63) / 64, (p.K + 63) / 64);
    if (mode == BWD_CUDA_CORES) {
        dgrad_kernel<<<grid, kThreads, 0, st>>>(p, G, g, dA);
    } else {
        // at least two CTAs per SM: split the reduction over N when the row / k tiles are few (results are scatter-added anyway)
        const int tiles = grid.x * grid.y, stages = (p.N + TM_BM - 1) / TM_BM;
        int nsplit = (2 * 148 + tiles - 1) / tiles;
        if (nsplit > stages) nsplit = stages;
        if (nsplit < 1) nsplit = 1;
        const int n_per_cta = (stages + nsplit - 1) / nsplit * TM_BM;
        grid.z = (p.N + n_per_cta - 1) / n_per_cta;
        if (mode == BWD_TF32)
            dgrad_mma_kernel<false><<<grid, kThreads, 0, st>>>(p, G, g, dA, n_per_cta);
        else
            dgrad_mma_kernel<true><<<grid, kThreads, 0, st>>>(p, G, g, dA, n_per_cta);
    }
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_gln_bwd(const GlnBwdParams& p, cudaStream_t st) {
    if (p.B <= 0) return 0;
    const int Ce = p.per_feature ? p.F * p.C : p.C;
    SE_CUDA_OK(cudaMemsetAsync(p.red, 0, (size_t)2 * p.B * sizeof(double), st));
    const int npos = p.per_feature ? p.T : p.T * p.F;
    const int lanes = Ce < kThreads ? Ce : kThreads;
    int split = npos / (4 * (kThreads / lanes));  // at least 4 positions per thread group
    if (split > 32) split = 32;
    if (split * p.B > 148 * 8) split = 148 * 8 / p.B;
    if (split < 1) split = 1;
    gln_bwd_reduce_kernel<<<dim3(p.B, split), kThreads, 0, st>>>(p);
    const int n = p.T * p.F * p.C;
    int gx = (n + kThreads * 4 - 1) / (kThreads * 4);
    gln_bwd_apply_kernel<<<dim3(gx < 1 ? 1 : gx, p.B), kThreads, 0, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_blend_bwd(const BlendBwdParams& p, cudaStream_t st) {
    if (p.B <= 0) return 0;
    const int Fm = p.Fs > p.Fy ? p.Fs : p.Fy;
    blend_bwd_kernel<<<grid_for((long long)p.B * p.T * Fm * p.C), kThreads, 0, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_gate_bwd(float* uv, const float* dy, long long rows, int C, cudaStream_t st) {
    if (rows <= 0) return 0;
    gate_bwd_kernel<<<grid_for(rows * C), kThreads, 0, st>>>(uv, dy, rows * C);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_elu_bwd(float* de, const float* e, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    elu_bwd_kernel<<<grid_for(n), kThreads, 0, st>>>(de, e, n);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_add_strided(float* dst, StridedRows d, const float* src, StridedRows s, int B, int T, int F, int C,
                       cudaStream_t st) {
    if (B <= 0) return 0;
    add_strided_kernel<<<grid_for((long long)B * T * F * C), kThreads, 0, st>>>(dst, d, src, s, B, T, F, C);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_permute4(float* dst, Strides4 d, const float* src, Strides4 s, int B, int T, int F, int C, int add,
                    cudaStream_t st) {
    if (B <= 0) return 0;
    permute4_kernel<<<grid_for((long long)B * T * F * C), kThreads, 0, st>>>(dst, d, src, s, B, T, F, C, add,
                                                                           d.sT == 1 ? 1 : 0);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_gru_bwd_pw(const float* gi, long long giB, const float* gh, long long ghB, const float* hprev, long long hB,
                      const float* dH, long long dHB, float* dhrec, float* dgi, float* dgh, long long dgB, int B, int H,
                      cudaStream_t st) {
    if (B <= 0) return 0;
    gru_bwd_pw_kernel<<<grid_for((long long)B * H), kThreads, 0, st>>>(gi, giB, gh, ghB, hprev, hB, dH, dHB, dhrec, dgi,
                                                                      dgh, dgB, B, H);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_copy_rows(float* dst, long long dB, const float* src, long long sB, int count, int nb, cudaStream_t st) {
    if (nb <= 0 || count <= 0) return 0;
    SE_REQUIRE(nb <= 65535, "copy_rows: at most 65535 rows per launch");
    int gx = (count + kThreads - 1) / kThreads;
    if (gx > 64) gx = 64;
    copy_rows_kernel<<<dim3(gx, nb), kThreads, 0, st>>>(dst, dB, src, sB, count);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_arena_gather(float* arena, const int* map, const float* theta, long long n, cudaStream_t st) {
    arena_gather_kernel<<<grid_for(n), kThreads, 0, st>>>(arena, map, theta, n);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_arena_scatter_add(const float* garena, const int* map, float* grad, long long n, cudaStream_t st) {
    arena_scatter_add_kernel<<<grid_for(n), kThreads, 0, st>>>(garena, map, grad, n);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace se
