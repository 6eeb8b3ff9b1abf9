// Exact-mode (fp32 FMA) gathered GEMM.  One kernel serves every dense contraction of the CRN path through the
// row/offset description in GemmParams; it is the bit-careful companion of the tcgen05 path in gemm_tc.cu and the
// only contraction kernel used when se_crn_config.precision == SE_PRECISION_FP32.
#include "se_internal.h"

namespace se {

namespace {

constexpr int kThreads = 256;
constexpr int BK = 16;

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// BM x BN tile, 4x4 outputs per thread (BM*BN == 4096)
template <int BM, int BN>
__global__ void __launch_bounds__(kThreads) gemm_fp32_kernel(GemmParams p) {
    static_assert(BM * BN == 16 * kThreads, "tile must give 4x4 outputs per thread");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Ws[BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    constexpr int TX = BN / 4;  // threads along n
    const int tx = tid % TX;
    const int ty = tid / TX;

    // ---- per-thread gather assignments: units of 4 floats -----------------------------------------------------
    constexpr int A_UNITS = BM * BK / 4;                          // per k-tile
    constexpr int A_PER_THREAD = (A_UNITS + kThreads - 1) / kThreads;
    constexpr int W_UNITS = BN * BK / 4;
    constexpr int W_PER_THREAD = (W_UNITS + kThreads - 1) / kThreads;
    const int rowsPerStream = p.Tn * p.Fo;

    const float* a_base[A_PER_THREAD];
    int a_row[A_PER_THREAD], a_ku[A_PER_THREAD];
#pragma unroll
    for (int i = 0; i < A_PER_THREAD; ++i) {
        const int u = tid + i * kThreads;
        a_row[i] = u / (BK / 4);
        a_ku[i] = u % (BK / 4);
        const int m = m0 + a_row[i];
        a_base[i] = nullptr;
        if (u < A_UNITS && m < p.M) {
            const int b = m / rowsPerStream;
            const int r = m - b * rowsPerStream;
            const int t = r / p.Fo;
            const int f = r - t * p.Fo;
            a_base[i] = reinterpret_cast<const float*>(p.A) + b * p.sB + t * p.sT + f * p.sF;
        }
    }

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
        for (int i = 0; i < A_PER_THREAD; ++i) {
            const int u = tid + i * kThreads;
            if (u < A_UNITS) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int k = k0 + a_ku[i] * 4;
                if (a_base[i] != nullptr && k < p.K)
                    v = *reinterpret_cast<const float4*>(a_base[i] + __ldg(p.koff + (k >> 2)));
                As[a_ku[i] * 4 + 0][a_row[i]] = v.x;
                As[a_ku[i] * 4 + 1][a_row[i]] = v.y;
                As[a_ku[i] * 4 + 2][a_row[i]] = v.z;
                As[a_ku[i] * 4 + 3][a_row[i]] = v.w;
            }
        }
#pragma unroll
        for (int i = 0; i < W_PER_THREAD; ++i) {
            const int u = tid + i * kThreads;
            if (u < W_UNITS) {
                const int row = u / (BK / 4), ku = u % (BK / 4);
                const int n = n0 + row, k = k0 + ku * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n < p.Npad && k < p.K)
                    v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.W) + (long long)n * p.K + k);
                Ws[ku * 4 + 0][row] = v.x;
                Ws[ku * 4 + 1][row] = v.y;
                Ws[ku * 4 + 2][row] = v.z;
                Ws[ku * 4 + 3][row] = v.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue ------------------------------------------------------------------------------------------
    const int n = n0 + tx * 4;
    float bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = (n + j < p.Npad) ? __ldg(p.bias + n + j) : 0.f;

    const bool paired = (p.epi == EPI_GATE_STATS || p.epi == EPI_SKIP);
    const bool want_stats = (p.epi == EPI_ELU_STATS || p.epi == EPI_GATE_STATS || p.epi == EPI_SKIP);

    float s_acc = 0.f, ss_acc = 0.f;
    int s_b = -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        float s = 0.f, ss = 0.f;
        int b = -1;
        if (m < p.M) {
            b = m / rowsPerStream;
            const int r = m - b * rowsPerStream;
            const int t = r / p.Fo;
            const int f = r - t * p.Fo;
            float* o = p.out + b * p.oB + t * p.oT + f * p.oF;
            const int nlim = (p.odd_tail && f == p.Fo - 1) ? p.N / 2 : p.N;
            if (!paired) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < nlim) {
                        float v = acc[i][j] + bias[j];
                        if (p.epi != EPI_BIAS) v = elu1(v);
                        o[n + j] = v;
                        s += v;
                        ss += v * v;
                    }
                }
            } else {
                float* o2 = (p.epi == EPI_SKIP) ? p.out2 + b * p.o2B + t * p.o2T + f * p.o2F : nullptr;
#pragma unroll
                for (int j = 0; j < 4; j += 2) {
                    if (n + j + 1 < p.N) {
                        const float a0 = acc[i][j] + bias[j];
                        const float a1 = acc[i][j + 1] + bias[j + 1];
                        const int c = (n + j) >> 1;
                        float v;
                        if (p.epi == EPI_GATE_STATS) {
                            v = a0 * sigmoidf_(a1);
                        } else {
                            v = a0;
                            o2[c] = elu1(a1);
                        }
                        o[c] = v;
                        s += v;
                        ss += v * v;
                    }
                }
            }
        }
        if (want_stats) {
            // reduce over the TX threads that share this row (consecutive lanes; TX divides 32)
#pragma unroll
            for (int off = TX / 2; off > 0; off >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, off);
                ss += __shfl_xor_sync(0xffffffffu, ss, off);
            }
            if (tx == 0 && b >= 0) {
                if (b != s_b) {
                    if (s_b >= 0) {
                        atomicAdd(p.stats + 2 * s_b, (double)s_acc);
                        atomicAdd(p.stats + 2 * s_b + 1, (double)ss_acc);
                    }
                    s_b = b;
                    s_acc = 0.f;
                    ss_acc = 0.f;
                }
                s_acc += s;
                ss_acc += ss;
            }
        }
    }
    if (want_stats && tx == 0 && s_b >= 0) {
        atomicAdd(p.stats + 2 * s_b, (double)s_acc);
        atomicAdd(p.stats + 2 * s_b + 1, (double)ss_acc);
    }
}

template <int BM, int BN>
int launch_tile(const GemmParams& p, cudaStream_t st) {
    dim3 grid((p.M + BM - 1) / BM, (p.Npad + BN - 1) / BN);
    gemm_fp32_kernel<BM, BN><<<grid, kThreads, 0, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_gemm_fp32(const GemmParams& p, cudaStream_t st) {
    SE_REQUIRE(p.K % 4 == 0, "gemm_fp32: K must be a multiple of 4");
    SE_REQUIRE(p.epi != EPI_GRU, "gemm_fp32: fused GRU epilogue exists only on the tf32 path");
    if (p.M <= 0) return 0;
    if (p.Npad <= 16) return launch_tile<256, 16>(p, st);
    if (p.Npad <= 32) return launch_tile<128, 32>(p, st);
    if (p.Npad <= 64) return launch_tile<64, 64>(p, st);
    // wide N: prefer 64x64 tiles (more CTAs) unless M is large
    return launch_tile<64, 64>(p, st);
}

}  // namespace se
