// Pre-convolution block of CRN_ELU.py:337-339,375-376 as ONE kernel per layer: 5x5 frequency-dilated causal
// convolution (5 -> 5 channels) + ELU + gated 1x1 pair + GlobalLayerNorm + residual add, with the causal state roll.
//
// The generic implicit-GEMM path would fetch every input value 25 times through L2 for this layer (K = 25 taps x 5
// channels, N = 5): it is bound by L2 bandwidth, not by math.  Here one CTA owns one stream: the whole zero-bordered
// input of the chunk (5 channels x 25 frames x <=220 bins, channel-planar, ~105 KB) is staged in shared memory once,
// (persistent CTAs prefetch the next stream's input with cp.async while computing the current one),
// each thread computes 4 adjacent bins x 5 output channels per pass in registers with fp32 FMAs (exact mode: no TF32
// rounding in this block), the per-stream normalisation statistics are reduced inside the CTA (no atomics, no second
// pass over HBM) and the normalised output + residual is written straight into the next layer's input buffer.
#include <cuda_fp16.h>

#include "se_internal.h"

namespace se {
namespace {

constexpr int T = kFramesPerChunk;  // 21
constexpr int NB = 201;             // bins
constexpr int CH = 5;               // channels (2M-1 for 3 microphones)
constexpr int KT = 5, KF = 5;       // taps
constexpr int TP = T + KT - 1;      // 25 frames: 4 carried + 21 new
constexpr int GPF = (NB + 3) / 4;   // 51 groups of 4 bins per frame
constexpr int NGROUPS = T * GPF;    // 1071
constexpr int kThreads = 384;
constexpr int kPasses = (NGROUPS + kThreads - 1) / kThreads;  // 3

__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int D>
__global__ void __launch_bounds__(kThreads, 1) preconv_kernel(PreconvParams p) {
    constexpr int FPP = PRECONV_FPP(D);
    constexpr int NV4 = 1 + D;  // float4 loads that cover the 4 + 4*D input bins one group needs per (frame, channel)
    extern __shared__ __align__(16) float sm[];
    constexpr int NIN = CH * TP * FPP;      // floats of one stream's input
    float* s_w = sm + 2 * NIN;              // packed weights (PRECONV_W_FLOATS) behind the two input buffers
    double* s_red = reinterpret_cast<double*>(s_w + PRECONV_W_FLOATS);  // [2][warps]
    const int tid = threadIdx.x;
    for (int i = tid; i < PRECONV_W_FLOATS; i += kThreads) s_w[i] = __ldg(p.w + i);

    // Persistent CTA: streams blockIdx.x, blockIdx.x + gridDim.x, ...  The input of the NEXT stream is fetched with
    // cp.async into the other shared-memory buffer while this one is computed (the load phase was 19 % of the kernel).
    auto prefetch = [&](int stream, float* dst) {
        const float* src = p.in + (long long)(p.b0 + stream) * p.in_sB;
        const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
        for (int i = tid; i < NIN / 4; i += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + 16u * i), "l"(src + 4 * i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cur = 0;
    if ((int)blockIdx.x < p.B) prefetch(blockIdx.x, sm);
  for (int stream = blockIdx.x; stream < p.B; stream += gridDim.x, cur ^= 1) {
    float* s_in = sm + cur * NIN;           // [CH][TP][FPP]
    const int b = p.b0 + stream;
    float* gin = p.in + (long long)b * p.in_sB;
    if (stream + (int)gridDim.x < p.B) {
        prefetch(stream + gridDim.x, sm + (cur ^ 1) * NIN);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // ---- causal state: the last 4 frames of this chunk's input become frames 0..3 of the next chunk (CRN_ELU.py:246)
    for (int i = tid; i < CH * (KT - 1) * FPP / 4; i += kThreads) {
        const int c = i / ((KT - 1) * FPP / 4);
        const int r = i - c * ((KT - 1) * FPP / 4);
        reinterpret_cast<float4*>(gin + (long long)c * TP * FPP)[r] =
            reinterpret_cast<const float4*>(s_in + (c * TP + T) * FPP)[r];
    }

    // ---- convolution + ELU + gate, 4 bins x 5 channels per thread and pass ----------------------------------------
    float yv[kPasses][4][CH];
    float psum = 0.f, psq = 0.f;
#pragma unroll
    for (int pass = 0; pass < kPasses; ++pass) {
        const int gi = tid + pass * kThreads;
        if (gi < NGROUPS) {
            const int t = gi / GPF;
            const int f0 = 4 * (gi - t * GPF);
            float acc[4][CH];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int co = 0; co < CH; ++co) acc[q][co] = s_w[PRECONV_W_BIAS + co];
#pragma unroll 1
            for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
                for (int ci = 0; ci < CH; ++ci) {
                    float x[4 * NV4];
                    const float4* row = reinterpret_cast<const float4*>(s_in + (ci * TP + t + kt) * FPP + f0);
#pragma unroll
                    for (int v = 0; v < NV4; ++v) {
                        const float4 r = row[v];
                        x[4 * v] = r.x;
                        x[4 * v + 1] = r.y;
                        x[4 * v + 2] = r.z;
                        x[4 * v + 3] = r.w;
                    }
                    float w[28];
                    const float4* wr = reinterpret_cast<const float4*>(s_w + (kt * CH + ci) * 28);
#pragma unroll
                    for (int v = 0; v < 7; ++v) {
                        const float4 r = wr[v];
                        w[4 * v] = r.x;
                        w[4 * v + 1] = r.y;
                        w[4 * v + 2] = r.z;
                        w[4 * v + 3] = r.w;
                    }
#pragma unroll
                    for (int kf = 0; kf < KF; ++kf)
#pragma unroll
                        for (int q = 0; q < 4; ++q)
#pragma unroll
                            for (int co = 0; co < CH; ++co)
                                acc[q][co] = fmaf(x[q + kf * D], w[kf * CH + co], acc[q][co]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float e[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) e[c] = elu1(acc[q][c]);
                const bool valid = f0 + q < NB;
#pragma unroll
                for (int co = 0; co < CH; ++co) {
                    float a = s_w[PRECONV_W_BT + co], g = s_w[PRECONV_W_BG + co];
#pragma unroll
                    for (int k = 0; k < CH; ++k) {
                        a = fmaf(s_w[PRECONV_W_WT + co * CH + k], e[k], a);
                        g = fmaf(s_w[PRECONV_W_WG + co * CH + k], e[k], g);
                    }
                    const float y = a * sigmoidf_(g);
                    yv[pass][q][co] = y;
                    if (valid) {
                        psum += y;
                        psq += y * y;
                    }
                }
            }
        }
    }

    // ---- GlobalLayerNorm statistics of this stream (CRN_ELU.py:40-41), reduced in double ----------------------------
    double ds = psum, dq = psq;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, off);
        dq += __shfl_xor_sync(0xffffffffu, dq, off);
    }
    constexpr int NW = kThreads / 32;
    if ((tid & 31) == 0) {
        s_red[tid >> 5] = ds;
        s_red[NW + (tid >> 5)] = dq;
    }
    __syncthreads();
    ds = 0.0;
    dq = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        ds += s_red[w];
        dq += s_red[NW + w];
    }
    const double count = (double)CH * NB * T;
    const double mu = ds / count;
    double var = dq / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const float varf = (float)var;
    const float den = p.student ? (sqrtf(varf) + 1e-8f) : (sqrtf(varf + 1e-8f) + 1e-8f);  // distillation_crn.py:51 / CRN_ELU.py:51
    const float mean = (float)mu;
    const float inv = 1.0f / den;

    // ---- normalise, add the block input (CRN_ELU.py:376) and write the next layer's input ---------------------------
    float nw[CH], nb[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        nw[c] = s_w[PRECONV_W_NW + c];
        nb[c] = s_w[PRECONV_W_NB + c];
    }
    float* gout = p.out + (long long)b * p.oB;
#pragma unroll
    for (int pass = 0; pass < kPasses; ++pass) {
        const int gi = tid + pass * kThreads;
        if (gi < NGROUPS) {
            const int t = gi / GPF;
            const int f0 = 4 * (gi - t * GPF);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int f = f0 + q;
                if (f < NB) {
                    float o[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c)
                        o[c] = (yv[pass][q][c] - mean) * inv * nw[c] + nb[c] +
                               s_in[(c * TP + t + KT - 1) * FPP + f + 2 * D];
                    float* dst = gout + t * p.oT + f * p.oF;
                    if (p.out_vec8 && p.out_half) {  // ... stored as fp16: 8 halves = one 16-byte store
                        __half* dh = reinterpret_cast<__half*>(p.out) + (long long)b * p.oB + t * p.oT + f * p.oF;
                        const __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]),
                                      h2 = __floats2half2_rn(o[4], 0.f);
                        uint4 u;
                        u.x = *reinterpret_cast<const unsigned*>(&h0);
                        u.y = *reinterpret_cast<const unsigned*>(&h1);
                        u.z = *reinterpret_cast<const unsigned*>(&h2);
                        u.w = 0u;
                        *reinterpret_cast<uint4*>(dh) = u;
                    } else if (p.out_vec8) {  // channels-last C = 8 destination (first encoder input)
                        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], 0.f, 0.f, 0.f);
                    } else {
#pragma unroll
                        for (int c = 0; c < CH; ++c) dst[c * p.oC] = o[c];
                    }
                }
            }
        }
    }
    __syncthreads();  // everyone is done with s_in before it is refilled two iterations later
  }
}

template <int D>
int launch_d(const PreconvParams& p, cudaStream_t st) {
    constexpr size_t bytes = (size_t)(2 * CH * TP * PRECONV_FPP(D) + PRECONV_W_FLOATS) * sizeof(float) +
                             2 * (kThreads / 32) * sizeof(double);
    static_assert(bytes <= 227 * 1024, "two input buffers must fit in shared memory");
    SE_DYN_SMEM(preconv_kernel<D>, bytes);
    int num_sms = 0;
    if (num_sms_current_device(&num_sms)) return 1;
    preconv_kernel<D><<<p.B < num_sms ? p.B : num_sms, kThreads, bytes, st>>>(p);
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_preconv(const PreconvParams& p, cudaStream_t st) {
    if (p.B <= 0) return 0;
    switch (p.d) {
        case 1: return launch_d<1>(p, st);
        case 2: return launch_d<2>(p, st);
        case 4: return launch_d<4>(p, st);
    }
    SE_REQUIRE(false, "preconv: frequency dilation must be 1, 2 or 4 (CRN_ELU.py:336)");
    return 2;
}

}  // namespace se
