// Direct CUDA-core kernels for the two layer shapes that the 128-row tensor-core tiles serve badly (fp16 operand mode):
//   * the last transposed convolution (CRN_ELU.py:352-358, 16 -> 2 channels): the GEMM has N = 4 useful columns of a
//     16-wide tile and gathers 9 taps x 16 channels per row -- 0.6 % of the tensor peak, 194 us per step;
//   * the 1x1 mask / residual pair on a 16- (or 8-) channel skip tensor (CRN_ELU.py:305-306): K = 16 is padded to a
//     64-element k-block, so 3/4 of the operand traffic and of the MMA work is zeros -- 151 us per step.
// Both are a few hundred FMAs per output position on data that is read once: one thread per position, weights
// broadcast from shared memory, per-stream GlobalLayerNorm statistics reduced per block.
#include <cuda_fp16.h>

#include "se_internal.h"

namespace se {
namespace {

constexpr int T = kFramesPerChunk;

__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : __expf(x) - 1.0f; }

template <int N>
__device__ __forceinline__ void load_halves(const __half* p, float* v) {  // N = 8 or 16 halves, 16-byte aligned
#pragma unroll
    for (int u = 0; u < N / 8; ++u) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(p) + u);
        const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h[i]);
            v[8 * u + 2 * i] = f.x;
            v[8 * u + 2 * i + 1] = f.y;
        }
    }
}

__device__ __forceinline__ void block_stats(float s, float ss, double* stats, int b) {
    __shared__ float red[2][8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, off);
        ss += __shfl_xor_sync(0xffffffffu, ss, off);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s;
        red[1][threadIdx.x >> 5] = ss;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            a += red[0][w];
            c += red[1][w];
        }
        atomicAdd(stats + 2 * b, (double)a);
        atomicAdd(stats + 2 * b + 1, (double)c);
    }
}

// grid (ceil(Fin / 32), B), 256 threads: lane = input bin f' inside the block's 32-bin slab, warp w = frames w, w+8, w+16
template <int CIN>
__global__ void __launch_bounds__(256) deconv_last_kernel(DeconvLastParams p) {
    __shared__ __align__(16) float sw[9 * CIN * 4];  // [kt][j][ci][n], n = parity * 2 + co
    __shared__ float sb[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 9 * CIN * 4; i += 256) {
        const int n = i & 3, k = i >> 2;  // k = (kt * 3 + j) * CIN + ci
        sw[i] = __ldg(p.w + (long long)n * p.Kp + k);
    }
    if (tid < 4) sb[tid] = __ldg(p.bias + tid);
    __syncthreads();
    const int b = blockIdx.y;
    const int f = blockIdx.x * 32 + lane;
    const bool valid = f < p.Fin;
    const int fl = valid ? f : p.Fin - 1;  // clamped: out-of-range lanes compute on valid memory and store nothing
    const __half* base = p.in + (long long)b * p.sB + (long long)fl * p.sF;
    const int Fy = 2 * p.Fin - 1;
    float s = 0.f, ss = 0.f;
    // the warp's (up to) three frames are processed together: one weight fetch feeds 12 FMAs and three independent
    // load streams are in flight
    int tt[3];
    bool tv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        tv[i] = warp + 8 * i < T;
        tt[i] = tv[i] ? warp + 8 * i : warp;  // clamped: computed, not stored
    }
    float acc[3][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[i][n] = sb[n];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float x[3][CIN];
#pragma unroll
            for (int i = 0; i < 3; ++i)
                load_halves<CIN>(base + (long long)(tt[i] + (2 - kt) * p.d) * p.sT + (long long)(2 - j) * p.sF, x[i]);
            const float4* w = reinterpret_cast<const float4*>(sw + (kt * 3 + j) * CIN * 4);
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) {
                const float4 q = w[ci];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    acc[i][0] = fmaf(x[i][ci], q.x, acc[i][0]);
                    acc[i][1] = fmaf(x[i][ci], q.y, acc[i][1]);
                    acc[i][2] = fmaf(x[i][ci], q.z, acc[i][2]);
                    acc[i][3] = fmaf(x[i][ci], q.w, acc[i][3]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (valid && tv[i]) {
            float2* y = reinterpret_cast<float2*>(p.y) + ((long long)b * T + tt[i]) * Fy + 2 * f;
            const float e0 = elu_fast(acc[i][0]), e1 = elu_fast(acc[i][1]);
            y[0] = make_float2(e0, e1);
            s += e0 + e1;
            ss += e0 * e0 + e1 * e1;
            if (f < p.Fin - 1) {  // bin 2 f' + 1 exists
                const float o0 = elu_fast(acc[i][2]), o1 = elu_fast(acc[i][3]);
                y[1] = make_float2(o0, o1);
                s += o0 + o1;
                ss += o0 * o0 + o1 * o1;
            }
        }
    }
    block_stats(s, ss, p.stats, b);
}

// grid (ceil(T * Fs / 256), B), one thread per position
template <int C>
__global__ void __launch_bounds__(256) skip_small_kernel(SkipSmallParams p) {
    __shared__ __align__(16) float sw[2 * C * C];  // [n][ci], n = 2 co (mask) / 2 co + 1 (residual)
    __shared__ float sb[2 * C];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * C * C; i += 256) sw[i] = __ldg(p.w + (long long)(i / C) * p.Kp + i % C);
    if (tid < 2 * C) sb[tid] = __ldg(p.bias + tid);
    __syncthreads();
    const int b = blockIdx.y;
    const int r = blockIdx.x * 256 + tid;
    float s = 0.f, ss = 0.f;
    if (r < T * p.Fs) {
        const int t = r / p.Fs, f = r - t * p.Fs;
        float x[C];
        load_halves<C>(p.in + (long long)b * p.sB + (long long)t * p.sT + (long long)f * p.sF, x);
        uint32_t pm[C / 2], pr[C / 2];  // packed half2 pairs (kept in registers: no local-memory staging)
#pragma unroll
        for (int cp = 0; cp < C / 2; ++cp) {
            float a[4] = {sb[4 * cp], sb[4 * cp + 1], sb[4 * cp + 2], sb[4 * cp + 3]};  // mask, residual of 2 channels
            const float4* w0 = reinterpret_cast<const float4*>(sw + (4 * cp) * C);      // 4 consecutive weight rows
#pragma unroll
            for (int c4 = 0; c4 < C / 4; ++c4) {
#pragma unroll
                for (int r4 = 0; r4 < 4; ++r4) {
                    const float4 q = w0[r4 * (C / 4) + c4];
                    a[r4] = fmaf(q.x, x[4 * c4], a[r4]);
                    a[r4] = fmaf(q.y, x[4 * c4 + 1], a[r4]);
                    a[r4] = fmaf(q.z, x[4 * c4 + 2], a[r4]);
                    a[r4] = fmaf(q.w, x[4 * c4 + 3], a[r4]);
                }
            }
            const __half2 hm = __floats2half2_rn(a[0], a[2]);
            const __half2 hr = __floats2half2_rn(elu_fast(a[1]), elu_fast(a[3]));
            pm[cp] = *reinterpret_cast<const uint32_t*>(&hm);
            pr[cp] = *reinterpret_cast<const uint32_t*>(&hr);
            s += a[0] + a[2];
            ss += a[0] * a[0] + a[2] * a[2];
        }
        const long long o = (((long long)b * T + t) * p.Fs + f) * C;
#pragma unroll
        for (int u = 0; u < C / 8; ++u) {
            reinterpret_cast<uint4*>(p.rm + o)[u] = make_uint4(pm[4 * u], pm[4 * u + 1], pm[4 * u + 2], pm[4 * u + 3]);
            reinterpret_cast<uint4*>(p.rr + o)[u] = make_uint4(pr[4 * u], pr[4 * u + 1], pr[4 * u + 2], pr[4 * u + 3]);
        }
    }
    block_stats(s, ss, p.stats, b);
}

}  // namespace

bool deconv_last_supported(int Cin) { return Cin == 8 || Cin == 16; }
bool skip_small_supported(int C) { return C == 8 || C == 16; }

int launch_deconv_last(const DeconvLastParams& p, int Cin, cudaStream_t st) {
    if (p.B <= 0) return 0;
    SE_REQUIRE(p.B <= 65535, "deconv_last: at most 65535 streams per launch");
    const dim3 grid((p.Fin + 31) / 32, p.B);
    if (Cin == 16) deconv_last_kernel<16><<<grid, 256, 0, st>>>(p);
    else if (Cin == 8) deconv_last_kernel<8><<<grid, 256, 0, st>>>(p);
    else SE_REQUIRE(false, "deconv_last: input channels must be 8 or 16");
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_skip_small(const SkipSmallParams& p, int C, cudaStream_t st) {
    if (p.B <= 0) return 0;
    SE_REQUIRE(p.B <= 65535, "skip_small: at most 65535 streams per launch");
    const dim3 grid((T * p.Fs + 255) / 256, p.B);
    if (C == 16) skip_small_kernel<16><<<grid, 256, 0, st>>>(p);
    else if (C == 8) skip_small_kernel<8><<<grid, 256, 0, st>>>(p);
    else SE_REQUIRE(false, "skip_small: channels must be 8 or 16");
    SE_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace se
