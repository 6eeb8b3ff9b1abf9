// Tensor-core gathered GEMM for sm_100a: tcgen05.mma kind::tf32 (fp32 operands read as TF32, fp32 accumulation in
// TMEM).  Same GemmParams contract as gemm_fp32.cu -- C[m][n] = sum_k A(m,k) W[n][k] with A gathered through the koff
// table (implicit im2col over zero-bordered channels-last activations) -- plus the fused GRU cell epilogue.
//
// One CTA computes a 128 x BN tile:
//   warps 0-3 : producers.  Thread r owns tile row r: per 32-float k-block it issues 8 x 16-byte cp.async gathers into
//               the canonical K-major SWIZZLE_128B layout (row r at (r/8)*1024 + (r%8)*128, 16-byte chunk j stored at
//               j ^ (r%8)) and the same for its share of the weight rows, then arrives on the stage's "full" mbarrier
//               through cp.async.mbarrier.arrive.noinc.  After the k loop the same threads run the epilogue: thread r
//               reads accumulator row r from TMEM (tcgen05.ld 32x32b) and writes its N outputs channels-last.
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

// accumulate (s, ss) of one tile row group into the per-stream statistics with as few atomics as possible
__device__ __forceinline__ void stats_commit(double* stats, int b, float s, float ss) {
    const unsigned full = 0xffffffffu;
    const int b0 = __shfl_sync(full, b, 0);
    const bool uniform = __all_sync(full, b == b0);
    if (uniform) {
        if (b0 < 0) return;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            s += __shfl_xor_sync(full, s, off);
            ss += __shfl_xor_sync(full, ss, off);
        }
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(stats + 2 * b0, (double)s);
            atomicAdd(stats + 2 * b0 + 1, (double)ss);
        }
    } else if (b >= 0) {
        atomicAdd(stats + 2 * b, (double)s);
        atomicAdd(stats + 2 * b + 1, (double)ss);
    }
}

template <int BN, int STAGES>
struct TcSmem {
    static constexpr int B_STAGE_BYTES = BN * BK * 4;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int KOFF_MAX = 512;  // K <= 2048
    static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + KOFF_MAX * 4 + 256;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads) gemm_tf32_kernel(GemmParams p) {
    using S = TcSmem<BN, STAGES>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* tiles_ptr = smem_raw + (tiles - raw);
    int* s_koff = reinterpret_cast<int*>(tiles_ptr + STAGES * S::STAGE_BYTES);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(tiles_ptr + STAGES * S::STAGE_BYTES + S::KOFF_MAX * 4);
    // barriers: [0,STAGES) full, [STAGES,2*STAGES) empty, [2*STAGES] accumulator; then the TMEM base address
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * STAGES + 1);
    const uint32_t bar0 = smem_u32(s_bar);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t acc_bar = bar0 + 8u * (2 * STAGES);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int nkb = (p.K + BK - 1) / BK;
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

    for (int i = tid; i < p.K / 4; i += kThreads) s_koff[i] = __ldg(p.koff + i);
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
        const int r = tid;  // tile row
        const int m = m0 + r;
        const int rowsPerStream = p.Tn * p.Fo;
        int b = -1, t = 0, f = 0;
        const float* a_row = p.A;
        uint32_t a_ok = 0;
        if (m < p.M) {
            b = m / rowsPerStream;
            const int rr = m - b * rowsPerStream;
            t = rr / p.Fo;
            f = rr - t * p.Fo;
            a_row = p.A + b * p.sB + t * p.sT + f * p.sF;
            a_ok = 16;
        }
        const uint32_t sw = (uint32_t)(r & 7);
        const uint32_t a_dst_row = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        constexpr int B_ROWS_PER_THREAD = (BN + kProducerThreads - 1) / kProducerThreads;

        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % STAGES;
            mbar_wait(empty_bar(s), ((kb / STAGES) & 1) ^ 1);
            const uint32_t stage = tiles + (uint32_t)s * S::STAGE_BYTES;
            const int k0 = kb * BK;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + 4 * j;
                const bool kin = k < p.K;
                const float* src = kin ? a_row + s_koff[k >> 2] : p.A;
                cp_async16(stage + a_dst_row + ((j ^ sw) << 4), src, kin ? a_ok : 0u);
            }
#pragma unroll
            for (int i = 0; i < B_ROWS_PER_THREAD; ++i) {
                const int nl = r + i * kProducerThreads;  // row inside the weight tile
                if (nl < BN) {
                    const int n = n0 + nl;
                    const bool nin = n < p.Npad;
                    const float* wrow = p.W + (long long)(nin ? n : 0) * p.K;
                    const uint32_t dst_row =
                        stage + A_STAGE_BYTES + (uint32_t)((nl >> 3) * 1024 + (nl & 7) * 128);
                    const uint32_t swb = (uint32_t)(nl & 7);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = k0 + 4 * j;
                        const bool ok = nin && k < p.K;
                        cp_async16(dst_row + ((j ^ swb) << 4), ok ? wrow + k : p.W, ok ? 16u : 0u);
                    }
                }
            }
            cp_async_mbar_arrive_noinc(full_bar(s));
        }

        // ============================ epilogue ============================
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool row_ok = m < p.M;
        float s_acc = 0.f, ss_acc = 0.f;

        if (p.epi == EPI_GRU) {
            // tile columns: [r | z | n] of hidden units j0 .. j0+BN/3
            constexpr int U = BN / 3;
            const int j0 = blockIdx.y * U;
            const float* gi = p.gi + (row_ok ? (long long)m * p.giB : 0);
            const float* hp = p.hprev + (row_ok ? (long long)m * p.hB : 0);
            float* ho = p.out + (row_ok ? (long long)m * p.oB : 0);
            const float* bias = p.bias + (long long)blockIdx.y * BN;
#pragma unroll 1
            for (int u0 = 0; u0 < U; u0 += 8) {
                float ar[8], az[8], an[8];
                tmem_ld8(trow + u0, ar);
                tmem_ld8(trow + U + u0, az);
                tmem_ld8(trow + 2 * U + u0, an);
                if (row_ok) {
                    float hn[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int j = j0 + u0 + i;
                        const float rg = sigmoidf_(gi[j] + ar[i] + __ldg(bias + u0 + i));
                        const float zg = sigmoidf_(gi[p.H + j] + az[i] + __ldg(bias + U + u0 + i));
                        const float ng = tanhf(gi[2 * p.H + j] + rg * (an[i] + __ldg(bias + 2 * U + u0 + i)));
                        hn[i] = (1.0f - zg) * ng + zg * hp[j];
                    }
                    *reinterpret_cast<float4*>(ho + j0 + u0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    *reinterpret_cast<float4*>(ho + j0 + u0 + 4) = make_float4(hn[4], hn[5], hn[6], hn[7]);
                }
            }
        } else {
            const bool paired = (p.epi == EPI_GATE_STATS || p.epi == EPI_SKIP);
            const bool want_stats = (p.epi == EPI_ELU_STATS || p.epi == EPI_GATE_STATS || p.epi == EPI_SKIP);
            float* o = row_ok ? p.out + b * p.oB + t * p.oT + f * p.oF : p.out;
            float* o2 = (p.epi == EPI_SKIP && row_ok) ? p.out2 + b * p.o2B + t * p.o2T + f * p.o2F : p.out2;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                float v[16];
                tmem_ld16(trow + c0, v);
                const int n = n0 + c0;
                if (row_ok && n < p.N) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(p.bias + n + i);
                if (!paired) {
                    if (p.epi != EPI_BIAS) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = elu1(v[i]);
                    }
                    if (n + 16 <= p.N) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(o + n + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            s_acc += v[i];
                            ss_acc += v[i] * v[i];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (n + i < p.N) {
                                o[n + i] = v[i];
                                s_acc += v[i];
                                ss_acc += v[i] * v[i];
                            }
                    }
                } else {
                    float w[8];
                    const int c = n >> 1;
                    const int nc = (p.N - n) >> 1;  // valid output channels in this group (>= 1)
                    if (p.epi == EPI_GATE_STATS) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = v[2 * i] * sigmoidf_(v[2 * i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = v[2 * i];
                    }
                    if (nc >= 8) {
                        *reinterpret_cast<float4*>(o + c) = make_float4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<float4*>(o + c + 4) = make_float4(w[4], w[5], w[6], w[7]);
                        if (p.epi == EPI_SKIP) {
                            *reinterpret_cast<float4*>(o2 + c) =
                                make_float4(elu1(v[1]), elu1(v[3]), elu1(v[5]), elu1(v[7]));
                            *reinterpret_cast<float4*>(o2 + c + 4) =
                                make_float4(elu1(v[9]), elu1(v[11]), elu1(v[13]), elu1(v[15]));
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            s_acc += w[i];
                            ss_acc += w[i] * w[i];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (i < nc) {
                                o[c + i] = w[i];
                                if (p.epi == EPI_SKIP) o2[c + i] = elu1(v[2 * i + 1]);
                                s_acc += w[i];
                                ss_acc += w[i] * w[i];
                            }
                    }
                }
                }  // row_ok && n < N
            }
            if (want_stats) stats_commit(p.stats, row_ok ? b : -1, s_acc, ss_acc);
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
    if (p.epi == EPI_GRU) return p.H % 32 == 0 && p.K % 4 == 0 && p.K <= 2048;
    // tiny contractions stay on the CUDA-core kernel: tile set-up would dominate
    return p.K % 4 == 0 && p.K >= 32 && p.K <= 2048 && p.N >= 8 && (p.oF % 4 == 0) && (p.oT % 4 == 0) &&
           (p.oB % 4 == 0);
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
