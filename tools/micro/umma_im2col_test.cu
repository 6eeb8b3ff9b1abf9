// Micro-test (GPU): tcgen05.mma with NO-SWIZZLE K-major descriptors whose 16-byte K chunks OVERLAP between rows, i.e.
// an implicit im2col of an 8-channel fp16 activation row held once in shared memory:
//   A(row p, K chunk j) = X[p + j]  (16 bytes = 8 channels)  via SBO = 128 B (8 rows x 16 B) and LBO = 16 B.
// out[p][n] = sum_{kf<5, c<8} X[p + kf][c] * W[n][kf][c], 128 positions x 16 output columns, checked against the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_im2col_test umma_im2col_test.cu && ./umma_im2col_test
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void k(const __half* X, const __half* W, float* out) {
    __shared__ __align__(128) __half sx[(128 + 8) * 8];   // positions 0..135, 16 B each
    __shared__ __align__(128) __half sw[3 * 2 * 16 * 8];  // [kstep][chunk][n][8]
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 136 * 8; i += blockDim.x) sx[i] = X[i];
    for (int i = tid; i < 3 * 2 * 16 * 8; i += blockDim.x) {
        const int c = i & 7, n = (i >> 3) & 15, j = (i >> 7) & 1, s = i >> 8;
        const int kf = 2 * s + j;
        sw[i] = (kf < 5 && n < 8) ? W[(n * 5 + kf) * 8 + c] : __float2half(0.f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int s = 0; s < 3; ++s) {
            const uint64_t ad = desc_nosw(smem_u32(sx) + 32u * s, 16, 128);        // taps 2s, 2s+1
            const uint64_t bd = desc_nosw(smem_u32(sw) + 512u * s, 256, 128);
            const uint32_t acc = s ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    {
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        uint32_t v[16];
        const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(ta));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) out[tid * 16 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}
int main() {
    const int NX = 136 * 8, NW = 8 * 5 * 8;
    __half hx[NX], hw[NW];
    float fx[NX], fw[NW];
    srand(1);
    for (int i = 0; i < NX; ++i) { fx[i] = (rand() % 2001 - 1000) / 1000.f; hx[i] = __float2half(fx[i]); fx[i] = __half2float(hx[i]); }
    for (int i = 0; i < NW; ++i) { fw[i] = (rand() % 2001 - 1000) / 1000.f; hw[i] = __float2half(fw[i]); fw[i] = __half2float(hw[i]); }
    __half *dx, *dw; float* dout;
    cudaMalloc(&dx, sizeof(hx)); cudaMalloc(&dw, sizeof(hw)); cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(dx, hx, sizeof(hx), cudaMemcpyHostToDevice); cudaMemcpy(dw, hw, sizeof(hw), cudaMemcpyHostToDevice);
    k<<<1, 128>>>(dx, dw, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    float ho[128 * 16];
    cudaMemcpy(ho, dout, sizeof(ho), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxpad = 0;
    for (int p = 0; p < 128; ++p)
        for (int n = 0; n < 16; ++n) {
            double ref = 0;
            if (n < 8) for (int kf = 0; kf < 5; ++kf) for (int c = 0; c < 8; ++c) ref += (double)fx[(p + kf) * 8 + c] * fw[(n * 5 + kf) * 8 + c];
            const double d = fabs(ref - ho[p * 16 + n]);
            if (n < 8) { if (d > maxerr) maxerr = d; } else if (d > maxpad) maxpad = d;
        }
    printf("implicit-im2col UMMA: max |err| = %.3e (real columns), %.3e (zero columns)  -> %s\n", maxerr, maxpad,
           maxerr < 1e-3 && maxpad == 0 ? "OK" : "MISMATCH");
    return maxerr < 1e-3 && maxpad == 0 ? 0 : 2;
}
