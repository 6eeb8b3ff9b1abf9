// Micro-benchmark (GPU): all-to-all exchange inside a thread-block cluster with the bulk-copy engine
// (cp.async.bulk.shared::cluster.shared::cta + mbarrier complete_tx on the destination CTA), the pattern a cluster-resident
// GRU recurrence would use to hand every CTA's slice of h_t to its peers without going through L2 / global memory:
// each CTA of a cluster of CS sends SLICE bytes to each of the CS-1 peers per step and waits until all peer slices arrived.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o dsmem_bulk_bw dsmem_bulk_bw.cu && ./dsmem_bulk_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS, int SLICE>
__global__ void __launch_bounds__(128, 1) k(int steps, long long* cycles, unsigned* checksum) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [2 buffers][CS slices][SLICE] receive area, then the CTA's own outgoing slice, then 2 mbarriers
    unsigned char* rx = smem;
    unsigned char* tx = smem + 2 * CS * SLICE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tx + SLICE);
    const uint32_t me = cluster_rank();
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < SLICE / 4; i += blockDim.x) reinterpret_cast<unsigned*>(tx)[i] = me * 1000003u + i;
    __syncthreads();
    cluster_sync();
    const long long t0 = clock64();
    unsigned acc = 0;
    for (int s = 0; s < steps; ++s) {
        const int buf = s & 1;
        if (tid == 0) {
            // expect the CS-1 peer slices of this step on my barrier, then push my slice to every peer
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + buf)),
                         "r"((unsigned)((CS - 1) * SLICE)) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (uint32_t d = 1; d < CS; ++d) {
                const uint32_t peer = (me + d) % CS;
                const uint32_t dst = mapa(smem_u32(rx + (buf * CS + me) * SLICE), peer);
                const uint32_t bar = mapa(smem_u32(bars + buf), peer);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "r"(smem_u32(tx)), "r"((unsigned)SLICE), "r"(bar) : "memory");
            }
        }
        // everyone waits for the peers' slices
        unsigned ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bars + buf)), "r"((unsigned)((s >> 1) & 1)) : "memory");
        acc += reinterpret_cast<unsigned*>(rx + (buf * CS + (me + 1) % CS) * SLICE)[tid];
        // a buffer is reused two steps later: peers must have consumed it -> one cluster barrier per step keeps it simple
        cluster_sync();
    }
    const long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    atomicAdd(checksum, acc);
}

template <int CS, int SLICE>
void run(int steps) {
    long long* dc; unsigned* ds;
    cudaMalloc(&dc, 8); cudaMalloc(&ds, 4); cudaMemset(ds, 0, 4);
    const size_t smem = (size_t)(2 * CS + 1) * SLICE + 64;
    cudaFuncSetAttribute(k<CS, SLICE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<CS, SLICE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS * (128 / CS));  // 128 CTAs like the GRU recurrence of 1024 streams
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    cudaOccupancyMaxActiveClusters(&ncl, k<CS, SLICE>, &cfg);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<CS, SLICE>, steps, dc, ds);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cluster %2d slice %5d: %s\n", CS, SLICE, cudaGetErrorString(e)); cudaGetLastError(); return; }
    long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / steps;
    printf("cluster %2d, slice %5d B (smem %3zu KB, max active clusters %d): %8.0f cycles per all-to-all step, %.1f B/clk received per CTA\n",
           CS, SLICE, smem / 1024, ncl, per, (double)(CS - 1) * SLICE / per);
    cudaFree(dc); cudaFree(ds);
}

int main() {
    run<2, 8192>(200);
    run<4, 8192>(200);
    run<8, 8192>(200);
    run<16, 8192>(200);
    run<16, 4096>(200);
    run<8, 16384>(200);
    return 0;
}
