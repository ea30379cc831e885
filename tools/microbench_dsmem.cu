// microbench_dsmem.cu — throughput of scattered 32-bit atomic adds into the shared memory of OTHER CTAs of a
// thread-block cluster (red.shared::cluster through mapa), next to the same atomics on the CTA's own shared memory.
// Design evidence for the k >= 9 path: a k = 10 row with 16-bit counters (1 MiB) fits the shared
// memory of an 8-CTA cluster, so a cluster-distributed histogram would need no zeroing of global rows, no L2
// waves and no grid barriers — IF remote shared-memory atomics are fast enough.
//   mode 0: every lane -> random word of the CTA's OWN histogram (ATOMS)
//   mode 1: every lane -> random word of a random CTA of the cluster (7/8 remote at cluster size 8)
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t &x) { x = x * 1664525u + 1013904223u; return x >> 4; }

template <int MODE>
__global__ void dsmem_kernel(uint32_t words_mask, int iters, uint32_t *sink) {
    extern __shared__ uint32_t hist[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t csize = cluster.num_blocks();
    for (uint32_t i = threadIdx.x; i <= words_mask; i += blockDim.x) hist[i] = 0;
    cluster.sync();
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(hist);
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
        const uint32_t r = lcg(x);
        const uint32_t w = r & words_mask;
        if (MODE == 0) {
            atomicAdd(hist + w, 1u);
        } else {
            const uint32_t target = (r >> 20) % csize;
            uint32_t remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(base + 4u * w), "r"(target));
            asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" :: "r"(remote), "r"(1u) : "memory");
        }
    }
    cluster.sync();
    uint32_t s = 0;
    for (uint32_t i = threadIdx.x; i <= words_mask; i += blockDim.x) s += hist[i];
    if (s == 0xFFFFFFFFu) sink[0] = s;
}

template <int MODE>
int run(int csize, size_t smem_bytes, int threads, int sms, int iters, int khz, uint32_t *sink) {
    auto kern = dsmem_kernel<MODE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    if (csize > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    const int grid = (sms / csize) * csize;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const uint32_t mask = (uint32_t)(smem_bytes / 4 - 1);
    CK(cudaLaunchKernelEx(&cfg, kern, mask, iters, sink));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, kern, mask, iters, sink));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)grid * threads * iters;
    printf("mode %d cluster %2d smem %3zu KB %4d thr, %3d CTAs: %7.1f G atom/s  %.3f atom/clk/SM  (%.3f ms)\n", MODE, csize,
           smem_bytes >> 10, threads, grid, n / ms / 1e6, n / (ms * 1e-3) / ((double)khz * 1e3) / grid, ms);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, khz = p.clockRate;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, khz);
    uint32_t *sink; CK(cudaMalloc(&sink, 4));
    const int it = 1 << 12;
    for (int thr : {256, 1024}) {
        run<0>(1, 128 << 10, thr, sms, it, khz, sink);
        for (int cs : {2, 4, 8}) run<1>(cs, 128 << 10, thr, sms, it, khz, sink);
    }
    run<1>(8, 64 << 10, 1024, sms, it, khz, sink);
    run<1>(16, 128 << 10, 1024, sms, it, khz, sink);
    return 0;
}
