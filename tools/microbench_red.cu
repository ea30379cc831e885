// microbench_red.cu — chip-wide throughput of scattered global RED.ADD.U32 on sm_100a (design evidence for the
// k >= 9 path in DESIGN.md).  Every thread issues `iters` REDs to pseudo-random words of a region of `bytes`
// (L2-resident when small); prints G RED/s and REDs per clock per SM.
//   pattern 0: 32 lanes -> 32 random words anywhere in the region
//   pattern 1: 32 lanes -> 32 random words, but 2 consecutive REDs of a lane hit the same 32-byte sector
//   pattern 2: coalesced (lane i -> word base+i), random base per instruction
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t lcg(uint32_t &x) { x = x * 1664525u + 1013904223u; return x >> 4; }

template <int PATTERN>
__global__ void red_kernel(uint32_t *region, uint32_t words_mask, int iters) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const int lane = threadIdx.x & 31;
    uint32_t prev = 0;
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
        uint32_t w = lcg(x) & words_mask;
        if (PATTERN == 1) { if (it & 1) w = prev ^ 1u; prev = w; }
        if (PATTERN == 2) w = ((__shfl_sync(0xffffffffu, w, 0) & ~31u) + lane) & words_mask;
        atomicAdd(region + w, 1u);
    }
}

template <int PATTERN>
int run(size_t bytes, int threads, int ctas_per_sm, int sms, int iters, int khz) {
    uint32_t *region;
    CK(cudaMalloc(&region, bytes));
    CK(cudaMemset(region, 0, bytes));
    const uint32_t mask = (uint32_t)(bytes / 4 - 1);
    red_kernel<PATTERN><<<sms * ctas_per_sm, threads>>>(region, mask, iters);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    red_kernel<PATTERN><<<sms * ctas_per_sm, threads>>>(region, mask, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)sms * ctas_per_sm * threads * iters;
    printf("pattern %d region %6zu MB  %4d thr x %d CTA/SM : %7.1f G RED/s  %.3f RED/clk/SM  (%.3f ms)\n", PATTERN,
           bytes >> 20, threads, ctas_per_sm, n / ms / 1e6, n / (ms * 1e-3) / ((double)khz * 1e3) / sms, ms);
    cudaFree(region);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount, khz = p.clockRate;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, khz);
    const int it = 1 << 12;
    for (size_t mb : {2, 16, 32, 64, 128, 512})
        for (int thr : {256, 1024}) run<0>(mb << 20, thr, 1, sms, it, khz);
    run<0>(16 << 20, 1024, 2, sms, it, khz);
    run<0>(16 << 20, 128, 1, sms, it, khz);
    run<0>(16 << 20, 32, 1, sms, it * 4, khz);
    for (size_t mb : {16, 64}) { run<1>(mb << 20, 1024, 1, sms, it, khz); run<2>(mb << 20, 1024, 1, sms, it, khz); }
    return 0;
}
