// microbench.cu — shared-memory histogram primitives on sm_100a (design evidence for DESIGN.md).
// Prints lane-updates per cycle per SM for: random-address ATOMS (u32), conflict-free byte RMW
// (LDS.U8/IADD/STS.U8), conflict-free packed ATOMS, for several histogram sizes and occupancies.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lcg(uint32_t &x) { x = x * 1664525u + 1013904223u; return x >> 8; }

// mode 0: ATOMS random over a CTA-shared histogram of nb u32 bins
// mode 1: byte RMW, lane-private bank (word w of lane t at w*32+t), nb bins per lane
// mode 2: packed ATOMS (1<<8*(b&3)) on the same lane-private layout
// mode 3: ATOMS random over a WARP-private histogram of nb bins
template <int MODE>
__global__ void k(uint32_t nb, int iters, unsigned long long *cycles, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t words = (MODE == 0) ? nb : (MODE == 3 ? nb * (blockDim.x / 32) : (nb / 4) * 32 * (blockDim.x / 32));
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    uint8_t *hb = reinterpret_cast<uint8_t *>(sm) + (size_t)warp * (nb / 4) * 128 + lane * 4;
    uint32_t *wh = sm + (size_t)warp * nb;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        const uint32_t b = lcg(x) & (nb - 1);
        if (MODE == 0) atomicAdd(&sm[b], 1u);
        else if (MODE == 3) atomicAdd(&wh[b], 1u);
        else if (MODE == 1) { const uint32_t off = (b >> 2) * 128 + (b & 3); hb[off] = (uint8_t)(hb[off] + 1); }
        else { const uint32_t off = (b >> 2) * 128; atomicAdd(reinterpret_cast<uint32_t *>(hb + off), 1u << ((b & 3) * 8)); }
    }
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) acc += sm[i];
    if (acc == 0xFFFFFFFFu) sink[0] = acc;
}

template <int MODE>
int run(const char *name, uint32_t nb, int threads, int iters, int sms) {
    size_t words = (MODE == 0) ? nb : (MODE == 3 ? (size_t)nb * (threads / 32) : (size_t)(nb / 4) * 32 * (threads / 32));
    size_t smem = words * 4;
    if (smem > 227 * 1024) return 0;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<MODE>, threads, smem));
    if (per_sm < 1) return 0;
    unsigned long long *cyc; uint32_t *sink;
    CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(cyc, 0, 8));
    k<MODE><<<sms * per_sm, threads, smem>>>(nb, iters, cyc, sink);   // warm
    CK(cudaMemset(cyc, 0, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms * per_sm, threads, smem>>>(nb, iters, cyc, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
    double upd_per_sm = (double)per_sm * threads * iters;
    printf("%-34s nb=%6u thr=%4d cta/sm=%2d smem=%7zu  %8.3f upd/cyc/SM  (%.3f ms, %.1f Gupd/s chip)\n", name, nb, threads,
           per_sm, smem, upd_per_sm / (double)c, ms, upd_per_sm * sms / ms / 1e6);
    cudaFree(cyc); cudaFree(sink);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    const int it = 1 << 14;
    for (uint32_t nb : {128u, 512u, 2048u, 8192u, 16384u, 32768u})
        for (int thr : {128, 256, 512, 1024}) run<0>("ATOMS random, CTA-shared u32", nb, thr, it, sms);
    for (uint32_t nb : {128u, 512u, 2048u})
        for (int thr : {32, 64, 128, 256}) run<3>("ATOMS random, warp-private u32", nb, thr, it, sms);
    for (uint32_t nb : {128u, 512u, 1024u})
        for (int thr : {128, 256, 320, 448}) {
            run<1>("byte RMW lane-private bank", nb, thr, it, sms);
            run<2>("packed ATOMS lane-private bank", nb, thr, it, sms);
        }
    return 0;
}
