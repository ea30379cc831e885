// microbench_d2h.cu — page-locked device->host copy bandwidth of 1 / 2 / 4 / 8 GPUs copying at the same time, with
// host buffers (a) wherever cudaHostAlloc puts them and (b) bound to the NUMA node of the GPU (mmap + mbind +
// cudaHostRegister, node from /sys/bus/pci/devices/<bdf>/numa_node).  The end-to-end figure of the headline config is
// bound by exactly this (20.5 GB of rows per GPU and step), so this is its ceiling.
//   nvcc -O3 -o tools/_build/microbench_d2h tools/microbench_d2h.cu && tools/_build/microbench_d2h [GiB per copy]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

static int numa_node_of(int dev) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev) != cudaSuccess) return -1;
    for (char *c = bdf; *c; ++c) *c = (char)tolower(*c);
    FILE *f = fopen((std::string("/sys/bus/pci/devices/") + bdf + "/numa_node").c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static void *alloc_near(size_t bytes, int node, bool *bound) {
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    *bound = false;
    if (node >= 0) {
        unsigned long mask[16] = {0};
        mask[node / 64] |= 1ul << (node % 64);
        *bound = syscall(SYS_mbind, p, bytes, 1 /* MPOL_PREFERRED */, mask, 1024, 0) == 0;
    }
    const unsigned nt = 8;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=] { for (size_t o = bytes / nt * t; o < bytes / nt * (t + 1); o += 4096) ((volatile char *)p)[o] = 0; });
    for (auto &x : th) x.join();
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { munmap(p, bytes); return nullptr; }
    return p;
}

int main(int argc, char **argv) {
    const size_t bytes = (size_t)((argc > 1 ? atof(argv[1]) : 2.0) * (1ull << 30));
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    printf("%d device(s), %.1f GiB per copy, 3 copies per GPU, host CPUs visible: %ld\n", ndev, bytes / 1073741824.0,
           sysconf(_SC_NPROCESSORS_ONLN));
    for (int d = 0; d < ndev; ++d) printf("  gpu %d: numa node %d\n", d, numa_node_of(d));
    for (int mode = 0; mode < 2; ++mode) {
        for (int G = 1; G <= ndev; G *= 2) {
            std::vector<void *> hbuf(G), dbuf(G);
            std::vector<cudaStream_t> st(G);
            std::vector<double> gbs(G);
            bool all_bound = true;
            for (int d = 0; d < G; ++d) {
                CK(cudaSetDevice(d));
                CK(cudaMalloc(&dbuf[d], bytes));
                CK(cudaMemset(dbuf[d], 1, bytes));
                CK(cudaStreamCreate(&st[d]));
                if (mode == 0) { CK(cudaHostAlloc(&hbuf[d], bytes, cudaHostAllocPortable)); }
                else { bool b = false; hbuf[d] = alloc_near(bytes, numa_node_of(d), &b); all_bound &= b; if (!hbuf[d]) { printf("alloc_near failed\n"); return 1; } }
            }
            for (int d = 0; d < G; ++d) { CK(cudaSetDevice(d)); CK(cudaMemcpyAsync(hbuf[d], dbuf[d], bytes, cudaMemcpyDeviceToHost, st[d])); CK(cudaStreamSynchronize(st[d])); }
            std::vector<std::thread> th;
            const auto t0 = std::chrono::steady_clock::now();
            for (int d = 0; d < G; ++d)
                th.emplace_back([&, d] {
                    cudaSetDevice(d);
                    const auto a = std::chrono::steady_clock::now();
                    for (int r = 0; r < 3; ++r) cudaMemcpyAsync(hbuf[d], dbuf[d], bytes, cudaMemcpyDeviceToHost, st[d]);
                    cudaStreamSynchronize(st[d]);
                    gbs[d] = 3.0 * bytes / std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count() / 1e9;
                });
            for (auto &x : th) x.join();
            const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double mn = 1e30, mx = 0;
            for (double g : gbs) { mn = g < mn ? g : mn; mx = g > mx ? g : mx; }
            printf("%-34s %d GPU(s): per GPU %.1f - %.1f GB/s, aggregate %.1f GB/s%s\n",
                   mode == 0 ? "cudaHostAlloc (first-touch node)" : "mmap + mbind(GPU node) + register", G, mn, mx,
                   3.0 * bytes * G / wall / 1e9, mode == 1 && !all_bound ? "  [mbind refused: cpuset?]" : "");
            for (int d = 0; d < G; ++d) {
                CK(cudaSetDevice(d));
                if (mode == 0) cudaFreeHost(hbuf[d]); else { cudaHostUnregister(hbuf[d]); munmap(hbuf[d], bytes); }
                cudaFree(dbuf[d]); cudaStreamDestroy(st[d]);
            }
        }
    }
    return 0;
}
