// device_guard.h — makes a CUDA device current for the duration of an entry point and puts the caller's device back
// afterwards (a caller that works on cuda:0 and owns a handle on cuda:1 must not find its current device switched).
#pragma once
#include <cuda_runtime.h>

namespace ktb {

struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) err = cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

}  // namespace ktb
