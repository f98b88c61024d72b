// pqp_device_guard.h — every handle of the library is bound to one CUDA device; an API call
// switches to it for its own duration and leaves the caller's current device as it found it.
#pragma once
#include <cuda_runtime.h>

namespace pqp {
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
}  // namespace pqp
