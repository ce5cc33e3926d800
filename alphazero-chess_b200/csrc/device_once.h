// cudaFuncSetAttribute and similar calls are per device: this remembers, per CUDA device, whether a one-time set-up ran.
#pragma once
#include <cuda_runtime.h>

namespace azb {
struct PerDeviceOnce {
    bool done[64] = {};
    // true exactly once per device (always true for ordinals outside the table, so nothing is ever skipped wrongly)
    bool first() {
        int d = -1;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};
}  // namespace azb
