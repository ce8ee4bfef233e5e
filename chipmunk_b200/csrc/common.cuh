// Small host/device helpers shared by every translation unit of libchipmunk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm {

template <typename T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Number of SMs of the current device, cached per process (persistent kernels size their grid
// with it; 148 on B200).
int sm_count();
// True when the current device is compute capability 10.x.
bool is_sm100();

}  // namespace cm
