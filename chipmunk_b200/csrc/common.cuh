// Small host/device helpers shared by every translation unit of libchipmunk_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm {

template <typename T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Number of SMs of the CURRENT device, cached per device (persistent kernels size their grid
// with it; 148 on B200).
int sm_count();
// True when the current device is compute capability 10.x.
bool is_sm100();
// Opt `func` into `bytes` of dynamic shared memory once per device; `mask` is the call site's static bit set
// (one bit per device ordinal).  Returns 0 or a cudaError_t.
int opt_in_dynamic_smem(unsigned long long& mask, const void* func, int bytes);
// CM_DEBUG_FLAGS of the environment, read once per process.  Timing experiments only (stage isolation: skip the
// gather / the MMAs / the softmax / the stores); kernels produce WRONG results when it is non-zero.
int debug_flags();

}  // namespace cm
