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
// Stage-isolation timing switches (skip the gather / the MMAs / the softmax / the stores: results are WRONG).
// They exist only in builds made with -DCM_DEBUG_STAGES (python chipmunk_b200/build.py --debug-stages); in the
// shipped library CM_DBG() is the constant false, the branches are compiled out and debug_flags() returns 0
// whatever the environment says.
int debug_flags();
#ifdef CM_DEBUG_STAGES
#define CM_DBG(P, bit) (((P).dbg & (bit)) != 0)
#else
#define CM_DBG(P, bit) false
#endif

}  // namespace cm
