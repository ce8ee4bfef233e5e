// Column-sparse MLP GEMMs for sm_100a (placeholder entry points until the tcgen05 kernels land).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/chipmunk_b200.h"

extern "C" int cm_csp_mlp_mm1(const void*, const void*, void*, const void*, void*, const int32_t*, const int32_t*,
                              int, int, int, int64_t, int, void*) { return CM_EUNSUPPORTED; }
extern "C" int cm_csp_mlp_mm2(const void*, const void*, void*, void*, const int32_t*, const int32_t*, int, int, int,
                              int64_t, int, void*) { return CM_EUNSUPPORTED; }
extern "C" int cm_csp_scatter_add(const void*, void*, const int32_t*, const int32_t*, int, int, int64_t, void*) {
    return CM_EUNSUPPORTED;
}
