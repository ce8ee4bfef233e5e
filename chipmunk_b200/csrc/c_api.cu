// ABI housekeeping: version, error strings, device queries.
#include <cuda_runtime.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"

namespace cm {
static int g_sms = 0, g_major = 0;
static void query() {
    if (g_sms) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&g_major, cudaDevAttrComputeCapabilityMajor, dev);
}
int sm_count() { query(); return g_sms > 0 ? g_sms : 148; }
bool is_sm100() { query(); return g_major == 10; }
}  // namespace cm

extern "C" int cm_abi_version(void) { return CM_ABI_VERSION; }
extern "C" int cm_sm_count(void) { return cm::sm_count(); }
extern "C" const char* cm_strerror(int code) {
    switch (code) {
        case CM_OK: return "ok";
        case CM_EINVAL: return "chipmunk_b200: invalid argument (shape, multiple or null pointer)";
        case CM_EALIGN: return "chipmunk_b200: pointer or stride is not 16-byte aligned";
        case CM_EUNSUPPORTED: return "chipmunk_b200: unsupported dtype / size for this kernel";
        case CM_EARCH: return "chipmunk_b200: device is not sm_100 (these kernels are sm_100a only)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "chipmunk_b200: unknown error";
}
