// ABI housekeeping: version, error strings, device queries.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"

namespace cm {
constexpr int MAX_DEV = 64;
static int g_sms[MAX_DEV] = {0}, g_major[MAX_DEV] = {0};
static int query() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return -1;
    if (!g_sms[dev]) {
        cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&g_major[dev], cudaDevAttrComputeCapabilityMajor, dev);
    }
    return dev;
}
int sm_count() { const int d = query(); return d >= 0 && g_sms[d] > 0 ? g_sms[d] : 148; }
bool is_sm100() { const int d = query(); return d >= 0 && g_major[d] == 10; }
int debug_flags() {
#ifdef CM_DEBUG_STAGES
    static const int flags = [] { const char* e = getenv("CM_DEBUG_FLAGS"); return e ? atoi(e) : 0; }();
    return flags;
#else
    return 0;
#endif
}
int opt_in_dynamic_smem(unsigned long long& mask, const void* func, int bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64 && ((mask >> dev) & 1ull)) return 0;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0 && dev < 64) mask |= 1ull << dev;
    return 0;
}
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

// ------------------------------------------------------------------------------------------
#include "tma.cuh"
namespace cm {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                        uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn) return e != cudaSuccess ? (int)e : (int)cudaErrorSymbolNotFound;
        g_encode = (EncodeTiledFn)fn;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

int encode_tmap_2d_bf16_sw128(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t pitch_bytes, uint32_t box_rows) {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || !fn) return e != cudaSuccess ? (int)e : (int)cudaErrorSymbolNotFound;
        g_encode = (EncodeTiledFn)fn;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {pitch_bytes};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}
}  // namespace cm
