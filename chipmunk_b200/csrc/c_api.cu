// ABI housekeeping: version, error strings, device queries.
#include <cuda_runtime.h>
#include <stdlib.h>

#include <mutex>

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

static int load_encoder() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn) return e != cudaSuccess ? (int)e : (int)cudaErrorSymbolNotFound;
    g_encode = (EncodeTiledFn)fn;
    return 0;
}

// ---- a small cache of encoded tensor maps: the operands of a denoising step (weights, caches, q/k/v buffers of a
// compiled block) come back with the same address / shape / strides step after step, and cuTensorMapEncodeTiled is a
// driver call of several microseconds -- comparable to a short kernel.  Keyed by every argument of the encoding and the
// current device; 64 entries, round-robin replacement; guarded by a mutex (ops may be called from several threads).
struct TmapKey {
    const void* base; uint64_t d[4]; uint64_t s[3]; uint32_t box[4]; int rank, swz, dev;
    bool operator==(const TmapKey& o) const {
        if (base != o.base || rank != o.rank || swz != o.swz || dev != o.dev) return false;
        for (int i = 0; i < 4; i++) if (box[i] != o.box[i]) return false;
        for (int i = 0; i < 4; i++) if (d[i] != o.d[i]) return false;
        for (int i = 0; i < 3; i++) if (s[i] != o.s[i]) return false;
        return true;
    }
};
constexpr int TMAP_CACHE = 64;
static TmapKey g_keys[TMAP_CACHE];
static CUtensorMap g_maps[TMAP_CACHE];
static int g_used = 0, g_next = 0;
static std::mutex g_tmap_mutex;

static int encode_cached(CUtensorMap* map, const TmapKey& key) {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    for (int i = 0; i < g_used; i++)
        if (g_keys[i] == key) { *map = g_maps[i]; return 0; }
    int rc = load_encoder();
    if (rc) return rc;
    cuuint64_t gdim[4]; cuuint64_t gstr[3]; cuuint32_t box[4] = {1, 1, 1, 1}; cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int i = 0; i < key.rank; i++) gdim[i] = key.d[i];
    for (int i = 0; i + 1 < key.rank; i++) gstr[i] = key.s[i];
    for (int i = 0; i < key.rank; i++) box[i] = key.box[i];
    const CUtensorMapSwizzle sw = key.swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : key.swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : key.swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)key.rank, const_cast<void*>(key.base), gdim, gstr,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return (int)cudaErrorInvalidValue;
    const int slot = g_used < TMAP_CACHE ? g_used++ : (g_next++ % TMAP_CACHE);
    g_keys[slot] = key;
    g_maps[slot] = *map;
    return 0;
}

int cached_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                        uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
    TmapKey k{};
    k.base = base; k.rank = 2; k.swz = swizzle_bytes; k.d[0] = cols; k.d[1] = rows; k.s[0] = pitch_bytes;
    k.box[0] = box_cols; k.box[1] = box_rows; k.box[2] = 1; k.box[3] = 1;
    cudaGetDevice(&k.dev);
    return encode_cached(map, k);
}

int encode_tmap_4d_bf16_sw128(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                              const uint32_t box[4]) {
    TmapKey k{};
    k.base = base; k.rank = 4; k.swz = 128;
    for (int i = 0; i < 4; i++) k.d[i] = dims[i];
    for (int i = 0; i < 3; i++) k.s[i] = strides_bytes[i];
    for (int i = 0; i < 4; i++) k.box[i] = box[i];
    cudaGetDevice(&k.dev);
    return encode_cached(map, k);
}

int encode_tmap_bhnd(CUtensorMap* m, const void* base, int B, int H, int N, const int64_t st[3], uint32_t box_rows, int8_t pos[3]) {
    struct Dim { uint64_t size, stride; int role; } d[3] = {{(uint64_t)N, (uint64_t)st[2] * 2, 0}, {(uint64_t)H, (uint64_t)st[1] * 2, 1},
                                                             {(uint64_t)B, (uint64_t)st[0] * 2, 2}};
    auto key = [](const Dim& x) { return x.size == 1 ? ~0ull : x.stride; };
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++)
            if (key(d[j]) < key(d[i])) { Dim t = d[i]; d[i] = d[j]; d[j] = t; }
    uint64_t dims[4] = {128, 0, 0, 0}, strides[3];
    uint32_t box[4] = {64, 1, 1, 1};
    uint64_t prev = 256;
    for (int i = 0; i < 3; i++) {
        dims[i + 1] = d[i].size;
        // a size-1 dimension's stride is never used for addressing; give it a valid monotone value
        strides[i] = d[i].size == 1 ? prev : d[i].stride;
        prev = strides[i] * d[i].size;
        pos[d[i].role] = (int8_t)(i + 1);
        if (d[i].role == 0) box[i + 1] = box_rows;
    }
    return encode_tmap_4d_bf16_sw128(m, base, dims, strides, box);
}

int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                        uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
    return cached_tmap_2d_bf16(map, base, rows, cols, pitch_bytes, box_cols, box_rows, swizzle_bytes);
}

int encode_tmap_2d_bf16_sw128(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t pitch_bytes, uint32_t box_rows) {
    return cached_tmap_2d_bf16(map, base, rows, cols, pitch_bytes, 64, box_rows, 128);
}
}  // namespace cm
