// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map encoding without linking libcuda
// (the driver entry point is fetched through the runtime), and the device-side 2-D tile load.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace cm {

// Encode a row-major 2-D bf16 tensor [rows, cols] (row pitch `pitch_bytes`) whose tiles of
// `box_rows` x 64 columns land in shared memory as [box_rows][128 B] with SWIZZLE_128B -- the
// K-major operand layout tcgen05.mma reads (see ptx.cuh: umma_smem_desc).
// Returns 0 or a CUresult.
int encode_tmap_2d_bf16_sw128(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t pitch_bytes, uint32_t box_rows);

// General form: boxes of `box_rows` x `box_cols` elements; `swizzle_bytes` in {0, 32, 64, 128} must be >= the
// box's inner extent in bytes (or 0).  Used for the epilogue stores (shared -> global boxes).
int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                        uint32_t box_cols, uint32_t box_rows, int swizzle_bytes);

// 4-D bf16 tensor [d3][d2][d1][d0] with d0 contiguous (a [B,H,N,128] attention operand with arbitrary batch / head / row
// strides): boxes of box[0..3] elements, SWIZZLE_128B (box[0] = 64).  `strides_bytes` = byte strides of d1, d2, d3
// (multiples of 16).  Out-of-range coordinates read as zero / are not written.  Results are cached per argument set
// (encoding costs microseconds on the host and the operands of a diffusion step recur), see c_api.cu.
int encode_tmap_4d_bf16_sw128(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                              const uint32_t box[4]);
// [B,H,N,128] bf16 view with element strides st = {batch, head, row}: a 4-D tensor map whose outer dimensions are ordered
// by ascending stride (size-1 dimensions last), boxes of `box_rows` rows x 64 columns.  pos[] = the coordinate slot
// (1..3) where (row, head, batch) landed: see tma_coords().
int encode_tmap_bhnd(CUtensorMap* map, const void* base, int B, int H, int N, const int64_t st[3], uint32_t box_rows, int8_t pos[3]);
// cached forms of the 2-D encoders (same arguments, same result)
int cached_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                        uint32_t box_cols, uint32_t box_rows, int swizzle_bytes);

// One thread: load the tile whose top-left element is (row, col) into `dst` (1024-byte aligned),
// completing `bytes` on `bar`.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col, int row) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes.cta_group::1 "
        "[%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst), "l"(map), "r"(col), "r"(row), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cta.global.tile.mbarrier::complete_tx::bytes.cta_group::1 "
        "[%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src_smem)
                 : "memory");
}
// coordinates (row, head, batch) placed at the slots an encode_tmap_bhnd map wants them
struct TmaCoord { int c[4]; };
__device__ __forceinline__ TmaCoord tma_coords(const int8_t pos[3], int col, int row, int h, int b) {
    TmaCoord r;
    r.c[0] = col;
#pragma unroll
    for (int i = 1; i < 4; i++) r.c[i] = pos[0] == i ? row : (pos[1] == i ? h : b);
    return r;
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];\n" ::"l"(map),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src_smem)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(map) : "memory");
}

}  // namespace cm
