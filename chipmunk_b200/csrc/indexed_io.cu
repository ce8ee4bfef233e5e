// Index selection / pack-unpack kernels for sm_100a (integer path, bit-exact).
//
// Replaces csrc/indexed_io/{mask_to_indices,topk_indices,copy_indices}.cu and the torch-compiled
// bit codec in src/chipmunk/ops/bitpack.py of the reference.  All kernels are HBM-bound byte /
// integer work: coalesced 128-bit loads, shared-memory staging of one row as 32-bit words,
// warp ballots + scans for compaction.  No tensor cores on purpose.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"

namespace cm {

// ------------------------------------------------------------------------------------------
// bit codec: 16 mask bytes <-> 2 packed bytes per thread
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nibble_of_bools(uint32_t x) {
    // x holds four bytes; any non-zero byte counts as true.  -> 4-bit little-endian nibble
    x = __vcmpne4(x, 0u) & 0x01010101u;
    return (x * 0x10204080u) >> 28;
}
__device__ __forceinline__ uint32_t bools_of_nibble(uint32_t nib) {
    return ((nib & 0xFu) * 0x00204081u) & 0x01010101u;
}

__global__ void __launch_bounds__(256) bitpack_kernel(const uint8_t* __restrict__ mask,
                                                      uint8_t* __restrict__ packed, int64_t n) {
    const int64_t n16 = n / 16;   // full 16-byte groups
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(mask) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(packed) & 1) == 0);
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += stride) {
        uint4 m;
        if (aligned) {
            m = __ldg(reinterpret_cast<const uint4*>(mask) + g);
        } else {
            const uint8_t* p = mask + g * 16;
            uint32_t w[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                w[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
            m = make_uint4(w[0], w[1], w[2], w[3]);
        }
        uint32_t bits = nibble_of_bools(m.x) | (nibble_of_bools(m.y) << 4) |
                        (nibble_of_bools(m.z) << 8) | (nibble_of_bools(m.w) << 12);
        if (aligned) {
            reinterpret_cast<uint16_t*>(packed)[g] = (uint16_t)bits;
        } else {
            packed[2 * g] = (uint8_t)bits;
            packed[2 * g + 1] = (uint8_t)(bits >> 8);
        }
    }
    // tail: fewer than 16 mask bytes -> at most 2 packed bytes, done by one thread
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t base = n16 * 16;
        uint32_t bits = 0;
        for (int64_t i = base; i < n; i++) bits |= (mask[i] != 0 ? 1u : 0u) << (i - base);
        int64_t nb = (n - base + 7) / 8;
        for (int64_t j = 0; j < nb; j++) packed[base / 8 + j] = (uint8_t)(bits >> (8 * j));
    }
}

__global__ void __launch_bounds__(256) bitunpack_kernel(const uint8_t* __restrict__ packed,
                                                        uint8_t* __restrict__ mask, int64_t n) {
    const int64_t n16 = n / 16;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(mask) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(packed) & 1) == 0);
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n16; g += stride) {
        uint32_t bits = aligned ? (uint32_t)__ldg(reinterpret_cast<const uint16_t*>(packed) + g)
                                : (uint32_t)packed[2 * g] | ((uint32_t)packed[2 * g + 1] << 8);
        uint4 m = make_uint4(bools_of_nibble(bits), bools_of_nibble(bits >> 4),
                             bools_of_nibble(bits >> 8), bools_of_nibble(bits >> 12));
        if (aligned) {
            reinterpret_cast<uint4*>(mask)[g] = m;
        } else {
            uint32_t w[4] = {m.x, m.y, m.z, m.w};
            for (int i = 0; i < 16; i++) mask[g * 16 + i] = (uint8_t)(w[i / 4] >> (8 * (i % 4)));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int64_t i = n16 * 16; i < n; i++) mask[i] = (packed[i >> 3] >> (i & 7)) & 1;
    }
}

// ------------------------------------------------------------------------------------------
// mask (bool bytes or packed bits) -> indices, counts.   One CTA per [b,h,m] row.
//
// The row is staged in shared memory as W = ceil(n/32) words, word i bit c = mask[32 i + c].
// The reference emits set columns in (c, i) order -- lane c of its single warp owns columns
// c, c+32, ... (mask_to_indices.cu:47-68).  Thread (warp w, lane c) here owns class c of the
// w-th slice of words; an exclusive scan over (c, w) gives every thread its output cursor.
// ------------------------------------------------------------------------------------------
constexpr int M2I_THREADS = 256;
constexpr int M2I_WARPS = M2I_THREADS / 32;

template <bool PACKED>
__global__ void __launch_bounds__(M2I_THREADS)
mask_to_indices_kernel(const uint8_t* __restrict__ src, int32_t* __restrict__ indices,
                       int32_t* __restrict__ counts, int n, int pad_n, int multiple_of,
                       int64_t total_bytes) {
    extern __shared__ uint32_t words[];          // [W]
    __shared__ int seg_cnt[M2I_WARPS][32];
    __shared__ int cursor[M2I_WARPS][32];
    __shared__ int s_total;

    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = (n + 31) >> 5;

    // ---- stage 1: build the row's words
    if (PACKED) {
        const int64_t bit0 = row * (int64_t)n;
        for (int i = tid; i < W; i += M2I_THREADS) {
            const int64_t off = bit0 + 32ll * i;
            const int64_t byte = off >> 3;
            const int sh = (int)(off & 7);
            uint64_t acc = 0;
#pragma unroll
            for (int b = 0; b < 5; b++) {
                int64_t a = byte + b;
                uint32_t v = (a < total_bytes) ? (uint32_t)__ldg(src + a) : 0u;
                acc |= (uint64_t)v << (8 * b);
            }
            uint32_t w = (uint32_t)(acc >> sh);
            int rem = n - 32 * i;
            if (rem < 32) w &= (1u << rem) - 1u;
            words[i] = w;
        }
    } else {
        const uint8_t* rowp = src + row * (int64_t)n;
        const bool vec = ((reinterpret_cast<uintptr_t>(rowp) & 15) == 0);
        for (int i = tid; i < W; i += M2I_THREADS) {
            uint32_t w = 0;
            if (vec && 32 * i + 32 <= n) {
                uint4 lo = __ldg(reinterpret_cast<const uint4*>(rowp + 32 * i));
                uint4 hi = __ldg(reinterpret_cast<const uint4*>(rowp + 32 * i) + 1);
                w = nibble_of_bools(lo.x) | (nibble_of_bools(lo.y) << 4) |
                    (nibble_of_bools(lo.z) << 8) | (nibble_of_bools(lo.w) << 12) |
                    (nibble_of_bools(hi.x) << 16) | (nibble_of_bools(hi.y) << 20) |
                    (nibble_of_bools(hi.z) << 24) | (nibble_of_bools(hi.w) << 28);
            } else {
                int rem = min(32, n - 32 * i);
                for (int c = 0; c < rem; c++) w |= (rowp[32 * i + c] != 0 ? 1u : 0u) << c;
            }
            words[i] = w;
        }
    }
    __syncthreads();

    // ---- stage 2: per (slice, class) population
    const int S = (W + M2I_WARPS - 1) / M2I_WARPS;
    const int i0 = min(W, warp * S), i1 = min(W, i0 + S);
    int cnt = 0;
    for (int i = i0; i < i1; i++) cnt += (words[i] >> lane) & 1u;
    seg_cnt[warp][lane] = cnt;
    __syncthreads();

    // ---- stage 3: cursors.  class totals -> exclusive scan over classes (warp 0)
    if (warp == 0) {
        int tot = 0;
#pragma unroll
        for (int w = 0; w < M2I_WARPS; w++) tot += seg_cnt[w][lane];
        int incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        int run = incl - tot;
#pragma unroll
        for (int w = 0; w < M2I_WARPS; w++) {
            cursor[w][lane] = run;
            run += seg_cnt[w][lane];
        }
        if (lane == 31) s_total = incl;
    }
    __syncthreads();

    // ---- stage 4: emit set columns
    int32_t* out = indices + row * (int64_t)pad_n;
    int pos = cursor[warp][lane];
    for (int i = i0; i < i1; i++) {
        if ((words[i] >> lane) & 1u) out[pos++] = 32 * i + lane;
    }

    // ---- stage 5: pad with the first unset columns (ascending), write the count
    const int total = s_total;
    const int padded = ((total + multiple_of - 1) / multiple_of) * multiple_of;
    if (warp == 0) {
        int need = padded - total;
        int done = 0;
        for (int base = 0; base < W && done < need; base += 32) {
            int i = base + lane;
            uint32_t inv = 0;
            if (i < W) {
                inv = ~words[i];
                int rem = n - 32 * i;
                if (rem < 32) inv &= (1u << rem) - 1u;
            }
            int pc = __popc(inv);
            int incl = pc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int r = done + incl - pc;
            while (inv && r < need) {
                int c = __ffs(inv) - 1;
                inv &= inv - 1;
                out[total + r] = 32 * i + c;
                r++;
            }
            done += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) counts[row] = padded;
    }
}

// ------------------------------------------------------------------------------------------
// topk_indices: one CTA (1024 threads) per (row, batch), like the reference, so that the
// per-thread XORWOW streams and their draw order are the same as topk_indices.cu:44-49,111.
// ------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 1024;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(TOPK_THREADS)
topk_indices_kernel(const T* __restrict__ act, int32_t* __restrict__ indices,
                    int32_t* __restrict__ counts, int rows, int cols, float quantile,
                    int multiple_of, float random_amount) {
    __shared__ float sorted[TOPK_THREADS];
    __shared__ int warp_keep[32], warp_rej[32];
    __shared__ int pad_buf[TOPK_THREADS];

    const int row = blockIdx.x, batch = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t roff = ((size_t)batch * rows + row) * cols;
    act += roff;
    indices += roff;
    counts += (size_t)batch * rows + row;

    if (quantile == 0.f) {          // keep everything (topk_indices.cu:51-60)
        for (int c = tid; c < cols; c += TOPK_THREADS) indices[c] = c;
        if (tid == 0) counts[0] = cols;
        return;
    }
    if (quantile == 1.f) {          // keep nothing (topk_indices.cu:61-70)
        for (int c = tid; c < cols; c += TOPK_THREADS) indices[c] = -1;
        if (tid == 0) counts[0] = 0;
        return;
    }

    // ---- threshold: bitonic sort of the first 1024 values, ascending
    sorted[tid] = tid < cols ? to_f32(act[tid]) : __int_as_float(0x7f800000);
    __syncthreads();
    for (int k = 2; k <= TOPK_THREADS; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            int p = tid ^ j;
            if (p > tid) {
                float a = sorted[tid], b = sorted[p];
                bool up = (tid & k) == 0;
                if ((a > b) == up) { sorted[tid] = b; sorted[p] = a; }
            }
            __syncthreads();
        }
    }
    const float thr = sorted[(int)(TOPK_THREADS * quantile)];

    curandState st;
    const bool use_rng = random_amount > 0.f;
    if (use_rng) {
        unsigned long long seed = (unsigned long long)blockIdx.x * blockDim.x +
                                  (unsigned long long)blockIdx.y * gridDim.x * blockDim.x + threadIdx.x;
        seed += reinterpret_cast<const int*>(act)[0];
        curand_init(seed, 0, 0, &st);
    }

    // ---- ordered compaction, 1024 columns per sweep
    int n_keep = 0, n_rej = 0;
    for (int base = 0; base < cols; base += TOPK_THREADS) {
        const int col = base + tid;
        const bool valid = col < cols;
        bool keep = false;
        if (valid) {
            keep = to_f32(act[col]) >= thr;
            if (!keep && use_rng) keep = curand_uniform(&st) < random_amount;
        }
        const unsigned kb = __ballot_sync(0xffffffffu, keep);
        const unsigned rb = __ballot_sync(0xffffffffu, valid && !keep);
        if (lane == 0) { warp_keep[warp] = __popc(kb); warp_rej[warp] = __popc(rb); }
        __syncthreads();
        int kpre = 0, rpre = 0, ktot = 0, rtot = 0;
#pragma unroll
        for (int w = 0; w < 32; w++) {
            int a = warp_keep[w], b = warp_rej[w];
            if (w < warp) { kpre += a; rpre += b; }
            ktot += a; rtot += b;
        }
        const unsigned lt = (1u << lane) - 1u;
        if (keep) indices[n_keep + kpre + __popc(kb & lt)] = col;
        else if (valid) {
            int r = n_rej + rpre + __popc(rb & lt);
            if (r < TOPK_THREADS) pad_buf[r] = col;
        }
        n_keep += ktot;
        n_rej += rtot;
        __syncthreads();
    }
    // ---- pad to a multiple with the first rejected columns (topk_indices.cu:126-140)
    const int mod = n_keep % multiple_of;
    const int r = mod == 0 ? 0 : multiple_of - mod;
    if (tid < r && tid < n_rej) indices[n_keep + tid] = pad_buf[tid];
    if (tid == 0) counts[0] = n_keep + r;
}

// ------------------------------------------------------------------------------------------
// copy_indices: dst[b, r, idx] = src[b, r, idx]
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
copy_indices_kernel(const T* __restrict__ src, T* __restrict__ dst, const int32_t* __restrict__ inds,
                    const int32_t* __restrict__ counts, int M, int Rr, int F) {
    const int64_t grow = blockIdx.x;             // over B * M * Rr
    const int64_t bm = grow / Rr;                // b * M + m
    const int n = counts[bm];
    const int32_t* ip = inds + bm * (int64_t)F;
    const T* s = src + grow * (int64_t)F;
    T* d = dst + grow * (int64_t)F;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        int f = __ldg(ip + c);
        if ((unsigned)f < (unsigned)F) d[f] = s[f];
    }
}

}  // namespace cm

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace cm;

extern "C" int cm_bitpack(const uint8_t* mask, uint8_t* packed, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!mask || !packed))) return CM_EINVAL;
    if (n == 0) return CM_OK;
    int64_t groups = n / 16;
    int64_t want = (groups + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    bitpack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(mask, packed, n);
    return (int)cudaGetLastError();
}

extern "C" int cm_bitunpack(const uint8_t* packed, uint8_t* mask, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && (!mask || !packed))) return CM_EINVAL;
    if (n == 0) return CM_OK;
    int64_t groups = n / 16;
    int64_t want = (groups + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    bitunpack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(packed, mask, n);
    return (int)cudaGetLastError();
}

template <bool PACKED>
static int launch_m2i(const uint8_t* src, int32_t* indices, int32_t* counts, int64_t rows, int n,
                      int pad_n, int multiple_of, void* stream) {
    if (rows < 0 || n <= 0 || pad_n < n || multiple_of <= 0) return CM_EINVAL;
    if (rows == 0) return CM_OK;
    if (!src || !indices || !counts) return CM_EINVAL;
    if (rows > 2147483647ll) return CM_EINVAL;
    size_t smem = (size_t)((n + 31) / 32) * 4;
    if (smem > 200 * 1024) return CM_EUNSUPPORTED;
    auto kern = mask_to_indices_kernel<PACKED>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int64_t total_bytes = PACKED ? (rows * (int64_t)n + 7) / 8 : rows * (int64_t)n;
    kern<<<(unsigned)rows, M2I_THREADS, smem, (cudaStream_t)stream>>>(src, indices, counts, n, pad_n,
                                                                     multiple_of, total_bytes);
    return (int)cudaGetLastError();
}

extern "C" int cm_mask_to_indices(const uint8_t* mask, int32_t* indices, int32_t* counts,
                                  int64_t rows, int n, int pad_n, int multiple_of, void* stream) {
    return launch_m2i<false>(mask, indices, counts, rows, n, pad_n, multiple_of, stream);
}
extern "C" int cm_bitmask_to_indices(const uint8_t* packed, int32_t* indices, int32_t* counts,
                                     int64_t rows, int n, int pad_n, int multiple_of, void* stream) {
    return launch_m2i<true>(packed, indices, counts, rows, n, pad_n, multiple_of, stream);
}

extern "C" int cm_topk_indices(const void* act, int dtype, int32_t* indices, int32_t* counts, int B,
                               int R, int C, float sparsity, int multiple_of, float random_amount,
                               void* stream) {
    if (B < 0 || R < 0 || C <= 0 || multiple_of <= 0 || multiple_of > TOPK_THREADS) return CM_EINVAL;
    if (!(sparsity >= 0.f && sparsity <= 1.f)) return CM_EINVAL;
    if (B == 0 || R == 0) return CM_OK;
    if (!act || !indices || !counts) return CM_EINVAL;
    dim3 grid(R, B);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case CM_BF16:
            topk_indices_kernel<__nv_bfloat16><<<grid, TOPK_THREADS, 0, s>>>(
                (const __nv_bfloat16*)act, indices, counts, R, C, sparsity, multiple_of, random_amount);
            break;
        case CM_F16:
            topk_indices_kernel<__half><<<grid, TOPK_THREADS, 0, s>>>(
                (const __half*)act, indices, counts, R, C, sparsity, multiple_of, random_amount);
            break;
        case CM_F32:
            topk_indices_kernel<float><<<grid, TOPK_THREADS, 0, s>>>(
                (const float*)act, indices, counts, R, C, sparsity, multiple_of, random_amount);
            break;
        default:
            return CM_EUNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

extern "C" int cm_copy_indices(const void* src, void* dst, int elem_size, const int32_t* indices,
                               const int32_t* counts, int B, int M, int Rr, int F, void* stream) {
    if (B < 0 || M < 0 || Rr <= 0 || F <= 0) return CM_EINVAL;
    int64_t rows = (int64_t)B * M * Rr;
    if (rows == 0) return CM_OK;
    if (!src || !dst || !indices || !counts || rows > 2147483647ll) return CM_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_size == 2)
        copy_indices_kernel<uint16_t><<<(unsigned)rows, 256, 0, s>>>((const uint16_t*)src, (uint16_t*)dst,
                                                                    indices, counts, M, Rr, F);
    else if (elem_size == 4)
        copy_indices_kernel<uint32_t><<<(unsigned)rows, 256, 0, s>>>((const uint32_t*)src, (uint32_t*)dst,
                                                                    indices, counts, M, Rr, F);
    else
        return CM_EUNSUPPORTED;
    return (int)cudaGetLastError();
}
