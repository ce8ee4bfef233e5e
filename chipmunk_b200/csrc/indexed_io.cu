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
// The index list is assembled in SHARED memory (a scatter there costs bank conflicts, not DRAM
// sectors) and leaves the SM as one coalesced run of 16-byte stores: the "warp-scan + shared-memory
// transpose" of north_star (iii).  Lists longer than the staging buffer are written directly.
// ------------------------------------------------------------------------------------------
constexpr int M2I_THREADS = 512;
constexpr int M2I_WARPS = M2I_THREADS / 32;
constexpr int STAGE_INTS = 16384;              // 64 KB staging buffer for one row's index list

struct EmitScratch {
    int warp_tot[32];
    int total;
};

// 32 x 32 bit transpose across the lanes of a warp: lane l gives word x_l, lane c receives T_c with bit b of T_c = bit c
// of x_b (five butterfly stages).
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) == 0 ? ((x & m) | ((y & m) << j)) : ((x & ~m) | ((y & ~m) >> j));
    }
    return x;
}

// words[0..W) complete and visible to the whole CTA (caller synchronised).  All threads of the CTA call this.
// The reference's emission order -- set columns sorted by (col % 32, col) -- is the order of the set bits of the
// TRANSPOSED bit matrix: class c = col % 32 first, then the word index.  So: transpose the row's [W x 32] bit matrix
// 32 words at a time (warp butterflies), popcount + block-scan the transposed words, and let every thread expand its few
// words with ffs loops: work per row ~ W word operations instead of 32 W single-bit tests.
template <int NWARPS>
__device__ __forceinline__ void emit_row(const uint32_t* words, uint32_t* tw, int W, int n, int multiple_of,
                                         int32_t* __restrict__ out, int32_t* __restrict__ count_out, EmitScratch& sc,
                                         int32_t* stage) {
    constexpr int NT = NWARPS * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int WJ = (W + 31) >> 5;                 // groups of 32 words; tw[c * WJ + j] bit b = words[32 j + b] bit c
    for (int j = warp; j < WJ; j += NWARPS) {
        const int i = 32 * j + lane;
        const uint32_t t = warp_bit_transpose(i < W ? words[i] : 0u, lane);
        tw[lane * WJ + j] = t;
    }
    __syncthreads();
    // ---- popcounts of this thread's run of transposed words, block-wide exclusive scan
    const int TW = 32 * WJ;
    const int CH = (TW + NT - 1) / NT;
    const int w0 = min(TW, tid * CH), w1 = min(TW, w0 + CH);
    int cnt = 0;
    for (int i = w0; i < w1; i++) cnt += __popc(tw[i]);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sc.warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = lane < NWARPS ? sc.warp_tot[lane] : 0;
        int in2 = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, in2, d);
            if (lane >= d) in2 += t;
        }
        if (lane < NWARPS) sc.warp_tot[lane] = in2 - v;
        if (lane == 31) sc.total = in2;
    }
    __syncthreads();
    const int total = sc.total;
    const int padded = ((total + multiple_of - 1) / multiple_of) * multiple_of;
    const bool staged = stage != nullptr && padded <= STAGE_INTS;
    int32_t* dest = staged ? stage : out;
    // ---- emit set columns
    int pos = sc.warp_tot[warp] + incl - cnt;
    int c = w0 / WJ, j = w0 - c * WJ;             // class and word group of transposed word w0 (one division per thread)
    for (int i = w0; i < w1; i++) {
        uint32_t t = tw[i];
        const int col0 = 1024 * j + c;            // column of bit b: 32 (32 j + b) + c
        while (t) {
            const int bbit = __ffs(t) - 1;
            t &= t - 1;
            dest[pos++] = col0 + 32 * bbit;
        }
        if (++j == WJ) { j = 0; c++; }
    }
    // ---- pad with the first unset columns (ascending), write the count
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
            int incl2 = pc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl2, d);
                if (lane >= d) incl2 += t;
            }
            int r = done + incl2 - pc;
            while (inv && r < need) {
                int c = __ffs(inv) - 1;
                inv &= inv - 1;
                dest[total + r] = 32 * i + c;
                r++;
            }
            done += __shfl_sync(0xffffffffu, incl2, 31);
        }
        if (lane == 0) *count_out = padded;
    }
    if (staged) {
        __syncthreads();
        // a row of fewer unset columns than the padding asks for leaves a gap the reference leaves uninitialised too:
        // only [0, min(padded, n)) is defined
        const int nout = min(padded, n);
        if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
            const int n4 = nout >> 2;
            const int4* s4 = reinterpret_cast<const int4*>(stage);
            int4* o4 = reinterpret_cast<int4*>(out);
            for (int j = tid; j < n4; j += NT) o4[j] = s4[j];
            for (int j = (n4 << 2) + tid; j < nout; j += NT) out[j] = stage[j];
        } else {
            for (int j = tid; j < nout; j += NT) out[j] = stage[j];
        }
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(M2I_THREADS)
mask_to_indices_kernel(const uint8_t* __restrict__ src, int32_t* __restrict__ indices,
                       int32_t* __restrict__ counts, int n, int pad_n, int multiple_of,
                       int64_t total_bytes, int use_stage) {
    extern __shared__ __align__(16) uint32_t dyn_smem[];     // [STAGE_INTS if use_stage][W words][32 ceil(W/32) transposed words]
    __shared__ EmitScratch sc;
    int32_t* stage = use_stage ? reinterpret_cast<int32_t*>(dyn_smem) : nullptr;
    uint32_t* words = dyn_smem + (use_stage ? STAGE_INTS : 0);
    uint32_t* tw = words + ((n + 31) >> 5);

    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x;
    const int W = (n + 31) >> 5;

    // ---- stage 1: build the row's words
    if (PACKED) {
        const int64_t bit0 = row * (int64_t)n;
        for (int i = tid; i < W; i += M2I_THREADS) {
            const int64_t off = bit0 + 32ll * i;
            const int64_t byte = off >> 3;
            const int sh = (int)(off & 7);
            uint64_t acc = 0;
            if ((byte & 3) == 0 && byte + 8 <= total_bytes) {       // two aligned words instead of five byte loads
                const uint32_t* p32 = reinterpret_cast<const uint32_t*>(src + byte);
                acc = (uint64_t)__ldg(p32) | ((uint64_t)__ldg(p32 + 1) << 32);
            } else {
#pragma unroll
                for (int b = 0; b < 5; b++) {
                    int64_t a = byte + b;
                    uint32_t v = (a < total_bytes) ? (uint32_t)__ldg(src + a) : 0u;
                    acc |= (uint64_t)v << (8 * b);
                }
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
    emit_row<M2I_WARPS>(words, tw, W, n, multiple_of, indices + row * (int64_t)pad_n, counts + row, sc, stage);
}

// ------------------------------------------------------------------------------------------
// select_columns: the full step's column selection in ONE kernel.
// Replaces `random_and_topk` + `bitpack` + `mask_to_indices` of the reference's SparseDiffAttn
// (src/chipmunk/modules/attn.py:76-84,132-139: torch.randint + torch.topk over the [B,H,G,N] column sums +
// scatter_ + static-mask algebra + bit packing + the index kernel; and :141-150, torch.topk, on the uncompressed path).
// One CTA (1024 threads) per (b,h,g) row of the column sums cs[row, 0:n] (bf16):
//   pass A/B  exact k-th largest by a two-level radix select on the order-preserving 16-bit key of the bf16 value
//             (256-bin histograms, one private copy per lane so a warp never collides with itself; runs of equal
//             bins are merged in registers, which matters because column sums share their exponent bits);
//   pass C0   ties at the threshold are ranked by column so that EXACTLY k columns are kept (torch.topk's choice
//             among equal values is unspecified; ours is "lowest column first");
//   pass C    bit = top-k | hash(seed,row,col) < random_prob; the row's bit words are assembled in shared memory;
//   then      words = (words & group_is_sparse[g]) | static_words[g]   (reference :80-82), the flat little-endian
//             bit-packed mask is written (what bitpack() would store) and the index list + padded count are emitted
//             exactly as mask_to_indices would emit them from that mask.
// Every pass streams the row in warp-contiguous, lane-coalesced 16-byte loads; after pass A it comes from L2.
// ------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 1024;
constexpr int SEL_WARPS = 32;

__device__ __forceinline__ uint32_t bf16_key(uint32_t u) {       // order-preserving: larger float <=> larger key
    return (u & 0x8000u) ? (~u & 0xffffu) : (u | 0x8000u);
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
    return x;
}

struct SelParams {
    const __nv_bfloat16* cs;
    int64_t cs_row_stride;
    int n, k;
    uint32_t rand_thr16;          // keep if 16-bit hash < rand_thr16  (0: no random columns)
    uint32_t seed;
    const uint32_t* static_words; // [static_rows, static_stride] or null
    int64_t static_stride;
    int static_rows;
    const uint8_t* group_is_sparse;   // [static_rows] or null
    uint32_t* packed_words;       // flat bit-packed mask viewed as aligned 32-bit words, or null
    int32_t* indices;             // or null
    int32_t* counts;
    int pad_n, multiple_of;
};

__global__ void __launch_bounds__(SEL_THREADS, 1) select_columns_kernel(const SelParams P) {
    extern __shared__ __align__(16) uint32_t dyn_smem[];     // [STAGE_INTS][W words][32 ceil(W/32) transposed words]
    __shared__ uint32_t hist[256 * 32];
    __shared__ uint32_t tot[256];
    __shared__ EmitScratch sc;
    __shared__ int s_bin, s_krem;
    __shared__ int warp_ties[SEL_WARPS];

    int32_t* stage = reinterpret_cast<int32_t*>(dyn_smem);
    uint32_t* words = dyn_smem + STAGE_INTS;
    const int64_t row = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n = P.n, W = (n + 31) >> 5;
    uint32_t* tw = words + W;
    const __nv_bfloat16* rowp = P.cs + row * P.cs_row_stride;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(rowp) & 15) == 0;
    // warp w owns columns [w * CW, (w + 1) * CW), CW a multiple of 256: a warp step covers 256 columns, 8 per lane
    const int CW = (((n + SEL_WARPS - 1) / SEL_WARPS) + 255) & ~255;
    const int wbeg = warp * CW, wend = min(n, wbeg + CW);

    // keys are handled two at a time: a 32-bit word holds the order-preserving keys of columns (c, c + 1)
    auto key2 = [](uint32_t w) -> uint32_t {
        const uint32_t m = ((w >> 15) & 0x00010001u) * 0xFFFFu;          // 0xFFFF in every negative half
        return w ^ (m | 0x80008000u);                                      // negative: ~bits, non-negative: bits | 0x8000
    };
    // 8 columns of this lane -> 4 key pairs; columns at or past n read as key 0 with their valid bit cleared
    auto load8 = [&](int col0, uint32_t (&kp)[4]) -> uint32_t {
        if (vec_ok && col0 + 8 <= n) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(rowp + col0));
            kp[0] = key2(v.x); kp[1] = key2(v.y); kp[2] = key2(v.z); kp[3] = key2(v.w);
            return 0xffu;
        }
        uint32_t vm = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t w = 0;
            if (col0 + 2 * j < n) { w |= (uint32_t)__bfloat16_as_ushort(rowp[col0 + 2 * j]); vm |= 1u << (2 * j); }
            if (col0 + 2 * j + 1 < n) { w |= (uint32_t)__bfloat16_as_ushort(rowp[col0 + 2 * j + 1]) << 16; vm |= 2u << (2 * j); }
            kp[j] = key2(w);
        }
        return vm;
    };
    auto clear_hist = [&]() {
        for (int i = tid; i < 256 * 32; i += SEL_THREADS) hist[i] = 0;
    };
    // bins -> totals, then the bin holding the `want`-th largest: s_bin, s_krem = how many to take from that bin
    auto find_bin = [&](int want) {
        __syncthreads();
        for (int bb = warp * 8; bb < warp * 8 + 8; bb++) {
            uint32_t v = hist[bb * 32 + lane];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) tot[bb] = v;
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns bins 255 - 8 l ... 248 - 8 l (descending)
            int sum = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) sum += (int)tot[255 - 8 * lane - j];
            int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int excl = incl - sum;
            const int all = __shfl_sync(0xffffffffu, incl, 31);
            if (want > all) {                       // fewer valid columns than asked for: take everything
                if (lane == 0) { s_bin = -1; s_krem = 0; }
            } else if (excl < want && want <= incl) {
                int run = excl;
                for (int j = 0; j < 8; j++) {
                    const int bb = 255 - 8 * lane - j;
                    const int c = (int)tot[bb];
                    if (want <= run + c) { s_bin = bb; s_krem = want - run; break; }
                    run += c;
                }
            }
        }
        __syncthreads();
    };
    uint32_t* hl = hist + lane;                   // this lane's private copy of the 256 bins (bank = lane: a warp never conflicts)

    // ------------------------------------------------------------------ threshold (k-th largest key) and tie quota
    uint32_t T = 0x20000u;        // keep key > T, and `need` of the keys == T   (0x20000: keep none)
    int need = 0;
    if (P.k >= n) {
        T = 0; need = 0x7fffffff;                 // keep every column
    } else if (P.k > 0) {
        clear_hist();
        __syncthreads();
        // pass A: high byte of every key
        for (int col = wbeg + lane * 8; col < wend; col += 256) {
            uint32_t kp[4];
            const uint32_t vm = load8(col, kp);
            if (vm == 0xffu) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    atomicAdd(hl + ((kp[j] >> 3) & 0x1fe0u), 1u);          // ((k >> 8) & 0xff) * 32
                    atomicAdd(hl + ((kp[j] >> 19) & 0x1fe0u), 1u);         // (k >> 24) * 32
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if ((vm >> j) & 1u) atomicAdd(hl + ((((kp[j >> 1] >> (16 * (j & 1))) >> 8) & 0xffu) << 5), 1u);
            }
        }
        find_bin(P.k);
        const int b1 = s_bin, k1 = s_krem;
        __syncthreads();
        if (b1 < 0) {
            T = 0; need = 0x7fffffff;
        } else {
            clear_hist();
            __syncthreads();
            // pass B: low byte of the keys whose high byte is b1
            for (int col = wbeg + lane * 8; col < wend; col += 256) {
                uint32_t kp[4];
                const uint32_t vm = load8(col, kp);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t kk = (kp[j >> 1] >> (16 * (j & 1))) & 0xffffu;
                    if (((vm >> j) & 1u) && (int)(kk >> 8) == b1) atomicAdd(hl + ((kk & 0xffu) << 5), 1u);
                }
            }
            find_bin(k1);
            T = ((uint32_t)b1 << 8) | (uint32_t)s_bin;
            need = s_krem;                        // 1 <= need <= #(key == T)
            __syncthreads();
        }
    }
    const uint32_t T2 = T | (T << 16);           // the threshold in both halves (T <= 0xffff whenever it is compared)

    // ------------------------------------------------------------------ pass C0: ties per warp (column order)
    // only needed when the threshold value occurs more often than it may be kept
    int tie_base = 0;
    const bool rank_ties = need > 0 && need != 0x7fffffff && need < (int)tot[T & 0xffu];
    if (need > 0 && need != 0x7fffffff && !rank_ties) need = 0x7ffffffe;      // keep every key == T: no ranking
    if (rank_ties) {
        int e = 0;
        for (int col = wbeg + lane * 8; col < wend; col += 256) {
            uint32_t kp[4];
            const uint32_t vm = load8(col, kp);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t x = kp[j] ^ T2;                            // a zero half = a key equal to T
                e += ((x & 0xffffu) == 0 && ((vm >> (2 * j)) & 1u)) ? 1 : 0;
                e += ((x >> 16) == 0 && ((vm >> (2 * j + 1)) & 1u)) ? 1 : 0;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(0xffffffffu, e, d);
        if (lane == 0) warp_ties[warp] = e;
        __syncthreads();
        for (int w = 0; w < warp; w++) tie_base += warp_ties[w];
    }

    // ------------------------------------------------------------------ pass C: the row's bit words
    for (int i = tid; i < W; i += SEL_THREADS) words[i] = 0;
    __syncthreads();
    {
        const uint32_t rowseed = mix32(P.seed ^ mix32((uint32_t)row * 0x9E3779B1u + 0x7F4A7C15u));
        uint8_t* wbytes = reinterpret_cast<uint8_t*>(words);
        const bool keep_all_eq = need >= 0x7ffffffe;
        int running = tie_base;
        for (int col0 = wbeg; col0 < wend; col0 += 256) {
            const int col = col0 + lane * 8;
            uint32_t kp[4] = {0, 0, 0, 0};
            uint32_t vm = 0;
            if (col < wend) vm = load8(col, kp);
            uint32_t gt = 0, eq = 0;
            if (T <= 0xffffu) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t lo = kp[j] & 0xffffu, hi = kp[j] >> 16;
                    gt |= (lo > T ? 1u : 0u) << (2 * j) | (hi > T ? 2u : 0u) << (2 * j);
                    eq |= (lo == T ? 1u : 0u) << (2 * j) | (hi == T ? 2u : 0u) << (2 * j);
                }
                gt &= vm; eq &= vm;
            }
            uint32_t keep = gt;
            if (keep_all_eq) keep |= eq;
            else if (rank_ties && __any_sync(0xffffffffu, eq != 0)) {
                const int e = __popc(eq);
                int incl = e;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d) incl += t;
                }
                int r = running + incl - e;
                uint32_t m = eq;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    if (r < need) keep |= 1u << j;
                    r++;
                }
                running += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (P.rand_thr16) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const uint32_t h = mix32(((uint32_t)(col + j) >> 1) * 0x9E3779B1u ^ rowseed);
                    keep |= ((h & 0xffffu) < P.rand_thr16 ? 1u : 0u) << j;
                    keep |= ((h >> 16) < P.rand_thr16 ? 1u : 0u) << (j + 1);
                }
                keep &= vm;
            }
            if (col < wend) wbytes[col >> 3] = (uint8_t)keep;
        }
    }
    __syncthreads();
    // ------------------------------------------------------------------ static mask algebra (reference :80-82)
    if (P.static_words != nullptr || P.group_is_sparse != nullptr) {
        const int g = (int)(row % P.static_rows);
        const bool sparse = P.group_is_sparse == nullptr || P.group_is_sparse[g] != 0;
        const uint32_t* sw = P.static_words ? P.static_words + (int64_t)g * P.static_stride : nullptr;
        for (int i = tid; i < W; i += SEL_THREADS) {
            uint32_t w = sparse ? words[i] : 0u;
            if (sw) w |= __ldg(sw + i);
            const int rem = n - 32 * i;
            if (rem < 32) w &= (1u << rem) - 1u;
            words[i] = w;
        }
        __syncthreads();
    }
    // ------------------------------------------------------------------ the flat bit-packed mask
    if (P.packed_words != nullptr) {
        const int64_t bit0 = row * (int64_t)n;
        const int64_t gw0 = bit0 >> 5, gw1 = (bit0 + n - 1) >> 5;
        const int sh = (int)(bit0 & 31);
        const bool last_shared = ((bit0 + n) & 31) != 0;
        for (int64_t gw = gw0 + tid; gw <= gw1; gw += SEL_THREADS) {
            const int j = (int)(gw - gw0);
            const uint32_t lo = j < W ? words[j] : 0u;
            const uint32_t prev = (j > 0 && j - 1 < W) ? words[j - 1] : 0u;
            const uint32_t v = sh ? ((lo << sh) | (prev >> (32 - sh))) : lo;
            // words shared with the neighbouring rows are OR-ed into the (zero-initialised) buffer
            if ((gw == gw0 && sh != 0) || (gw == gw1 && last_shared)) { if (v) atomicOr(P.packed_words + gw, v); }
            else P.packed_words[gw] = v;
        }
    }
    // ------------------------------------------------------------------ indices + counts (mask_to_indices order)
    if (P.indices != nullptr)
        emit_row<SEL_WARPS>(words, tw, W, n, P.multiple_of, P.indices + row * (int64_t)P.pad_n, P.counts + row, sc, stage);
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
    const size_t W_ = (size_t)((n + 31) / 32);
    const size_t wbytes = (W_ + 32 * ((W_ + 31) / 32)) * 4;          // the row's words + their transpose
    if (wbytes > 128 * 1024) return CM_EUNSUPPORTED;
    // every thread emits a run of consecutive list positions (its transposed words are consecutive in emission order), so
    // the list goes straight to global memory in ~64-byte runs; a shared-memory staging copy only cost occupancy
    // (measured at the 720p shape: 0.69 ms with the 64 KB staging buffer, 2 CTAs / SM)
    const int use_stage = 0;
    const size_t smem = wbytes + (use_stage ? (size_t)STAGE_INTS * 4 : 0);
    auto kern = mask_to_indices_kernel<PACKED>;
    static unsigned long long configured = 0;
    int rc = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(kern), 200 * 1024);   // + 8 KB static
    if (rc) return rc;
    int64_t total_bytes = PACKED ? (rows * (int64_t)n + 7) / 8 : rows * (int64_t)n;
    kern<<<(unsigned)rows, M2I_THREADS, smem, (cudaStream_t)stream>>>(src, indices, counts, n, pad_n,
                                                                     multiple_of, total_bytes, use_stage);
    return (int)cudaGetLastError();
}

extern "C" int cm_select_columns(const void* cs, int64_t cs_row_stride, int64_t rows, int n, int k, float random_prob,
                                 uint64_t seed, const uint32_t* static_words, int64_t static_stride_words, int static_rows,
                                 const uint8_t* group_is_sparse, uint8_t* packed_out, int32_t* indices, int32_t* counts,
                                 int pad_n, int multiple_of, void* stream) {
    if (rows < 0 || n <= 0 || k < 0 || !(random_prob >= 0.f && random_prob < 1.f)) return CM_EINVAL;
    if (rows == 0) return CM_OK;
    if (!cs || cs_row_stride < n || rows > 2147483647ll) return CM_EINVAL;
    if (!packed_out && !indices) return CM_EINVAL;
    if (indices && (!counts || pad_n < n || multiple_of <= 0)) return CM_EINVAL;
    if ((static_words || group_is_sparse) && static_rows <= 0) return CM_EINVAL;
    if (static_words && static_stride_words < (n + 31) / 32) return CM_EINVAL;
    if (packed_out && (reinterpret_cast<uintptr_t>(packed_out) & 3)) return CM_EALIGN;
    const size_t W_ = (size_t)((n + 31) / 32);
    const size_t wbytes = (W_ + 32 * ((W_ + 31) / 32)) * 4;          // the row's words + their transpose
    if (wbytes > 112 * 1024) return CM_EUNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (packed_out && (n & 31)) {
        // rows share boundary words: those are OR-ed in, so the buffer starts from zero
        cudaError_t e = cudaMemsetAsync(packed_out, 0, (size_t)(((rows * (int64_t)n + 31) / 32) * 4), s);
        if (e != cudaSuccess) return (int)e;
    }
    SelParams P{};
    P.cs = (const __nv_bfloat16*)cs; P.cs_row_stride = cs_row_stride; P.n = n; P.k = k;
    uint32_t thr = (uint32_t)(random_prob * 65536.f + 0.5f);
    P.rand_thr16 = thr > 65535u ? 65535u : thr;
    P.seed = (uint32_t)(seed ^ (seed >> 32));
    P.static_words = static_words; P.static_stride = static_stride_words; P.static_rows = static_rows > 0 ? static_rows : 1;
    P.group_is_sparse = group_is_sparse;
    P.packed_words = reinterpret_cast<uint32_t*>(packed_out); P.indices = indices; P.counts = counts;
    P.pad_n = pad_n; P.multiple_of = multiple_of;
    static unsigned long long configured = 0;
    int rc = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(select_columns_kernel), 180 * 1024);   // + 42 KB static
    if (rc) return rc;
    const size_t smem = wbytes + (size_t)STAGE_INTS * 4;
    select_columns_kernel<<<(unsigned)rows, SEL_THREADS, smem, s>>>(P);
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

// ------------------------------------------------------------------------------------------
// Token reordering: dst[o, i, :] = src[o, perm[i], :].
// The reference's 2-level patchify / 3-D voxel chunking (src/chipmunk/ops/patch.py:7-80, ops/voxel.py:9-99) are fixed
// permutations of the token axis built from chains of einops rearranges (each a full copy); here the permutation is
// cached on the host side and ONE gather moves the data: rows of `row_bytes` bytes, 16 bytes per thread when the rows
// allow it (a [.., 128] bf16 token row is 16 such chunks: coalesced 256-byte reads and writes), else 2/4-byte elements.
// HBM-bound: bytes = 2 x tensor size + 4 n.
// ------------------------------------------------------------------------------------------
namespace cm {
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const V* __restrict__ src, V* __restrict__ dst,
                                                          const int32_t* __restrict__ perm, int64_t n_src, int64_t n_dst,
                                                          int chunks_per_row, int64_t total_chunks) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < total_chunks; c += stride) {
        const int64_t r = c / chunks_per_row;
        const int ch = (int)(c - r * chunks_per_row);
        const int64_t o = r / n_dst, i = r - o * n_dst;
        const int64_t j = __ldg(perm + i);
        dst[c] = __ldg(src + (o * n_src + j) * chunks_per_row + ch);
    }
}
}  // namespace cm

extern "C" int cm_gather_rows(const void* src, void* dst, const int32_t* perm, int64_t outer, int64_t n_src, int64_t n_dst,
                              int64_t row_bytes, void* stream) {
    if (outer < 0 || n_src <= 0 || n_dst < 0 || row_bytes <= 0) return CM_EINVAL;
    if (outer == 0 || n_dst == 0) return CM_OK;
    if (!src || !dst || !perm) return CM_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)row_bytes;
    const int unit = (al & 15) == 0 ? 16 : (al & 7) == 0 ? 8 : (al & 3) == 0 ? 4 : (al & 1) == 0 ? 2 : 1;
    const int64_t cpr = row_bytes / unit;
    if (cpr > 2147483647ll) return CM_EINVAL;
    const int64_t total = outer * n_dst * cpr;
    int64_t want = (total + 255) / 256;
    const int blocks = (int)(want < 1 ? 1 : (want > 148 * 32 ? 148 * 32 : want));
    switch (unit) {
        case 16: cm::gather_rows_kernel<uint4><<<blocks, 256, 0, s>>>((const uint4*)src, (uint4*)dst, perm, n_src, n_dst, (int)cpr, total); break;
        case 8: cm::gather_rows_kernel<uint2><<<blocks, 256, 0, s>>>((const uint2*)src, (uint2*)dst, perm, n_src, n_dst, (int)cpr, total); break;
        case 4: cm::gather_rows_kernel<uint32_t><<<blocks, 256, 0, s>>>((const uint32_t*)src, (uint32_t*)dst, perm, n_src, n_dst, (int)cpr, total); break;
        case 2: cm::gather_rows_kernel<uint16_t><<<blocks, 256, 0, s>>>((const uint16_t*)src, (uint16_t*)dst, perm, n_src, n_dst, (int)cpr, total); break;
        default: cm::gather_rows_kernel<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)src, (uint8_t*)dst, perm, n_src, n_dst, (int)cpr, total); break;
    }
    return (int)cudaGetLastError();
}
