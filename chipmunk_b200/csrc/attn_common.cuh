// Softmax steps of the column-sparse attention kernel (csp_attn.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace cm {
namespace attn {

constexpr int D = 128;              // head dim
constexpr int QG = 192;             // query rows per tile (= index group of the reference)
constexpr int KT = 128;             // key columns per step
constexpr float SCALE_LOG2 = 0.08838834764f * 1.44269504089f;   // log2(e)/sqrt(128), csp_attn.cu:265 of the reference
constexpr float RESCALE_THRESHOLD = 8.0f;                        // in log2 units

// One softmax step of one query row (= one thread): S row (128 fp32 in TMEM) -> P row (bf16, written
// over the first 64 columns of S).  m_ref is the row's reference maximum in raw score units; it is
// only moved (and O / l rescaled) when the new tile maximum exceeds it by more than 2^8 after
// scaling, so most steps skip the correction.  TAIL masks packed positions >= valid by position
// (reference csp_attn.cu:272) by loading them as -inf.
template <bool TAIL>
__device__ __forceinline__ void softmax_step(uint32_t tS, uint32_t tO, int valid, int kk, float& m_ref,
                                             float& l_sum, bool active = true) {
    uint32_t s[KT];
#pragma unroll
    for (int c = 0; c < KT; c += 32) tmem_ld32(tS + c, s + c);
    tmem_ld_wait();
    if (TAIL) {
#pragma unroll
        for (int j = 0; j < KT; j++) s[j] = j < valid ? s[j] : 0xff800000u;
    }
    // ---- tile max: 4 independent chains of 3-input max
    float mx[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        mx[c] = __uint_as_float(s[c * 32]);
#pragma unroll
        for (int j = 1; j < 31; j += 2)
            mx[c] = fmax3(mx[c], __uint_as_float(s[c * 32 + j]), __uint_as_float(s[c * 32 + j + 1]));
        mx[c] = fmaxf(mx[c], __uint_as_float(s[c * 32 + 31]));
    }
    const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    // ---- lazy rescale of the running state (always taken on the first step: m_ref = -inf)
    // Lanes that own no row -- lanes 16-31 of a block-1 warp (M = 64 accumulator, csp_attn.cu) -- run the same instruction
    // stream on whatever their TMEM lanes hold, but never vote.  (Branching them off the arithmetic was measured: the
    // divergence costs 8 % of the kernel, far more than the idle lanes' power.)
    const bool need = active && (m_tile - m_ref) * SCALE_LOG2 > RESCALE_THRESHOLD;
    if (__any_sync(0xffffffffu, need)) {
        float alpha = 1.f;
        if (need) {
            alpha = fast_exp2((m_ref - m_tile) * SCALE_LOG2);
            m_ref = m_tile;
            l_sum *= alpha;
        }
        if (kk > 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tO + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
                tmem_st_32x32b_x32(tO + c0, r);
            }
        }
    }
    // ---- P = exp2(s*c - m*c) (packed fp32x2 FMA), row sum, bf16 pack, write over S
    const float neg_m = -m_ref * SCALE_LOG2;
    const uint64_t c2 = pack_f32x2(SCALE_LOG2, SCALE_LOG2), nm2 = pack_f32x2(neg_m, neg_m);
    uint64_t acc[2] = {0ull, 0ull};
    const int cols = TAIL ? ((valid + 15) & ~15) : KT;
#pragma unroll
    for (int c0 = 0; c0 < KT; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const uint64_t x = ffma2(pack_f32x2(__uint_as_float(s[c0 + j]), __uint_as_float(s[c0 + j + 1])), c2, nm2);
            float x0, x1;
            unpack_f32x2(x, x0, x1);
            const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            acc[(j >> 1) & 1] = fadd2(acc[(j >> 1) & 1], pack_f32x2(p0, p1));
            pk[j >> 1] = pack_bf16x2(p0, p1);
        }
        if (!TAIL || c0 < cols) tmem_st_32x32b_x16(tS + (c0 >> 1), pk);
    }
    float a0, a1, a2, a3;
    unpack_f32x2(acc[0], a0, a1);
    unpack_f32x2(acc[1], a2, a3);
    l_sum += (a0 + a1) + (a2 + a3);
}


// A step with at most 32 valid key columns (the 16-column tail that a count = k*112 or k*128+16 list ends with):
// one 32-column chunk instead of four.  The step sits on the tile's critical chain like every other one, and at
// FLUX sizes (7 steps per tile) a full-width pass over a 16-column tail is ~4 % of the kernel.
__device__ __forceinline__ void softmax_step_narrow(uint32_t tS, uint32_t tO, int valid, int kk, float& m_ref, float& l_sum,
                                                    bool active = true) {
    uint32_t s[32];
    tmem_ld32(tS, s);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j++) s[j] = j < valid ? s[j] : 0xff800000u;
    float mx = __uint_as_float(s[0]);
#pragma unroll
    for (int j = 1; j < 31; j += 2) mx = fmax3(mx, __uint_as_float(s[j]), __uint_as_float(s[j + 1]));
    const float m_tile = fmaxf(mx, __uint_as_float(s[31]));
    const bool need = active && (m_tile - m_ref) * SCALE_LOG2 > RESCALE_THRESHOLD;
    if (__any_sync(0xffffffffu, need)) {
        float alpha = 1.f;
        if (need) {
            alpha = fast_exp2((m_ref - m_tile) * SCALE_LOG2);
            m_ref = m_tile;
            l_sum *= alpha;
        }
        if (kk > 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tO + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
                tmem_st_32x32b_x32(tO + c0, r);
            }
        }
    }
    const float neg_m = -m_ref * SCALE_LOG2;
    float acc = 0.f;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[j]), SCALE_LOG2, neg_m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[j + 1]), SCALE_LOG2, neg_m));
        acc += p0 + p1;
        pk[j >> 1] = pack_bf16x2(p0, p1);
    }
    tmem_st_32x32b_x16(tS, pk);
    l_sum += acc;
}

}  // namespace attn
}  // namespace cm
