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


// One softmax step of a warp that owns 16 query rows with FOUR threads per row (16-lane TMEM shapes, ptx.cuh): thread t
// holds, of rows t/4 ("A") and 8 + t/4 ("B") of the 16-lane window at tS / tO, the 32 score columns 8j + 2 (t%4) + {0,1}.
// A row's maximum is completed with two shuffles inside its four-thread group, so the four threads take identical
// decisions without any shared-memory exchange; row sums stay per-thread partials until the epilogue.  Every lane of
// the warp works: a 128-row block is eight such warps (two per TMEM lane quadrant), the M = 64 block four.
template <bool TAIL>
__device__ __forceinline__ void softmax_step16(uint32_t tS, uint32_t tO, int valid, int kk, float (&m_ref)[2], float (&l_sum)[2],
                                               int c4) {
    uint32_t s[64];
    tmem_ld_16x256b_x16(tS, s);
    tmem_ld_wait();
    if (TAIL) {
#pragma unroll
        for (int j = 0; j < 16; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const bool ok = 8 * j + 2 * c4 + e < valid;
                s[4 * j + e] = ok ? s[4 * j + e] : 0xff800000u;
                s[4 * j + 2 + e] = ok ? s[4 * j + 2 + e] : 0xff800000u;
            }
    }
    // ---- row maxima: two chains per row over this thread's 32 columns, then the four-thread group
    float mA0 = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1])), mA1 = fmaxf(__uint_as_float(s[4]), __uint_as_float(s[5]));
    float mB0 = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3])), mB1 = fmaxf(__uint_as_float(s[6]), __uint_as_float(s[7]));
#pragma unroll
    for (int j = 2; j < 16; j += 2) {
        mA0 = fmax3(mA0, __uint_as_float(s[4 * j]), __uint_as_float(s[4 * j + 1]));
        mB0 = fmax3(mB0, __uint_as_float(s[4 * j + 2]), __uint_as_float(s[4 * j + 3]));
        mA1 = fmax3(mA1, __uint_as_float(s[4 * j + 4]), __uint_as_float(s[4 * j + 5]));
        mB1 = fmax3(mB1, __uint_as_float(s[4 * j + 6]), __uint_as_float(s[4 * j + 7]));
    }
    float mA = fmaxf(mA0, mA1), mB = fmaxf(mB0, mB1);
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
    // ---- lazy rescale of the running state (always taken on the first step: m_ref = -inf)
    const bool needA = (mA - m_ref[0]) * SCALE_LOG2 > RESCALE_THRESHOLD;
    const bool needB = (mB - m_ref[1]) * SCALE_LOG2 > RESCALE_THRESHOLD;
    if (__any_sync(0xffffffffu, needA || needB)) {
        float alphaA = 1.f, alphaB = 1.f;
        if (needA) { alphaA = fast_exp2((m_ref[0] - mA) * SCALE_LOG2); m_ref[0] = mA; l_sum[0] *= alphaA; }
        if (needB) { alphaB = fast_exp2((m_ref[1] - mB) * SCALE_LOG2); m_ref[1] = mB; l_sum[1] *= alphaB; }
        if (kk > 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t r[16];
                tmem_ld_16x256b_x4(tO + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * alphaA);
                    r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * alphaA);
                    r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * alphaB);
                    r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * alphaB);
                }
                tmem_st_16x256b_x4(tO + c0, r);
            }
        }
    }
    // ---- P = exp2(s*c - m*c), row sums, bf16 pack; P word 4j + t%4 of rows A / B goes out as registers 2j / 2j+1 (16x128b)
    const float negA = -m_ref[0] * SCALE_LOG2, negB = -m_ref[1] * SCALE_LOG2;
    const uint64_t c2 = pack_f32x2(SCALE_LOG2, SCALE_LOG2), nA2 = pack_f32x2(negA, negA), nB2 = pack_f32x2(negB, negB);
    uint64_t accA = 0ull, accB = 0ull;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t pk[8];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const int j = 4 * q + jj;
            float x0, x1;
            unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(s[4 * j]), __uint_as_float(s[4 * j + 1])), c2, nA2), x0, x1);
            const float a0 = fast_exp2(x0), a1 = fast_exp2(x1);
            accA = fadd2(accA, pack_f32x2(a0, a1));
            pk[2 * jj] = pack_bf16x2(a0, a1);
            unpack_f32x2(ffma2(pack_f32x2(__uint_as_float(s[4 * j + 2]), __uint_as_float(s[4 * j + 3])), c2, nB2), x0, x1);
            const float b0 = fast_exp2(x0), b1 = fast_exp2(x1);
            accB = fadd2(accB, pack_f32x2(b0, b1));
            pk[2 * jj + 1] = pack_bf16x2(b0, b1);
        }
        tmem_st_16x128b_x4(tS + 16 * q, pk);
    }
    float a0, a1;
    unpack_f32x2(accA, a0, a1);
    l_sum[0] += a0 + a1;
    unpack_f32x2(accB, a0, a1);
    l_sum[1] += a0 + a1;
}

// softmax_step16 for a step with at most 32 valid key columns (see softmax_step_narrow): one 32-column piece, eight scores
// per thread and row.
__device__ __forceinline__ void softmax_step16_narrow(uint32_t tS, uint32_t tO, int valid, int kk, float (&m_ref)[2], float (&l_sum)[2],
                                                      int c4) {
    uint32_t s[16];
    tmem_ld_16x256b_x4(tS, s);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const bool ok = 8 * j + 2 * c4 + e < valid;
            s[4 * j + e] = ok ? s[4 * j + e] : 0xff800000u;
            s[4 * j + 2 + e] = ok ? s[4 * j + 2 + e] : 0xff800000u;
        }
    float mA = fmaxf(__uint_as_float(s[0]), __uint_as_float(s[1])), mB = fmaxf(__uint_as_float(s[2]), __uint_as_float(s[3]));
#pragma unroll
    for (int j = 1; j < 4; j++) {
        mA = fmax3(mA, __uint_as_float(s[4 * j]), __uint_as_float(s[4 * j + 1]));
        mB = fmax3(mB, __uint_as_float(s[4 * j + 2]), __uint_as_float(s[4 * j + 3]));
    }
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 1));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 1));
    mA = fmaxf(mA, __shfl_xor_sync(0xffffffffu, mA, 2));
    mB = fmaxf(mB, __shfl_xor_sync(0xffffffffu, mB, 2));
    const bool needA = (mA - m_ref[0]) * SCALE_LOG2 > RESCALE_THRESHOLD;
    const bool needB = (mB - m_ref[1]) * SCALE_LOG2 > RESCALE_THRESHOLD;
    if (__any_sync(0xffffffffu, needA || needB)) {
        float alphaA = 1.f, alphaB = 1.f;
        if (needA) { alphaA = fast_exp2((m_ref[0] - mA) * SCALE_LOG2); m_ref[0] = mA; l_sum[0] *= alphaA; }
        if (needB) { alphaB = fast_exp2((m_ref[1] - mB) * SCALE_LOG2); m_ref[1] = mB; l_sum[1] *= alphaB; }
        if (kk > 0) {
#pragma unroll 1
            for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t r[16];
                tmem_ld_16x256b_x4(tO + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) * alphaA);
                    r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) * alphaA);
                    r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) * alphaB);
                    r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) * alphaB);
                }
                tmem_st_16x256b_x4(tO + c0, r);
            }
        }
    }
    const float negA = -m_ref[0] * SCALE_LOG2, negB = -m_ref[1] * SCALE_LOG2;
    float accA = 0.f, accB = 0.f;
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const float a0 = fast_exp2(fmaf(__uint_as_float(s[4 * j]), SCALE_LOG2, negA)), a1 = fast_exp2(fmaf(__uint_as_float(s[4 * j + 1]), SCALE_LOG2, negA));
        const float b0 = fast_exp2(fmaf(__uint_as_float(s[4 * j + 2]), SCALE_LOG2, negB)), b1 = fast_exp2(fmaf(__uint_as_float(s[4 * j + 3]), SCALE_LOG2, negB));
        accA += a0 + a1;
        accB += b0 + b1;
        pk[2 * j] = pack_bf16x2(a0, a1);
        pk[2 * j + 1] = pack_bf16x2(b0, b1);
    }
    tmem_st_16x128b_x4(tS, pk);
    l_sum[0] += accA;
    l_sum[1] += accB;
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
