// Column-sparse delta attention, 64-key steps with DOUBLE-BUFFERED S (third schedule of `cm_csp_attn`).
//
// csp_attn.cu aliases P onto S and fills TMEM with S0,S1,O0,O1 (4 x 128 columns), so each query block's
// step is a serial chain  S -> softmax -> P.V -> next S  and the softmax (~1900 clk) is exposed.
// Here a step covers 64 keys: S tiles are 64 columns wide and TMEM holds TWO of them per block
//     S0a S0b S1a S1b (4 x 64)  +  O0 O1 (2 x 128)  = 512 columns,
// so the tensor pipe computes S(k+1), S(k+2) of a block while its softmax warps work on S(k): the
// softmax runs back to back and the chain is throughput- instead of latency-bound.
// Everything else (roles, gather layout, epilogue) follows csp_attn.cu.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/chipmunk_b200.h"
#include "attn_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace cm {
namespace attn3 {

using namespace cm::attn;

constexpr int KS = 64;                  // keys per step
constexpr int NSLOT = 10;               // 16 KB K/V slots
constexpr int SLOT_BYTES = KS * D * 2;  // 16384: K: 2 d-halves x [64 keys][128 B];  V: 2 d-halves x [64 keys][128 B]
constexpr int Q_HALF_BYTES = QG * 128;
constexpr int Q_BYTES = 2 * Q_HALF_BYTES;
constexpr int SMEM_BYTES = Q_BYTES + NSLOT * SLOT_BYTES + 1024;
constexpr int NUM_THREADS = 384;        // warps 0-3 softmax blk0 | 4-5 softmax blk1, 6 MMA, 7 idle | 8-11 producers
constexpr int WARP_MMA = 6, WARP_PROD0 = 8, NUM_PROD = 128;
// TMEM columns: S[blk][buf] at (blk*2 + buf) * 64, O[blk] at 256 + blk*128
constexpr uint32_t TM_O = 256;

struct Params {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;
    const int32_t* indices;
    const int32_t* counts;
    float* l;
    int B, H, Nq, Nk, G;
    int64_t qs[3], ks[3], vs[3], os[3];
    int64_t idx_row_stride;
    float o_scale;
    int accumulate;
    int num_tiles;
    int dense;
};

struct __align__(8) Barriers {
    uint64_t q_full, q_empty;
    uint64_t kv_full[NSLOT], kv_empty[NSLOT];
    uint64_t s_full[2][2];        // [block][buffer]
    uint64_t p_full[2][2];        // [block][buffer]: with S running ahead the softmax can be two steps ahead of the
                                  // issuer, so one barrier per S buffer (a single one would alias its phase parity)
    uint64_t pv_done[2][2];       // [block][step & 1]: P.V of a step has landed (needed to rescale O, and by the epilogue)
};

__device__ __forceinline__ int tile_count(const Params& P, int tile) {
    if (P.dense) return P.Nk;
    int c = __ldg(P.counts + tile);
    c = c < 0 ? 0 : c;
    return c > (int)P.idx_row_stride ? (int)P.idx_row_stride : c;
}

// One softmax step of one query row over a 64-column S tile; P (bf16) overwrites the first 32 columns.
template <bool TAIL>
__device__ __forceinline__ void softmax_step64(uint32_t tS, uint32_t tO, int valid, int kk, float& m_ref, float& l_sum,
                                               uint64_t* pv_done, uint32_t pv_parity) {
    uint32_t s[KS];
    tmem_ld32(tS, s);
    tmem_ld32(tS + 32, s + 32);
    tmem_ld_wait();
    if (TAIL) {
#pragma unroll
        for (int j = 0; j < KS; j++) s[j] = j < valid ? s[j] : 0xff800000u;
    }
    float mx[2];
#pragma unroll
    for (int c = 0; c < 2; c++) {
        mx[c] = __uint_as_float(s[c * 32]);
#pragma unroll
        for (int j = 1; j < 31; j += 2)
            mx[c] = fmax3(mx[c], __uint_as_float(s[c * 32 + j]), __uint_as_float(s[c * 32 + j + 1]));
        mx[c] = fmaxf(mx[c], __uint_as_float(s[c * 32 + 31]));
    }
    const float m_tile = fmaxf(mx[0], mx[1]);
    const bool need = (m_tile - m_ref) * SCALE_LOG2 > RESCALE_THRESHOLD;
    if (__any_sync(0xffffffffu, need)) {
        float alpha = 1.f;
        if (need) {
            alpha = fast_exp2((m_ref - m_tile) * SCALE_LOG2);
            m_ref = m_tile;
            l_sum *= alpha;
        }
        if (kk > 0) {
            mbar_wait(pv_done, pv_parity);      // S runs ahead of P.V: O is only stable once P.V(kk-1) has landed
            tc_fence_after_sync();
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
    const uint64_t c2 = pack_f32x2(SCALE_LOG2, SCALE_LOG2), nm2 = pack_f32x2(neg_m, neg_m);
    uint64_t acc[2] = {0ull, 0ull};
    const int cols = TAIL ? ((valid + 15) & ~15) : KS;
#pragma unroll
    for (int c0 = 0; c0 < KS; c0 += 32) {
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

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1) attn3_kernel(const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = sbase;
    const uint32_t sKV = sbase + Q_BYTES;

    if (tid == 0) {
        mbar_init(&bar.q_full, NUM_PROD);
        mbar_init(&bar.q_empty, 1);
        for (int i = 0; i < NSLOT; i++) { mbar_init(&bar.kv_full[i], NUM_PROD); mbar_init(&bar.kv_empty[i], 1); }
        for (int b = 0; b < 2; b++) {
            mbar_init(&bar.s_full[b][0], 1); mbar_init(&bar.s_full[b][1], 1);
            mbar_init(&bar.pv_done[b][0], 1); mbar_init(&bar.pv_done[b][1], 1);
        }
        mbar_init(&bar.p_full[0][0], 128); mbar_init(&bar.p_full[0][1], 128);
        mbar_init(&bar.p_full[1][0], 64); mbar_init(&bar.p_full[1][1], 64);
        fence_mbar_init();
    }
    if (warp == WARP_MMA) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;

    // =========================================================================== producers
    if (warp >= WARP_PROD0) {
        setmaxnreg_dec<80>();
        const int pt = tid - WARP_PROD0 * 32;       // 0..127
        const int chunk = pt & 15;                  // 16-byte chunk of the 256-byte row
        const int rsub = pt >> 4;                   // 0..7
        const uint32_t half_off = (uint32_t)(chunk >> 3);
        const uint32_t c8 = (uint32_t)(chunk & 7);
        uint32_t job = 0, it = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            const int count = tile_count(P, tile);
            if (count <= 0) continue;
            const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
            mbar_wait(&bar.q_empty, (it & 1) ^ 1);
            {
                const __nv_bfloat16* qb = P.q + b * P.qs[0] + h * P.qs[1];
#pragma unroll 4
                for (int i = 0; i < QG / 8; i++) {
                    const int r = rsub + 8 * i;
                    const int row = g * QG + r;
                    const bool ok = row < P.Nq;
                    cp_async_16_zfill(sQ + half_off * Q_HALF_BYTES + r * 128 + ((c8 ^ (r & 7)) << 4),
                                      qb + (int64_t)(ok ? row : 0) * P.qs[2] + chunk * 8, ok ? 16u : 0u);
                }
                cp_async_mbar_arrive_noinc(&bar.q_full);
            }
            it++;
            const __nv_bfloat16* kb = P.k + b * P.ks[0] + h * P.ks[1];
            const __nv_bfloat16* vb = P.v + b * P.vs[0] + h * P.vs[1];
            const int32_t* ip = P.dense ? nullptr : P.indices + (int64_t)tile * P.idx_row_stride;
            const int nk = (count + KS - 1) / KS;
            // lane j (j & 15 < 8) of a half-warp fetches the index of key row rsub + 8*(j & 7) of the step,
            // two steps ahead of its use; the 8 rows a thread copies are read back with shuffles
            const int my_r = rsub + 8 * (lane & 7);
            auto fetch_idx = [&](int kk) -> int {
                const int pos = kk * KS + my_r;
                int idx = pos;
                if (!P.dense) idx = pos < count ? __ldg(ip + pos) : 0;
                idx = idx < 0 ? 0 : (idx >= P.Nk ? P.Nk - 1 : idx);
                return pos < count ? idx : -1;
            };
            int idx_a = fetch_idx(0), idx_b = nk > 1 ? fetch_idx(1) : -1;
            for (int kk = 0; kk < nk; kk++) {
                const int idx_cur = idx_a;
                idx_a = idx_b;
                if (kk + 2 < nk) idx_b = fetch_idx(kk + 2);
                int rowidx[KS / 8];
#pragma unroll
                for (int i = 0; i < KS / 8; i++) rowidx[i] = __shfl_sync(0xffffffffu, idx_cur, (lane & 16) + i);
#pragma unroll
                for (int op = 0; op < 2; op++) {
                    const uint32_t slot = job % NSLOT;
                    mbar_wait(&bar.kv_empty[slot], ((job / NSLOT) & 1) ^ 1);
                    const uint32_t dst0 = sKV + slot * SLOT_BYTES + half_off * (SLOT_BYTES / 2);
                    const __nv_bfloat16* base = (op == 0 ? kb : vb) + chunk * 8;
                    const int64_t rs = op == 0 ? P.ks[2] : P.vs[2];
#pragma unroll
                    for (int i = 0; i < KS / 8; i++) {
                        const int r = rsub + 8 * i;
                        const bool ok = rowidx[i] >= 0;
                        cp_async_16_zfill(dst0 + r * 128 + ((c8 ^ (r & 7)) << 4), base + (int64_t)(ok ? rowidx[i] : 0) * rs, ok ? 16u : 0u);
                    }
                    cp_async_mbar_arrive_noinc(&bar.kv_full[slot]);
                    job++;
                }
            }
        }
        cp_async_wait_all();
    }
    // =========================================================================== MMA issuer
    else if (warp == WARP_MMA) {
        setmaxnreg_inc<208>();
        uint32_t jobbase = 0, it = 0, gs = 0;       // slot jobs before this tile, tiles, global step counter
        const uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);
        const uint64_t desc_q = umma_smem_desc(sQ, 16, 1024);
        const uint64_t desc_k = umma_smem_desc(sKV, 16, 1024);
        const uint64_t desc_v = umma_smem_desc(sKV, SLOT_BYTES / 2, 1024);
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            const int count = tile_count(P, tile);
            if (count <= 0) continue;
            const int nk = (count + KS - 1) / KS;
            auto ncols = [&](int kk) { int v = count - kk * KS; v = v > KS ? KS : v; return (v + 15) & ~15; };
            auto wait_slot = [&](uint32_t job) {
                const uint32_t slot = job % NSLOT;
                mbar_wait(&bar.kv_full[slot], (job / NSLOT) & 1);
                return slot;
            };
            // S[blk](kk) = Q_blk K(kk)^T into buffer (gs+kk)&1
            auto issue_S = [&](int blk, int kk, uint32_t slot) {
                const uint32_t idesc = umma_idesc_bf16(128, ncols(kk), 0, 0);
                const uint32_t buf = (gs + kk) & 1;
                const uint32_t d = tm + (blk * 2 + buf) * KS;
                const uint64_t ad0 = desc_q + (uint64_t)(blk * ((128 * 128) >> 4));
                const uint64_t bd0 = desc_k + (uint64_t)(slot * (SLOT_BYTES >> 4));
#pragma unroll
                for (int k16 = 0; k16 < D / 16; k16++) {
                    const uint64_t ad = ad0 + (uint64_t)((((k16 >> 2) * Q_HALF_BYTES) + (k16 & 3) * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)((((k16 >> 2) * (SLOT_BYTES / 2)) + (k16 & 3) * 32) >> 4);
                    umma_ss(d, ad, bd, idesc, k16 > 0);
                }
                umma_commit(&bar.s_full[blk][buf]);
            };
            auto issue_PV = [&](int blk, int kk, uint32_t slot) {
                const uint32_t buf = (gs + kk) & 1;
                const uint32_t d = tm + TM_O + blk * D;
                const uint32_t a = tm + (blk * 2 + buf) * KS;
                const uint64_t bd0 = desc_v + (uint64_t)(slot * (SLOT_BYTES >> 4));
                const int steps = ncols(kk) / 16;
                for (int j = 0; j < steps; j++) umma_ts(d, a + j * 8, bd0 + (uint64_t)(j * (2048 >> 4)), idesc_pv, (kk | j) != 0);
                umma_commit(&bar.pv_done[blk][(gs + kk) & 1]);
            };
            mbar_wait(&bar.q_full, it & 1);
            // prologue: S(0), S(1) of both blocks
            for (int kk = 0; kk < 2 && kk < nk; kk++) {
                const uint32_t slot = wait_slot(jobbase + 2 * kk);
                tc_fence_after_sync();
                if (lane == 0) {
                    issue_S(0, kk, slot);
                    issue_S(1, kk, slot);
                    umma_commit(&bar.kv_empty[slot]);
                    if (kk == nk - 1) umma_commit(&bar.q_empty);
                }
                __syncwarp();
            }
            for (int kk = 0; kk < nk; kk++) {
                const uint32_t slotV = wait_slot(jobbase + 2 * kk + 1);
                const bool more = kk + 2 < nk;
                uint32_t slotKn = 0;
                if (more) slotKn = wait_slot(jobbase + 2 * (kk + 2));
                // block 0
                mbar_wait(&bar.p_full[0][(gs + kk) & 1], ((gs + kk) >> 1) & 1);
                tc_fence_after_sync();
                if (lane == 0) {
                    issue_PV(0, kk, slotV);
                    if (more) issue_S(0, kk + 2, slotKn);
                }
                __syncwarp();
                // block 1
                mbar_wait(&bar.p_full[1][(gs + kk) & 1], ((gs + kk) >> 1) & 1);
                tc_fence_after_sync();
                if (lane == 0) {
                    issue_PV(1, kk, slotV);
                    umma_commit(&bar.kv_empty[slotV]);
                    if (more) {
                        issue_S(1, kk + 2, slotKn);
                        umma_commit(&bar.kv_empty[slotKn]);
                        if (kk + 2 == nk - 1) umma_commit(&bar.q_empty);
                    }
                }
                __syncwarp();
            }
            gs += nk;
            jobbase += 2 * nk;
            it++;
        }
    }
    // =========================================================================== softmax + epilogue
    else if (warp < 6) {
        setmaxnreg_inc<208>();
        const int blk = warp >> 2;
        const int r_in_tile = blk * 128 + (warp & 3) * 32 + lane;
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tO = tm + TM_O + blk * D + lane_off;
        uint32_t gs = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            const int count = tile_count(P, tile);
            const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
            const int row = g * QG + r_in_tile;
            const bool row_ok = row < P.Nq;
            __nv_bfloat16* orow = P.o + b * P.os[0] + h * P.os[1] + (int64_t)(row_ok ? row : 0) * P.os[2];
            if (count <= 0) {
                if (!P.accumulate && row_ok) {
#pragma unroll
                    for (int c = 0; c < 16; c++) reinterpret_cast<uint4*>(orow)[c] = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            const int nk = (count + KS - 1) / KS;
            float m_ref = -INFINITY, l_sum = 0.f;
            for (int kk = 0; kk < nk; kk++, gs++) {
                const int valid = min(KS, count - kk * KS);
                const uint32_t buf = gs & 1;
                mbar_wait(&bar.s_full[blk][buf], (gs >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t tS = tm + (blk * 2 + buf) * KS + lane_off;
                uint64_t* pvb = &bar.pv_done[blk][(gs - 1) & 1];
                const uint32_t pvp = ((gs - 1) >> 1) & 1;
                if (valid == KS) softmax_step64<false>(tS, tO, KS, kk, m_ref, l_sum, pvb, pvp);
                else softmax_step64<true>(tS, tO, valid, kk, m_ref, l_sum, pvb, pvp);
                tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(&bar.p_full[blk][buf]);
            }
            // ---- epilogue
            mbar_wait(&bar.pv_done[blk][(gs - 1) & 1], ((gs - 1) >> 1) & 1);
            tc_fence_after_sync();
            const float inv = P.o_scale / l_sum;
            if (P.dense && row_ok && P.l) P.l[(int64_t)bh * P.Nq + row] = 1.f / (fast_exp2(m_ref * SCALE_LOG2) * l_sum);
            for (int c0 = 0; c0 < D; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tO + c0, r);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            w[j] = pack_bf16x2(__uint_as_float(r[q4 * 8 + 2 * j]) * inv, __uint_as_float(r[q4 * 8 + 2 * j + 1]) * inv);
                        uint4* dst = reinterpret_cast<uint4*>(orow + c0 + q4 * 8);
                        if (P.accumulate) red_add_bf16x8(dst, w[0], w[1], w[2], w[3]);
                        else *dst = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            tc_fence_before_sync();
        }
    } else {
        setmaxnreg_inc<208>();      // warp 7: idle
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tm, 512);
}

int launch(const void* q, const void* k, const void* v, void* o, float* l, const int32_t* indices, const int32_t* counts,
           int B, int H, int Nq, int Nk, const int64_t qs[3], const int64_t ks[3], const int64_t vs[3], const int64_t os[3],
           int64_t idx_row_stride, int o_scale, int accumulate, int dense, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (int rc0 = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(attn3_kernel), SMEM_BYTES)) return rc0;
    Params P{};
    P.q = (const __nv_bfloat16*)q; P.k = (const __nv_bfloat16*)k; P.v = (const __nv_bfloat16*)v; P.o = (__nv_bfloat16*)o;
    P.l = l; P.indices = indices; P.counts = counts;
    P.B = B; P.H = H; P.Nq = Nq; P.Nk = Nk; P.G = (Nq + QG - 1) / QG;
    for (int i = 0; i < 3; i++) { P.qs[i] = qs[i]; P.ks[i] = ks[i]; P.vs[i] = vs[i]; P.os[i] = os[i]; }
    P.idx_row_stride = idx_row_stride;
    P.o_scale = (float)o_scale;
    P.accumulate = accumulate ? 1 : 0;
    P.dense = dense;
    const int64_t tiles = (int64_t)B * H * P.G;
    if (tiles > 2147483647ll) return CM_EINVAL;
    P.num_tiles = (int)tiles;
    const int grid = P.num_tiles < sm_count() ? P.num_tiles : sm_count();
    attn3_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(P);
    return (int)cudaGetLastError();
}

}  // namespace attn3
}  // namespace cm
