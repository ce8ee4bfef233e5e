// Dense attention with the statistics the sparse steps need, in ONE pass: o, l and (optionally) the per-group
// column sums cs.  Replaces csrc/attn/dense_attn.cu and csrc/attn/dense_colsum_attn.cu of the reference (Hopper
// wgmma kernels that compute o, l -- and cs -- in one flash-attention loop, dense_colsum_attn.cu:205-341).
//
// Work unit: one (b, h, 128 query rows) tile, walked over all keys 128 at a time.  All eight softmax warps work on
// that one tile, two threads per query row (64 keys each).  P has its own (triple-buffered) TMEM columns, so the next
// S = Q K^T is issued as soon as the softmax threads have READ the current one -- before they have computed anything --
// and the P.V products trail behind: the tensor pipe never waits for a softmax step and a softmax step never waits
// for the tensor pipe (the 192/256-row kernels of csp_attn.cu keep P inside S and are bound by each block's serial
// softmax -> P.V -> S chain).  The two threads of a row share nothing per step: each reads the whole 128-wide S row
// for the maximum (TMEM reads are cheap) and exponentiates its own half, so the two warps of an SM sub-partition
// drift apart and one's MUFU phase overlaps the other's loads / maxima / conversions.
//
//   warp 9       TMA producer: Q tile once per tile; K / V tiles (128 rows x 256 B, two 64-wide halves, 128B-swizzled)
//                through a 4-slot ring.  4-D tensor maps: any batch / head / row stride, rows past N read as zero.
//   warp 8       MMA issuer (one elected thread), per step k:
//                  CS(k-1) = P(k-1)^T F(k-1)   SS, M=128 keys, N=16, K=128 queries -> TMEM CS   (column sums, see below)
//                  S(k+1)  = Q K(k+1)^T        SS, M=128 N=128 K=128  -> TMEM S        (after s_free(k))
//                  O      += P(k-1) V(k-1)     TS (P from its TMEM buffer), V as MN-major smem   (after p_full(k-1))
//   warps 0-7    softmax: thread (row r, half h): S row -> maximum (whole row) -> lazy rescale -> exp2 of keys
//                [64h, 64h+64) -> bf16 P into TMEM buffer k%3 (and, for the column sums, into shared memory).
//                Epilogue (same threads): O / l -> bf16 -> 128B-swizzled staging tile -> TMA store; l to global.
//
// Column sums.  cs[b,h,g,j] = sum_{i in 192-row group g} exp(s_ij / sqrt(d)) p_i   (dense_colsum_attn.cu:267-277).
// With P_ij = exp2(s_ij c - m_i c) in hand (m_i = the row's running reference maximum), that is
//     cs[g, j] = sum_i P_ij f_i ,   f_i = exp2(m_i c + log2 p_i)   over the rows i of the tile that lie in group g:
// a [128 keys x 128 queries] x [128 queries x 16] product per step, i.e. the TENSOR PIPE does the cross-row reduction
// at 1/8 of the cost of S (A = P^T: the P tile in shared memory read MN-major; B = F: row 0 / row 1 hold the f_i of the
// tile's first / second group, rows 2-15 zero).  Key j of the step lands on TMEM lane j, so the softmax threads of
// half 0 pick their key's two sums up at the next step and add them into cs with a bf16 reduction at the L2.
// A 128-row tile overlaps at most two 192-row groups and every group is covered by two tiles: each cs element
// receives exactly two partial sums (fp32-accumulated over up to 128 rows each, then bf16).  The reference reduces
// twelve warps' bf16 partials with shared-memory atomics.
//
// TMEM (512 columns): S [0,128)  P0/P1/P2 [128,320)  O [320,448)  CS [448,464).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace cm {
namespace dense {

constexpr int D = 128, BM = 128, KT = 128, QG = 192;
constexpr float SCALE_LOG2 = 0.08838834764f * 1.44269504089f;   // log2(e)/sqrt(128) (reference csp_attn.cu:265)
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr int NSLOT = 4;
constexpr int TILE_BYTES = 128 * 256;          // 32 KB: Q tile, K/V slot, P tile, output staging
constexpr int HALF_BYTES = TILE_BYTES / 2;     // one 64-wide half: 128 rows x 128 B
constexpr int F_BYTES = 4096;                  // B operand of the column-sum MMA: two 64-query halves x (16 rows x 128 B)
constexpr int SMEM_BYTES = TILE_BYTES /*Q*/ + TILE_BYTES /*P / staging*/ + F_BYTES + NSLOT * TILE_BYTES + 1024 /*align*/;
constexpr int NUM_THREADS = 384;               // warps 0-7 softmax | 8 MMA | 9 TMA | 10-11 idle
constexpr int WARP_MMA = 8, WARP_TMA = 9;
constexpr uint32_t TM_S = 0, TM_P = 128, TM_O = 320, TM_CS = 448;

struct Params {
    float* l;                  // [B,H,Nq] or null
    const float* p;            // [B,H,Nq] previous step's l (column sums only)
    __nv_bfloat16* cs;         // [B*H*G, cs_stride], zero-initialised by the launcher
    int64_t cs_stride;
    int B, H, Nq, Nk, G, tiles_per_head, num_tiles, nk;
    // position (1..3) of the row / head / batch coordinate in each tensor map (q, k, v, o): the maps order their outer
    // dimensions by ascending stride, whatever view the caller passes
    int8_t pos[4][3];
};

struct __align__(8) Barriers {
    uint64_t q_full, q_empty;
    uint64_t kv_full[NSLOT], kv_empty[NSLOT];
    uint64_t s_full, s_free;
    uint64_t p_full[2];        // by step parity: the MMA thread may lag the softmax threads by more than one step
    uint64_t pv_done[3];       // one per P buffer
    uint64_t cs_full;
};

// coordinates (row, head, batch) placed at the positions the tensor map wants them
struct Coord { int c[4]; };
__device__ __forceinline__ Coord coords(const int8_t pos[3], int col, int row, int h, int b) {
    Coord r;
    r.c[0] = col;
#pragma unroll
    for (int i = 1; i < 4; i++) r.c[i] = pos[0] == i ? row : (pos[1] == i ? h : b);
    return r;
}

template <bool HAS_CS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dense_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
             const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_lx[2 * BM];         // partial row sums, exchanged in the epilogue

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = sbase, sP = sbase + TILE_BYTES, sF = sbase + 2 * TILE_BYTES, sKV = sF + F_BYTES;

    if (tid == 0) {
        mbar_init(&bar.q_full, 1); mbar_init(&bar.q_empty, 1);
        for (int i = 0; i < NSLOT; i++) { mbar_init(&bar.kv_full[i], 1); mbar_init(&bar.kv_empty[i], 1); }
        mbar_init(&bar.s_full, 1); mbar_init(&bar.s_free, 256);
        mbar_init(&bar.p_full[0], 256); mbar_init(&bar.p_full[1], 256);
        for (int i = 0; i < 3; i++) mbar_init(&bar.pv_done[i], 1);
        mbar_init(&bar.cs_full, 1);
        fence_mbar_init();
    }
    if (warp == WARP_MMA) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    }
    if (HAS_CS) {
        // rows 2-15 of F stay zero for the whole kernel
        for (int i = tid; i < F_BYTES / 4; i += NUM_THREADS)
            asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(sF + 4 * i), "r"(0u) : "memory");
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const int nk = P.nk;

    if (warp == WARP_TMA) {
        // =================================================================================== TMA producer
        if (lane == 0) {
            uint32_t job = 0, it = 0;
            auto load_kv = [&](const CUtensorMap* map, const int8_t* pos, int kstep, int h, int b) {
                const uint32_t slot = job % NSLOT;
                mbar_wait(&bar.kv_empty[slot], ((job / NSLOT) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar.kv_full[slot], TILE_BYTES);
                const Coord c = coords(pos, 0, kstep * KT, h, b);
                tma_load_4d(sKV + slot * TILE_BYTES, map, &bar.kv_full[slot], 0, c.c[1], c.c[2], c.c[3]);
                tma_load_4d(sKV + slot * TILE_BYTES + HALF_BYTES, map, &bar.kv_full[slot], 64, c.c[1], c.c[2], c.c[3]);
                job++;
            };
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, it++) {
                const int t = tile % P.tiles_per_head, bh = tile / P.tiles_per_head, h = bh % P.H, b = bh / P.H;
                mbar_wait(&bar.q_empty, (it & 1) ^ 1);
                mbar_arrive_expect_tx(&bar.q_full, TILE_BYTES);
                const Coord cq = coords(P.pos[0], 0, t * BM, h, b);
                tma_load_4d(sQ, &tm_q, &bar.q_full, 0, cq.c[1], cq.c[2], cq.c[3]);
                tma_load_4d(sQ + HALF_BYTES, &tm_q, &bar.q_full, 64, cq.c[1], cq.c[2], cq.c[3]);
                // consumption order of the ring: K0, then per step k: K(k+1), V(k-1); finally V(nk-1)
                load_kv(&tm_k, P.pos[1], 0, h, b);
                for (int k = 0; k < nk; k++) {
                    if (k + 1 < nk) load_kv(&tm_k, P.pos[1], k + 1, h, b);
                    if (k > 0) load_kv(&tm_v, P.pos[2], k - 1, h, b);
                }
                load_kv(&tm_v, P.pos[2], nk - 1, h, b);
            }
        }
    } else if (warp == WARP_MMA) {
        // =================================================================================== MMA issuer
        uint32_t job = 0, it = 0, g = 0;       // ring jobs, tiles, global step index of the tile's first step
        const uint32_t idesc_s = umma_idesc_bf16(128, KT, 0, 0);
        const uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);
        const uint32_t idesc_cs = umma_idesc_bf16(128, 16, 1, 0);
        const uint64_t desc_q = umma_smem_desc(sQ, 16, 1024);                 // K-major A: Q rows
        const uint64_t desc_k = umma_smem_desc(sKV, 16, 1024);                // K-major B: K rows
        const uint64_t desc_v = umma_smem_desc(sKV, HALF_BYTES, 1024);        // MN-major B: V rows (k = key, n = head dim)
        const uint64_t desc_pt = umma_smem_desc(sP, HALF_BYTES, 1024);        // MN-major A: P rows (k = query, m = key)
        const uint64_t desc_f = umma_smem_desc(sF, 16, 1024);                 // K-major B: F rows (n = group row, k = query)
        auto issue_S = [&](uint32_t slot) {
            const uint64_t bd0 = desc_k + (uint64_t)(slot * (TILE_BYTES >> 4));
#pragma unroll
            for (int k16 = 0; k16 < D / 16; k16++) {
                const uint64_t off = (uint64_t)((((k16 >> 2) * HALF_BYTES) + (k16 & 3) * 32) >> 4);
                umma_ss(tm + TM_S, desc_q + off, bd0 + off, idesc_s, k16 > 0);
            }
        };
        auto issue_PV = [&](uint32_t pbuf, uint32_t slot, bool first) {
            const uint64_t bd0 = desc_v + (uint64_t)(slot * (TILE_BYTES >> 4));
#pragma unroll
            for (int j = 0; j < KT / 16; j++)
                umma_ts(tm + TM_O, tm + TM_P + pbuf * 64 + j * 8, bd0 + (uint64_t)(j * (2048 >> 4)), idesc_pv, (!first) || j > 0);
        };
        auto issue_CS = [&]() {
#pragma unroll
            for (int j = 0; j < BM / 16; j++) {       // 16 queries per MMA
                const uint64_t boff = (uint64_t)((((j >> 2) * (F_BYTES / 2)) + (j & 3) * 32) >> 4);
                umma_ss(tm + TM_CS, desc_pt + (uint64_t)(j * (2048 >> 4)), desc_f + boff, idesc_cs, j > 0);
            }
        };
        // step j (global index): softmax has written P(j) -> CS(j), P(j).V(j)
        auto wait_p = [&](uint32_t j) { mbar_wait(&bar.p_full[j & 1], (j >> 1) & 1); };
        auto do_pv = [&](uint32_t j, bool first) {
            const uint32_t sv = job % NSLOT;
            mbar_wait(&bar.kv_full[sv], (job / NSLOT) & 1);
            job++;
            tc_fence_after_sync();
            if (elect_one()) {
                issue_PV(j % 3, sv, first);
                umma_commit(&bar.kv_empty[sv]);
                umma_commit(&bar.pv_done[j % 3]);
            }
            __syncwarp();
        };
        auto do_cs = [&]() {
            tc_fence_after_sync();
            if (elect_one()) { issue_CS(); umma_commit(&bar.cs_full); }
            __syncwarp();
        };
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, it++) {
            mbar_wait(&bar.q_full, it & 1);
            {   // S(0)
                const uint32_t s0 = job % NSLOT;
                mbar_wait(&bar.kv_full[s0], (job / NSLOT) & 1);
                job++;
                tc_fence_after_sync();
                if (elect_one()) {
                    issue_S(s0); umma_commit(&bar.s_full); umma_commit(&bar.kv_empty[s0]);
                    if (nk == 1) umma_commit(&bar.q_empty);
                }
                __syncwarp();
            }
            for (int k = 0; k < nk; k++) {
                if (k > 0) {
                    wait_p(g + k - 1);
#ifndef CM_CS_LATE
                    if (HAS_CS) do_cs();
#endif
                }
                if (k + 1 < nk) {
                    mbar_wait(&bar.s_free, (g + k) & 1);            // every softmax thread has read S(k)
                    const uint32_t sk = job % NSLOT;
                    mbar_wait(&bar.kv_full[sk], (job / NSLOT) & 1);
                    job++;
                    tc_fence_after_sync();
                    if (elect_one()) {
                        issue_S(sk); umma_commit(&bar.s_full); umma_commit(&bar.kv_empty[sk]);
                        if (k + 2 == nk) umma_commit(&bar.q_empty);
                    }
                    __syncwarp();
                }
                if (k > 0) do_pv(g + k - 1, k == 1);
#ifdef CM_CS_LATE
                if (HAS_CS && k > 0) do_cs();
#endif
            }
            wait_p(g + nk - 1);
            if (HAS_CS) do_cs();
            do_pv(g + nk - 1, nk == 1);
            g += nk;
        }
    } else if (warp < 8) {
        // =================================================================================== softmax + epilogue
        const int q4 = warp & 3, hf = warp >> 2;
        const int r_in_tile = q4 * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
        const uint32_t tS = tm + TM_S + lane_off;
        const uint32_t tO = tm + TM_O + lane_off + hf * 64;
        const uint32_t bar_id = 1 + q4;
        // MUFU turn-taking between the two warps of an SM sub-partition (they share its 4-lane MUFU): warp (q4, 0) and
        // warp (q4, 1) exponentiate alternately, so one's loads / maxima / conversions / stores run under the other's
        // exp2 phase instead of both phases colliding (FA3-style ping-pong with 64-thread named barriers)
        const uint32_t tok_mine = 7 + 2 * q4 + hf, tok_other = 7 + 2 * q4 + (hf ^ 1);
        if (hf == 1) named_bar_arrive(tok_other, 64);        // half 0 goes first
        uint32_t g = 0;                    // global step index (phases of s_full / s_free / p_full / pv_done / cs_full)
        const uint32_t sw = (uint32_t)(r_in_tile & 7);
        const uint64_t c2 = pack_f32x2(SCALE_LOG2, SCALE_LOG2);

        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            const int t = tile % P.tiles_per_head, bh = tile / P.tiles_per_head, h = bh % P.H, b = bh / P.H;
            const int row = t * BM + r_in_tile;
            const bool row_ok = row < P.Nq;
            // the previous tile's output box has been read out of the staging tile (= the P tile)
            if (tid == 0) bulk_wait_read<0>();
            named_bar_sync(5, 256);
            float lp = -INFINITY;          // log2 of the previous step's l of this row (column sums)
            uint32_t f_addr = 0, f_other = 0;
            // column sums: this thread drains key (32 q4 + lane) of every step into the tile's first / second group
            const int gA = (t * BM) / QG;
            const bool has_b = t * BM + BM - 1 >= QG * (gA + 1) && gA + 1 < P.G;
            __nv_bfloat16* cs_a = nullptr;
            __nv_bfloat16* cs_b = nullptr;
            if (HAS_CS && hf == 0) {
                const float pv = row_ok ? __ldg(P.p + (int64_t)bh * P.Nq + row) : 0.f;
                lp = pv > 0.f ? __log2f(pv) : -INFINITY;
                const uint32_t r01 = row >= QG * (gA + 1) ? 1u : 0u;      // F row of this query's group
                const uint32_t base = sF + (uint32_t)(r_in_tile >> 6) * (F_BYTES / 2) + (uint32_t)(r_in_tile & 7) * 2;
                const uint32_t chunk = (uint32_t)((r_in_tile & 63) >> 3);
                f_addr = base + r01 * 128 + ((chunk ^ r01) << 4);
                f_other = base + (r01 ^ 1u) * 128 + ((chunk ^ (r01 ^ 1u)) << 4);
                cs_a = P.cs + ((int64_t)bh * P.G + gA) * P.cs_stride + r_in_tile;      // (drain_cs rebases to the 8-key chunk)
                cs_b = cs_a + P.cs_stride;
            }
            auto drain_cs = [&](int kstep) {       // column sums of step `kstep` (cs_full already waited): TMEM lane = key
                uint32_t c0, c1;
                tmem_ld_32x32b_x2(tm + TM_CS + lane_off, c0, c1);
                tmem_ld_wait();
                // lanes 8j .. 8j+7 hold 8 consecutive keys: gather them into lane 8j (group A) / lane 8j+4 (group B) as four
                // bf16 pairs, so that the sums leave as 16-byte reductions (a scalar bf16 red per key is 8x the L2 atomics)
                const uint32_t nb = __shfl_down_sync(0xffffffffu, c0, 1), nb1 = __shfl_down_sync(0xffffffffu, c1, 1);
                const uint32_t pa = pack_bf16x2(__uint_as_float(c0), __uint_as_float(nb));     // valid on even lanes
                const uint32_t pb = pack_bf16x2(__uint_as_float(c1), __uint_as_float(nb1));
                const int base = lane & ~7;
                const bool for_b = (lane & 4) != 0;
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t va = __shfl_sync(0xffffffffu, pa, base + 2 * j);
                    const uint32_t vb = __shfl_sync(0xffffffffu, pb, base + 2 * j);
                    w[j] = for_b ? vb : va;
                }
                const int key = kstep * KT + q4 * 32 + base;         // first of the 8 keys
                if ((lane & 3) == 0 && key < P.cs_stride) {
                    __nv_bfloat16* dst = (for_b ? cs_b : cs_a) - r_in_tile + key;
                    if (!for_b || has_b) red_add_bf16x8(dst, w[0], w[1], w[2], w[3]);
                }
            };
            float m_ref = -INFINITY, l_sum = 0.f;

            for (int k = 0; k < nk; k++, g++) {
                const int valid = P.Nk - k * KT;                 // >= 128 except on the last step
                mbar_wait(&bar.s_full, g & 1);
                tc_fence_after_sync();
                uint32_t s[64];
                float m_tile;
                {
                    uint32_t o[64];                              // the partner's half: only its maximum is needed
                    tmem_ld32(tS + (hf ^ 1) * 64, o);
                    tmem_ld32(tS + (hf ^ 1) * 64 + 32, o + 32);
                    tmem_ld32(tS + hf * 64, s);
                    tmem_ld32(tS + hf * 64 + 32, s + 32);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    mbar_arrive(&bar.s_free);                    // S may be overwritten by the next Q K^T
                    if (valid < KT) {
#pragma unroll
                        for (int j = 0; j < 64; j++) {
                            s[j] = (hf * 64 + j < valid) ? s[j] : 0xff800000u;
                            o[j] = ((hf ^ 1) * 64 + j < valid) ? o[j] : 0xff800000u;
                        }
                    }
                    float mx[4] = {__uint_as_float(s[0]), __uint_as_float(s[32]), __uint_as_float(o[0]), __uint_as_float(o[32])};
#pragma unroll
                    for (int j = 1; j < 31; j += 2) {
                        mx[0] = fmax3(mx[0], __uint_as_float(s[j]), __uint_as_float(s[j + 1]));
                        mx[1] = fmax3(mx[1], __uint_as_float(s[32 + j]), __uint_as_float(s[32 + j + 1]));
                        mx[2] = fmax3(mx[2], __uint_as_float(o[j]), __uint_as_float(o[j + 1]));
                        mx[3] = fmax3(mx[3], __uint_as_float(o[32 + j]), __uint_as_float(o[32 + j + 1]));
                    }
                    m_tile = fmaxf(fmaxf(fmax3(mx[0], mx[1], __uint_as_float(s[31])), __uint_as_float(s[63])),
                                   fmaxf(fmax3(mx[2], mx[3], __uint_as_float(o[31])), __uint_as_float(o[63])));
                }
                // both threads of the row see the same S row and take the same decision
                const bool need = (m_tile - m_ref) * SCALE_LOG2 > RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, need)) {
                    float alpha = 1.f;
                    if (need) {
                        alpha = fast_exp2((m_ref - m_tile) * SCALE_LOG2);
                        m_ref = m_tile;
                        l_sum *= alpha;
                    }
                    if (k > 0) {
                        mbar_wait(&bar.pv_done[(g - 1) % 3], ((g - 1) / 3) & 1);     // O += P(k-1) V(k-1) has landed
                        tc_fence_after_sync();
#pragma unroll 1
                        for (int c0 = 0; c0 < 64; c0 += 32) {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(tO + c0, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__uint_as_float(r[j]) * alpha);
                            tmem_st_32x32b_x32(tO + c0, r);
                        }
                    }
                }
                // P buffer g % 3 was last read by P(g-3).V(g-3)
                if (g >= 3) mbar_wait(&bar.pv_done[g % 3], ((g - 3) / 3) & 1);
                const uint32_t tP = tm + TM_P + (g % 3) * 64 + lane_off + hf * 32;
                const float neg_m = -m_ref * SCALE_LOG2;
                const uint64_t nm2 = pack_f32x2(neg_m, neg_m);
                uint64_t acc[2] = {0ull, 0ull};
                uint32_t pk[32];
                named_bar_sync(tok_mine, 64);                        // my turn on the MUFU
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint64_t x = ffma2(pack_f32x2(__uint_as_float(s[c0 + j]), __uint_as_float(s[c0 + j + 1])), c2, nm2);
                        float x0, x1;
                        unpack_f32x2(x, x0, x1);
                        const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
                        acc[(j >> 1) & 1] = fadd2(acc[(j >> 1) & 1], pack_f32x2(p0, p1));
                        pk[(c0 + j) >> 1] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(tP + (c0 >> 1), *reinterpret_cast<uint32_t(*)[16]>(pk + (c0 >> 1)));
                }
                named_bar_arrive(tok_other, 64);                     // the partner warp's turn
                float a0, a1, a2, a3;
                unpack_f32x2(acc[0], a0, a1);
                unpack_f32x2(acc[1], a2, a3);
                l_sum += (a0 + a1) + (a2 + a3);
                if (HAS_CS) {
                    if (k > 0) {
                        // the previous step's column-sum MMA is done: its sums can be collected, and the P tile / F rewritten
                        mbar_wait(&bar.cs_full, (g - 1) & 1);
                        tc_fence_after_sync();
                        if (hf == 0) drain_cs(k - 1);
                    }
                    const uint32_t prow = sP + hf * HALF_BYTES + r_in_tile * 128;
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        st_shared_v4(prow + (((uint32_t)q ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    if (hf == 0) {
                        const float f = fast_exp2(m_ref * SCALE_LOG2 + lp);
                        const unsigned short fb = __bfloat16_as_ushort(__float2bfloat16(f));
                        asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(f_addr), "h"(fb) : "memory");
                        asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(f_other), "h"((unsigned short)0) : "memory");
                    }
                    fence_proxy_async_smem();
                }
                tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(&bar.p_full[g & 1]);
            }
            // ---- epilogue: O / l -> bf16 -> staging tile -> TMA store
            s_lx[r_in_tile * 2 + hf] = l_sum;
            named_bar_sync(bar_id, 64);
            const float l_tot = l_sum + s_lx[r_in_tile * 2 + (hf ^ 1)];
            const float inv = 1.f / l_tot;
            if (hf == 0 && row_ok && P.l) P.l[(int64_t)bh * P.Nq + row] = 1.f / (fast_exp2(m_ref * SCALE_LOG2) * l_tot);
            if (HAS_CS) {
                mbar_wait(&bar.cs_full, (g - 1) & 1);            // the last column sums; the P tile is free for the staging
                tc_fence_after_sync();
                if (hf == 0) drain_cs(nk - 1);
            }
            mbar_wait(&bar.pv_done[(g - 1) % 3], ((g - 1) / 3) & 1);
            tc_fence_after_sync();
            {
                uint32_t r[64];
                tmem_ld32(tO, r);
                tmem_ld32(tO + 32, r + 32);
                tmem_ld_wait();
                const uint32_t orow = sP + hf * HALF_BYTES + r_in_tile * 128;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        w[j] = pack_bf16x2(__uint_as_float(r[8 * c + 2 * j]) * inv, __uint_as_float(r[8 * c + 2 * j + 1]) * inv);
                    st_shared_v4(orow + (((uint32_t)c ^ sw) << 4), w[0], w[1], w[2], w[3]);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            named_bar_sync(6, 256);
            if (tid == 0) {
                const Coord co = coords(P.pos[3], 0, t * BM, h, b);
                tma_store_4d(&tm_o, sP, 0, co.c[1], co.c[2], co.c[3]);
                tma_store_4d(&tm_o, sP + HALF_BYTES, 64, co.c[1], co.c[2], co.c[3]);
                bulk_commit();
            }
        }
        if (tid == 0) bulk_wait<0>();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tm, 512);
}

}  // namespace dense
}  // namespace cm

// ------------------------------------------------------------------------------------------
using namespace cm;
using namespace cm::dense;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool st_ok(const int64_t s[3]) { return s[0] % 8 == 0 && s[1] % 8 == 0 && s[2] % 8 == 0 && s[2] >= D; }

// [B,H,N,128] view with element strides st = {batch, head, row}: a 4-D tensor map whose outer dimensions are ordered by
// ascending stride (size-1 dimensions last), boxes of 128 rows x 64 columns.  pos[] = where (row, head, batch) landed.
static int make_map(CUtensorMap* m, const void* base, int B, int H, int N, const int64_t st[3], int8_t pos[3]) {
    struct Dim { uint64_t size, stride; int role; } d[3] = {{(uint64_t)N, (uint64_t)st[2] * 2, 0}, {(uint64_t)H, (uint64_t)st[1] * 2, 1},
                                                             {(uint64_t)B, (uint64_t)st[0] * 2, 2}};
    auto key = [](const Dim& x) { return x.size == 1 ? ~0ull : x.stride; };
    for (int i = 0; i < 3; i++)
        for (int j = i + 1; j < 3; j++)
            if (key(d[j]) < key(d[i])) { Dim t = d[i]; d[i] = d[j]; d[j] = t; }
    uint64_t dims[4] = {(uint64_t)D, 0, 0, 0}, strides[3];
    uint32_t box[4] = {64, 1, 1, 1};
    uint64_t prev = D * 2;
    for (int i = 0; i < 3; i++) {
        dims[i + 1] = d[i].size;
        // a size-1 dimension's stride is never used for addressing; give it a valid monotone value
        strides[i] = d[i].size == 1 ? prev : d[i].stride;
        prev = strides[i] * d[i].size;
        pos[d[i].role] = (int8_t)(i + 1);
        if (d[i].role == 0) box[i + 1] = 128;
    }
    return encode_tmap_4d_bf16_sw128(m, base, dims, strides, box);
}

extern "C" int cm_dense_attn_strided(const void* q, const void* k, const void* v, void* o, float* l, void* cs, const float* p,
                                     int B, int H, int Nq, int Nk, const int64_t q_strides[3], const int64_t k_strides[3],
                                     const int64_t v_strides[3], const int64_t o_strides[3], int64_t cs_row_stride,
                                     void* stream) {
    if (B < 0 || H < 0 || Nq < 0 || Nk <= 0) return CM_EINVAL;
    if ((int64_t)B * H * Nq == 0) return CM_OK;
    if (!q || !k || !v || !o) return CM_EINVAL;
    if (cs && (!p || cs_row_stride < Nk || cs_row_stride % 8 != 0)) return CM_EINVAL;
    if (!al16(q) || !al16(k) || !al16(v) || !al16(o) || (cs && !al16(cs))) return CM_EALIGN;
    if (!st_ok(q_strides) || !st_ok(k_strides) || !st_ok(v_strides) || !st_ok(o_strides)) return CM_EALIGN;
    if (!is_sm100()) return CM_EARCH;
    CUtensorMap mq, mk, mv, mo;
    Params P{};
    int rc = make_map(&mq, q, B, H, Nq, q_strides, P.pos[0]);
    if (!rc) rc = make_map(&mk, k, B, H, Nk, k_strides, P.pos[1]);
    if (!rc) rc = make_map(&mv, v, B, H, Nk, v_strides, P.pos[2]);
    if (!rc) rc = make_map(&mo, o, B, H, Nq, o_strides, P.pos[3]);
    if (rc) return rc;
    P.l = l; P.p = p; P.cs = (__nv_bfloat16*)cs; P.cs_stride = cs_row_stride;
    P.B = B; P.H = H; P.Nq = Nq; P.Nk = Nk; P.G = (Nq + QG - 1) / QG;
    P.tiles_per_head = (Nq + BM - 1) / BM;
    const int64_t tiles = (int64_t)B * H * P.tiles_per_head;
    if (tiles > 2147483647ll) return CM_EINVAL;
    P.num_tiles = (int)tiles;
    P.nk = (Nk + KT - 1) / KT;
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = P.num_tiles < sm_count() ? P.num_tiles : sm_count();
    if (cs) {
        // every element receives two partial sums through red.add: start from zero
        cudaError_t e = cudaMemsetAsync(cs, 0, (size_t)B * H * P.G * cs_row_stride * 2, s);
        if (e != cudaSuccess) return (int)e;
        static unsigned long long configured = 0;
        rc = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(dense_kernel<true>), SMEM_BYTES);
        if (rc) return rc;
        dense_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(mq, mk, mv, mo, P);
    } else {
        static unsigned long long configured = 0;
        rc = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(dense_kernel<false>), SMEM_BYTES);
        if (rc) return rc;
        dense_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(mq, mk, mv, mo, P);
    }
    return (int)cudaGetLastError();
}
