// Dense attention with the statistics the sparse steps need, in ONE pass: o, l and (optionally) the per-group
// column sums cs.  Replaces csrc/attn/dense_attn.cu and csrc/attn/dense_colsum_attn.cu of the reference (Hopper
// wgmma kernels that compute o, l -- and cs -- in one flash-attention loop, dense_colsum_attn.cu:205-341).
//
// Work unit: one (b, h, 128 query rows) tile, walked over all keys 128 at a time.  All eight softmax warps work on
// that one tile -- two threads per query row, 64 keys each -- and S is double-buffered in TMEM, so the tensor pipe
// always has the next S ready when a softmax step ends (the 192/256-row kernels of csp_attn.cu keep two query
// blocks in flight instead and are bound by each block's serial softmax -> P.V -> S chain).
//
//   warp 13      TMA producer: Q tile once per tile; K / V tiles (128 rows x 256 B, two 64-wide halves, 128B-swizzled)
//                through a 4-slot ring.  4-D tensor maps: any batch / head / row stride, rows past N read as zero.
//   warp 12      MMA issuer (one elected thread):
//                  S(k)    = Q K(k)^T           SS, M=128 N=128 K=128  -> TMEM S buffer k&1
//                  CS(k)^T = P(k)^T F(k)^T      SS, M=128 keys, N=16, K=128 queries -> TMEM CS   (column sums, see below)
//                  O      += P(k) V(k)          TS (P read from TMEM, where it overwrote S), V as MN-major smem
//   warps 0-7    softmax: thread (row r, half h) owns keys [64h, 64h+64) of the step and head dims [64h, 64h+64) of O;
//                the pair shares its tile maximum through shared memory (64-thread named barrier), running maximum
//                with lazy rescale, exp2, bf16 P written over S in TMEM (and, for the column sums, to shared memory).
//                Epilogue (same threads): O / l -> bf16 -> 128B-swizzled staging tile -> TMA store; l to global.
//   warps 8-11   column-sum drain, one warp per TMEM lane quadrant: CS^T (key on the lane, group on the column) ->
//                bf16 -> 16-byte red.global.add into cs.
//
// Column sums.  cs[b,h,g,j] = sum_{i in 192-row group g} exp(s_ij / sqrt(d)) p_i   (dense_colsum_attn.cu:267-277).
// With P_ij = exp2(s_ij c - m_i c) in hand (m_i = the row's running reference maximum), that is
//     cs[g, j] = sum_i P_ij f_i ,   f_i = exp2(m_i c + log2 p_i)   over the rows i of the tile that lie in group g:
// a [128 keys x 128 queries] x [128 queries x 16] product per step, i.e. the TENSOR PIPE does the cross-row reduction
// at 1/8 of the cost of S (A = P^T: the P tile in shared memory read MN-major; B = F: row 0 / row 1 hold the f_i of the
// tile's first / second group, rows 2-15 zero).  (Round 2 first used an M=64 MMA of F x P: same result, but as
// expensive as a full S and on the S -> softmax -> P.V chain.)  A 128-row tile overlaps at most two 192-row groups
// and every group is covered by two tiles, so each cs element receives exactly two bf16 partial sums
// (fp32-accumulated over up to 128 rows each), added at the L2.  The reference reduces twelve warps' bf16 partials
// with shared-memory atomics.
//
// TMEM (512 columns): S0 [0,128)  S1 [128,256)  O [256,384)  CS [384,400).
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
constexpr int SMEM_BYTES = TILE_BYTES /*Q*/ + TILE_BYTES /*P / staging*/ + NSLOT * TILE_BYTES + F_BYTES + 1024 /*align*/;
constexpr int NUM_THREADS = 512;               // warps 0-7 softmax | 8-11 column-sum drain (one per TMEM lane quadrant) | 12 MMA | 13 TMA | 14-15 idle
constexpr int WARP_DRAIN0 = 8, WARP_MMA = 12, WARP_TMA = 13;
constexpr int REG_SOFTMAX = 192, REG_OTHER = 64;    // setmaxnreg budgets: 256 x 192 + 256 x 64 = 64 K registers
constexpr uint32_t TM_S = 0, TM_O = 256, TM_CS = 384;

#ifndef CM_DENSE_POLY
#define CM_DENSE_POLY 4       // every CM_DENSE_POLY-th pair of scores is exponentiated on the FMA pipe (0: all on the MUFU)
#endif
// exp2 of two fp32 on the FMA / ALU pipes: x = n + f (round to nearest with the 1.5*2^23 trick), 2^f by a degree-3
// Chebyshev fit on [-0.5, 0.5] (relative error 1.0e-4, far below the bf16 rounding of P), 2^n added into the exponent
// field.  The two softmax warps of an SM sub-partition exponentiate at the same time (they meet on the pair barrier every
// step), so the 4-lane MUFU is contended during that phase while the FMA pipe idles: moving a share of the exponentials
// there shortens the phase.  (In the column-sparse kernel, one warp per sub-partition, the same trick gains nothing.)
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x) {
    float x0, x1;
    unpack_f32x2(x, x0, x1);
    const uint64_t xc = pack_f32x2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
    const uint64_t t = fadd2(xc, pack_f32x2(12582912.f, 12582912.f));
    const uint64_t n = fadd2(t, pack_f32x2(-12582912.f, -12582912.f));
    const uint64_t f = ffma2(n, pack_f32x2(-1.f, -1.f), xc);
    uint64_t p = ffma2(f, pack_f32x2(0.0559220356f, 0.0559220356f), pack_f32x2(0.2426400828f, 0.2426400828f));
    p = ffma2(p, f, pack_f32x2(0.6931210340f, 0.6931210340f));
    p = ffma2(p, f, pack_f32x2(0.9999244815f, 0.9999244815f));
    float t0, t1, q0, q1;
    unpack_f32x2(t, t0, t1);
    unpack_f32x2(p, q0, q1);
    return pack_f32x2(__uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23)),
                      __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23)));
}

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
    uint64_t s_full[2];
    uint64_t p_full, pv_done, cs_full, cs_empty;
};

using Coord = TmaCoord;
__device__ __forceinline__ Coord coords(const int8_t pos[3], int col, int row, int h, int b) { return tma_coords(pos, col, row, h, b); }

template <bool HAS_CS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dense_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
             const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_mx[2][2 * BM];      // tile maxima of the two half-row threads, double-buffered by step parity
    __shared__ float s_lx[2 * BM];         // partial row sums, exchanged in the epilogue

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = sbase, sP = sbase + TILE_BYTES, sF = sbase + 2 * TILE_BYTES, sKV = sF + F_BYTES;

    if (tid == 0) {
        mbar_init(&bar.q_full, 1); mbar_init(&bar.q_empty, 1);
        for (int i = 0; i < NSLOT; i++) { mbar_init(&bar.kv_full[i], 1); mbar_init(&bar.kv_empty[i], 1); }
        mbar_init(&bar.s_full[0], 1); mbar_init(&bar.s_full[1], 1);
        mbar_init(&bar.p_full, 256); mbar_init(&bar.pv_done, 1);
        mbar_init(&bar.cs_full, 1); mbar_init(&bar.cs_empty, 4);
        fence_mbar_init();
    }
    if (warp == WARP_MMA) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (warp == WARP_TMA && lane == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v); tma_prefetch_desc(&tm_o);
    }
    if (HAS_CS) {
        // rows 2-15 of the F operand stay zero for the whole kernel
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
        setmaxnreg_dec<REG_OTHER>();
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
                // consumption order of the ring: K0, K1, then V(k), K(k+2) for k = 0 ...
                load_kv(&tm_k, P.pos[1], 0, h, b);
                if (nk > 1) load_kv(&tm_k, P.pos[1], 1, h, b);
                for (int k = 0; k < nk; k++) {
                    load_kv(&tm_v, P.pos[2], k, h, b);
                    if (k + 2 < nk) load_kv(&tm_k, P.pos[1], k + 2, h, b);
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // =================================================================================== MMA issuer
        setmaxnreg_dec<REG_OTHER>();
        uint32_t job = 0, it = 0, pc = 0, cc = 0;       // ring jobs, tiles, steps (p_full phases), cs phases
        const uint32_t idesc_s = umma_idesc_bf16(128, KT, 0, 0);
        const uint32_t idesc_pv = umma_idesc_bf16(128, D, 0, 1);
        const uint32_t idesc_cs = umma_idesc_bf16(128, 16, 1, 0);
        const uint64_t desc_q = umma_smem_desc(sQ, 16, 1024);                 // K-major A: Q rows
        const uint64_t desc_k = umma_smem_desc(sKV, 16, 1024);                // K-major B: K rows
        const uint64_t desc_v = umma_smem_desc(sKV, HALF_BYTES, 1024);        // MN-major B: V rows (k = key, n = head dim)
        const uint64_t desc_f = umma_smem_desc(sF, 16, 1024);                 // K-major B: F rows (n = group row, k = query)
        const uint64_t desc_p = umma_smem_desc(sP, HALF_BYTES, 1024);         // MN-major A: P rows (k = query, m = key)
        auto issue_S = [&](uint32_t buf, uint32_t slot) {
            const uint64_t bd0 = desc_k + (uint64_t)(slot * (TILE_BYTES >> 4));
#pragma unroll
            for (int k16 = 0; k16 < D / 16; k16++) {
                const uint64_t off = (uint64_t)((((k16 >> 2) * HALF_BYTES) + (k16 & 3) * 32) >> 4);
                umma_ss(tm + TM_S + buf * 128, desc_q + off, bd0 + off, idesc_s, k16 > 0);
            }
        };
        auto issue_PV = [&](uint32_t buf, uint32_t slot, bool first) {
            const uint64_t bd0 = desc_v + (uint64_t)(slot * (TILE_BYTES >> 4));
#pragma unroll
            for (int j = 0; j < KT / 16; j++)
                umma_ts(tm + TM_O, tm + TM_S + buf * 128 + j * 8, bd0 + (uint64_t)(j * (2048 >> 4)), idesc_pv, (!first) || j > 0);
        };
        auto issue_CS = [&]() {       // CS^T [128 keys x 16] = P^T [128 keys x 128 queries] . F^T [128 queries x 16]: 1/8 of an S
#pragma unroll
            for (int j = 0; j < BM / 16; j++) {
                const uint64_t boff = (uint64_t)((((j >> 2) * (F_BYTES / 2)) + (j & 3) * 32) >> 4);
                umma_ss(tm + TM_CS, desc_p + (uint64_t)(j * (2048 >> 4)), desc_f + boff, idesc_cs, j > 0);
            }
        };
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, it++) {
            mbar_wait(&bar.q_full, it & 1);
            {   // prologue: S(0), S(1)
                const uint32_t s0 = job % NSLOT;
                mbar_wait(&bar.kv_full[s0], (job / NSLOT) & 1);
                job++;
                tc_fence_after_sync();
                if (elect_one()) {
                    issue_S(0, s0); umma_commit(&bar.s_full[0]); umma_commit(&bar.kv_empty[s0]);
                    if (nk == 1) umma_commit(&bar.q_empty);
                }
                __syncwarp();
                if (nk > 1) {
                    const uint32_t s1 = job % NSLOT;
                    mbar_wait(&bar.kv_full[s1], (job / NSLOT) & 1);
                    job++;
                    tc_fence_after_sync();
                    if (elect_one()) {
                        issue_S(1, s1); umma_commit(&bar.s_full[1]); umma_commit(&bar.kv_empty[s1]);
                        if (nk == 2) umma_commit(&bar.q_empty);
                    }
                    __syncwarp();
                }
            }
            for (int k = 0; k < nk; k++) {
                const uint32_t sv = job % NSLOT;
                mbar_wait(&bar.kv_full[sv], (job / NSLOT) & 1);
                job++;
                uint32_t sk = 0;
                const bool more = k + 2 < nk;
                if (more) {
                    sk = job % NSLOT;
                    mbar_wait(&bar.kv_full[sk], (job / NSLOT) & 1);
                    job++;
                }
                if (HAS_CS && cc > 0) mbar_wait(&bar.cs_empty, (cc - 1) & 1);      // the previous column sums have left TMEM
                mbar_wait(&bar.p_full, pc & 1); pc++;
                tc_fence_after_sync();
                if (elect_one()) {
                    if (HAS_CS) { issue_CS(); umma_commit(&bar.cs_full); }
                    issue_PV(k & 1, sv, k == 0);
                    umma_commit(&bar.kv_empty[sv]);
                    umma_commit(&bar.pv_done);
                    if (more) {
                        issue_S(k & 1, sk); umma_commit(&bar.s_full[k & 1]); umma_commit(&bar.kv_empty[sk]);
                        if (k + 3 == nk) umma_commit(&bar.q_empty);
                    }
                }
                __syncwarp();
                if (HAS_CS) cc++;
            }
        }
    } else if (warp >= WARP_DRAIN0 && warp < WARP_DRAIN0 + 4) {
        // =================================================================================== column-sum drain
        // CS^T sits in TMEM as [128 lanes = keys of the step] x [column 0 / 1 = the tile's first / second group]: warp q
        // reads lane quadrant q, gathers 8 consecutive keys per lane group with shuffles and adds them into cs with
        // 16-byte bf16 reductions at the L2 (each cs element receives two such partial sums, from two tiles)
        setmaxnreg_dec<REG_OTHER>();
        if (HAS_CS) {
            const int q4 = warp - WARP_DRAIN0;
            const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
            uint32_t cc = 0;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
                const int t = tile % P.tiles_per_head, bh = tile / P.tiles_per_head;
                const int row0 = t * BM;
                const int gA = row0 / QG;
                const bool has_b = row0 + BM - 1 >= QG * (gA + 1) && gA + 1 < P.G;
                __nv_bfloat16* cs_a = P.cs + ((int64_t)bh * P.G + gA) * P.cs_stride;
                __nv_bfloat16* cs_b = cs_a + P.cs_stride;
                const int base = lane & ~7;
                const bool for_b = (lane & 4) != 0;
                for (int k = 0; k < nk; k++, cc++) {
                    mbar_wait(&bar.cs_full, cc & 1);
                    tc_fence_after_sync();
                    uint32_t c0, c1;
                    tmem_ld_32x32b_x2(tm + TM_CS + lane_off, c0, c1);
                    tmem_ld_wait();
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar.cs_empty);
                    const uint32_t nb = __shfl_down_sync(0xffffffffu, c0, 1), nb1 = __shfl_down_sync(0xffffffffu, c1, 1);
                    const uint32_t pa = pack_bf16x2(__uint_as_float(c0), __uint_as_float(nb));     // valid on even lanes
                    const uint32_t pb = pack_bf16x2(__uint_as_float(c1), __uint_as_float(nb1));
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t va = __shfl_sync(0xffffffffu, pa, base + 2 * j);
                        const uint32_t vb = __shfl_sync(0xffffffffu, pb, base + 2 * j);
                        w[j] = for_b ? vb : va;
                    }
                    const int key = k * KT + q4 * 32 + base;             // first of this lane group's 8 keys
                    if ((lane & 3) == 0 && key < P.cs_stride && (!for_b || has_b))
                        red_add_bf16x8((for_b ? cs_b : cs_a) + key, w[0], w[1], w[2], w[3]);
                }
            }
        }
    } else if (warp < 8) {
        // =================================================================================== softmax + epilogue
        setmaxnreg_inc<REG_SOFTMAX>();
        const int q4 = warp & 3, hf = warp >> 2;
        const int r_in_tile = q4 * 32 + lane;
        const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
        const uint32_t tO = tm + TM_O + lane_off + hf * 64;
        const uint32_t bar_id = 1 + q4;
        uint32_t sc0 = 0, sc1 = 0;         // uses of each S buffer
        uint32_t gstep = 0;                // steps done by this CTA (phases of p_full / pv_done / cs_full)
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
            if (HAS_CS && hf == 0) {
                const float pv = row_ok ? __ldg(P.p + (int64_t)bh * P.Nq + row) : 0.f;
                lp = pv > 0.f ? __log2f(pv) : -INFINITY;
                const int gA = (t * BM) / QG;
                const uint32_t r01 = row >= QG * (gA + 1) ? 1u : 0u;      // F row of this query's group
                const uint32_t base = sF + (uint32_t)(r_in_tile >> 6) * (F_BYTES / 2) + (uint32_t)(r_in_tile & 7) * 2;
                const uint32_t chunk = (uint32_t)((r_in_tile & 63) >> 3);
                f_addr = base + r01 * 128 + ((chunk ^ r01) << 4);
                f_other = base + (r01 ^ 1u) * 128 + ((chunk ^ (r01 ^ 1u)) << 4);
            }
            float m_ref = -INFINITY, l_sum = 0.f;

            for (int k = 0; k < nk; k++, gstep++) {
                const uint32_t buf = k & 1;
                const uint32_t tS = tm + TM_S + buf * 128 + lane_off;
                const int valid = P.Nk - k * KT;                 // >= 128 except on the last step
                if (buf == 0) { mbar_wait(&bar.s_full[0], sc0 & 1); sc0++; } else { mbar_wait(&bar.s_full[1], sc1 & 1); sc1++; }
                tc_fence_after_sync();
                uint32_t s[64];
                tmem_ld32(tS + hf * 64, s);
                tmem_ld32(tS + hf * 64 + 32, s + 32);
                tmem_ld_wait();
                if (valid < KT) {
#pragma unroll
                    for (int j = 0; j < 64; j++) s[j] = (hf * 64 + j < valid) ? s[j] : 0xff800000u;
                }
                float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[32]);
#pragma unroll
                for (int j = 1; j < 31; j += 2) {
                    mx0 = fmax3(mx0, __uint_as_float(s[j]), __uint_as_float(s[j + 1]));
                    mx1 = fmax3(mx1, __uint_as_float(s[32 + j]), __uint_as_float(s[32 + j + 1]));
                }
                const float m_part = fmaxf(fmaxf(mx0, __uint_as_float(s[31])), fmaxf(mx1, __uint_as_float(s[63])));
                s_mx[buf][r_in_tile * 2 + hf] = m_part;
                named_bar_sync(bar_id, 64);          // also: both threads have loaded their S half before either writes P over it
                const float m_tile = fmaxf(m_part, s_mx[buf][r_in_tile * 2 + (hf ^ 1)]);
                // (m_tile - m_ref) is NaN when both are -inf (a fully masked half cannot happen: valid >= 1): m_ref = -inf only on step 0
                const bool need = (m_tile - m_ref) * SCALE_LOG2 > RESCALE_THRESHOLD;
                if (__any_sync(0xffffffffu, need)) {
                    float alpha = 1.f;
                    if (need) {
                        alpha = fast_exp2((m_ref - m_tile) * SCALE_LOG2);
                        m_ref = m_tile;
                        l_sum *= alpha;
                    }
                    if (k > 0) {
                        mbar_wait(&bar.pv_done, (gstep - 1) & 1);        // O += P(k-1) V(k-1) has landed
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
                const float neg_m = -m_ref * SCALE_LOG2;
                const uint64_t nm2 = pack_f32x2(neg_m, neg_m);
                if (HAS_CS) {
                    // the previous step's column-sum MMA has finished reading the P tile and F
                    if (k > 0) mbar_wait(&bar.cs_full, (gstep - 1) & 1);
                    if (hf == 0) {
                        const float f = fast_exp2(m_ref * SCALE_LOG2 + lp);
                        const uint32_t fb = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(f));
                        asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(f_addr), "h"((unsigned short)fb) : "memory");
                        asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(f_other), "h"((unsigned short)0) : "memory");
                    }
                }
                uint64_t acc[2] = {0ull, 0ull};
#pragma unroll
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint64_t x = ffma2(pack_f32x2(__uint_as_float(s[c0 + j]), __uint_as_float(s[c0 + j + 1])), c2, nm2);
                        float x0, x1;
                        unpack_f32x2(x, x0, x1);
                        float p0, p1;
                        if (CM_DENSE_POLY > 0 && ((j >> 1) % (CM_DENSE_POLY > 0 ? CM_DENSE_POLY : 1)) == 0) unpack_f32x2(exp2_poly2(x), p0, p1);
                        else { p0 = fast_exp2(x0); p1 = fast_exp2(x1); }
                        acc[(j >> 1) & 1] = fadd2(acc[(j >> 1) & 1], pack_f32x2(p0, p1));
                        pk[j >> 1] = pack_bf16x2(p0, p1);
                    }
                    tmem_st_32x32b_x16(tS + hf * 32 + (c0 >> 1), pk);
                    if (HAS_CS) {
                        const uint32_t prow = sP + hf * HALF_BYTES + r_in_tile * 128;
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            st_shared_v4(prow + ((((uint32_t)(c0 >> 3) + q) ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                    }
                }
                float a0, a1, a2, a3;
                unpack_f32x2(acc[0], a0, a1);
                unpack_f32x2(acc[1], a2, a3);
                l_sum += (a0 + a1) + (a2 + a3);
                tmem_st_wait();
                if (HAS_CS) fence_proxy_async_smem();
                tc_fence_before_sync();
                mbar_arrive(&bar.p_full);
            }
            // ---- epilogue: O / l -> bf16 -> staging tile -> TMA store
            s_lx[r_in_tile * 2 + hf] = l_sum;
            named_bar_sync(bar_id, 64);
            const float l_tot = l_sum + s_lx[r_in_tile * 2 + (hf ^ 1)];
            const float inv = 1.f / l_tot;
            if (hf == 0 && row_ok && P.l) P.l[(int64_t)bh * P.Nq + row] = 1.f / (fast_exp2(m_ref * SCALE_LOG2) * l_tot);
            mbar_wait(&bar.pv_done, (gstep - 1) & 1);
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
    } else {
        setmaxnreg_dec<REG_OTHER>();        // warps 14-15: idle (setmaxnreg is warpgroup-wide)
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

static int make_map(CUtensorMap* m, const void* base, int B, int H, int N, const int64_t st[3], int8_t pos[3]) {
    return encode_tmap_bhnd(m, base, B, H, N, st, 128, pos);
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
