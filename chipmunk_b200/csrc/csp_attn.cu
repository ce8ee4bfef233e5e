// Column-sparse "delta" attention for sm_100a.
//
// Replaces csrc/attn/{csp_attn,csp_128_attn}.cu of the reference (Hopper wgmma + ThunderKittens) with one
// warp-specialised tcgen05 kernel.  (The dense full-step operators live in dense_attn.cu.)
//
// Work unit ("tile"): one (batch, head, group of 192 query rows).  Per tile the kernel walks the
// group's selected key columns 128 at a time:
//     TMA thread           loads the dense Q tile and the cached output tile (4-D tensor maps: any batch / head / row
//                          stride, rows past N read as zero), and stores / reduce-adds the finished output tile,
//     producers (4 warps)  gather K[idx] / V[idx] rows (256 B each) with 16-byte cp.async into
//                          128B-swizzled shared-memory slots (a 4-deep ring of 32 KB slots),
//     MMA warp (1 thread)  S = Q K^T   (tcgen05.mma SS, M=128 N<=128 K=128, fp32 in TMEM), twice:
//                          query rows 0-127 ("block 0") and 128-191 ("block 1"),
//                          O += P V    (tcgen05.mma TS: P read from TMEM, V as MN-major smem),
//     softmax warps (4+4)  TMEM -> regs, running max with lazy rescale, exp2, row sum, P (bf16) written back over S.
//                          Block 0 (M = 128): one thread per query row, one warp per TMEM lane quadrant.  Block 1 is
//                          an M = 64 accumulator (row m on lane (m % 16) + 32 (m / 16): 16 rows per quadrant), one
//                          warp per quadrant with FOUR threads per row (16-lane TMEM shapes, softmax_step16): every
//                          lane works.  An M = 64 MMA takes as many cycles as an M = 128 one but far less power
//                          (tests/probes/probe_mma_power.cu: under the 1 kW cap a pure MMA stream clocks 1.81 GHz at
//                          M = 64, 1.57 GHz at M = 128 with 64 zero rows), and the kernel runs at the cap,
//     epilogue             (same threads) O / l * o_scale -> bf16 (+ the cached tile, read from shared memory) into a
//                          128B-swizzled staging tile that the TMA thread stores (csp_128_attn, csp_attn_add) or
//                          reduce-adds at the L2 (csp_attn, like the reference's TMA store_add, csp_attn.cu:300).
// The two query blocks ping-pong on the tensor pipe: while the softmax warps of block 0 work on
// S0(k), the pipe runs S1(k) / P1 V(k-1), and vice versa.
//
// TMEM map (512 columns): S0 [0,128)  S1 [128,256)  O0 [256,384)  O1 [384,512);  P aliases the
// first 64 columns of its S.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <type_traits>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include "attn_common.cuh"
#include "tma.cuh"

#ifndef CM_ATTN_REG_SOFTMAX
#define CM_ATTN_REG_SOFTMAX 208      // one thread per row: a 128-column score row per thread
#endif

namespace cm {
namespace attn {

constexpr int NSLOT = 4;            // 32 KB K/V slots
constexpr int SLOT_BYTES = KT * D * 2;
// Geometry: the reference's 192-query index groups = one M=128 block + one M=64 block.
struct GEO {
    static constexpr int QROWS = QG;                             // query rows per tile
    static constexpr int Q_HALF_BYTES = QROWS * 128;             // one 64-wide d-half of the Q tile
    static constexpr int Q_BYTES = 2 * Q_HALF_BYTES;             // 48 KB
    static constexpr int STAGE_BYTES = Q_BYTES;                  // 48 KB: the cached output tile in, the output tile out (TMA both ways)
    static constexpr int SMEM_BYTES = Q_BYTES + STAGE_BYTES + NSLOT * SLOT_BYTES + 1024 /*align slack*/;
    // warps 0-3 softmax blk0 (rows 32 w + lane) | 4-7 softmax blk1 (rows 128 + 16 (w&3) + 0..15, four threads per row) |
    // 8 MMA, 9 TMA (Q tile, cached tile, output tile), 10-11 idle | 12-15 K/V gather producers
    static constexpr int NUM_THREADS = 512;
    static constexpr int WARP_MMA = 8;
    static constexpr int WARP_TMA = 9;
    static constexpr int NUM_SOFTMAX_WARPS = 8;
    // setmaxnreg budgets per warpgroup (softmax block 0 | softmax block 1 | MMA, TMA | producers); their sum must not exceed the
    // launch allocation (512 x 128): a larger sum can never be granted and the kernel hangs in setmaxnreg.inc.
    // QUAD = block 1 with four threads per row (two 32-column pieces per thread) leaves room for the other roles.
    static constexpr int REG_SOFTMAX = CM_ATTN_REG_SOFTMAX;
    template <bool QUAD> static constexpr int reg_softmax1() { return QUAD ? 160 : 208; }
    template <bool QUAD> static constexpr int reg_mma() { return QUAD ? 64 : 32; }
    template <bool QUAD> static constexpr int reg_prod() { return QUAD ? 80 : 64; }
    static_assert(128 * (REG_SOFTMAX + 160 + 64 + 80) <= 65536 && 128 * (REG_SOFTMAX + 208 + 32 + 64) <= 65536, "setmaxnreg budgets exceed the register file");
};
constexpr int WARP_PROD0 = 12;
constexpr int INACTIVE_ROW = 1 << 28;          // r_in_tile of the lanes 16-31 of a block-1 softmax warp: beyond every Nq
constexpr int NUM_PROD = 128;

constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O0 = 256, TM_O1 = 384;


struct Params {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;
    const int32_t* indices;
    const int32_t* counts;
    const __nv_bfloat16* cache;  // fused add-back: o = bf16(cache + o_scale * delta), out of place (null: see `accumulate`)
    int B, H, Nq, Nk, G;
    int64_t qs[3], ks[3], vs[3], os[3], cs[3];
    int64_t idx_row_stride;
    float o_scale;
    int accumulate;
    int wide;                    // o (and cache) rows are 32-byte aligned: the epilogue moves full sectors per access
    int64_t mc_delta;            // != 0: o lives in a symmetric buffer; rows are stored to (o + mc_delta bytes), its NVLS multicast alias
    int n_peers;                 // > 0: no multicast; rows are stored to (o + peer_delta[p] bytes) for every GPU p of the group (own copy included)
    int64_t peer_delta[8];
    int stage;                   // 1: output (and cache) tiles go through shared memory and the TMA; 0: multicast epilogue (direct stores)
    int8_t pos[3][3];            // coordinate slots of (row, head, batch) in the q / cache / o tensor maps
    int num_tiles;
    int dbg;                     // stage-isolation timing switches: only read in -DCM_DEBUG_STAGES builds (CM_DBG below)
};

struct __align__(8) Barriers {
    uint64_t q_full, q_empty;
    uint64_t kv_full[NSLOT], kv_empty[NSLOT];
    uint64_t s_full[2], p_full[2], o_full[2];
    uint64_t c_full, st_full, st_free;       // cached tile landed | staging tile written by the epilogue | read out by the TMA store
};

// one 16-byte piece of an output row into every GPU's copy of the symmetric output buffer: one multimem.st through the
// NVSwitch (NVLS), or one plain store per peer over NVLink when the group has no multicast object
__device__ __forceinline__ void bcast_st_v4(const Params& P, char* local, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    if (P.mc_delta != 0) { multimem_st_v4(local + P.mc_delta, a, b, c, d); return; }
    for (int p = 0; p < P.n_peers; p++) *reinterpret_cast<uint4*>(local + P.peer_delta[p]) = make_uint4(a, b, c, d);
}

__device__ __forceinline__ int tile_count(const Params& P, int tile) {
    int c = __ldg(P.counts + tile);
    c = c < 0 ? 0 : c;
    return c > (int)P.idx_row_stride ? (int)P.idx_row_stride : c;
}

// ------------------------------------------------------------------------------------------
// QUAD: block 1's softmax runs with four threads per row, every lane busy (softmax_step16) -- the lower-power form, faster
// on long key lists under the power cap (-2.2 % at the 720p shape); with one thread per row on lanes 0-15 the kernel is smaller
// and has the lower per-tile cost (-2.9 % at FLUX sizes, 7 steps per tile).  Same TMEM layout: the MMA side is identical.
template <bool QUAD>
__global__ void __launch_bounds__(GEO::NUM_THREADS, 1)
attn_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ CUtensorMap tm_o, const Params P) {
    constexpr int QROWS = GEO::QROWS, Q_HALF_BYTES = GEO::Q_HALF_BYTES, Q_BYTES = GEO::Q_BYTES, WARP_MMA = GEO::WARP_MMA;
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = sbase;
    const uint32_t sStage = sbase + Q_BYTES;
    const uint32_t sKV = sStage + GEO::STAGE_BYTES;

    if (tid == 0) {
        mbar_init(&bar.q_full, 1);
        mbar_init(&bar.c_full, 1); mbar_init(&bar.st_full, QROWS); mbar_init(&bar.st_free, 1);
        mbar_init(&bar.q_empty, 1);
        for (int i = 0; i < NSLOT; i++) { mbar_init(&bar.kv_full[i], NUM_PROD); mbar_init(&bar.kv_empty[i], 1); }
        mbar_init(&bar.s_full[0], 1);  mbar_init(&bar.s_full[1], 1);
        mbar_init(&bar.p_full[0], 128); mbar_init(&bar.p_full[1], 128);     // every thread of a block's four softmax warps
        mbar_init(&bar.o_full[0], 1);  mbar_init(&bar.o_full[1], 1);
        fence_mbar_init();
    }
    if (warp == WARP_MMA) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;

    // =========================================================================== producers
    if (warp >= WARP_PROD0 && warp < WARP_PROD0 + 4) {
        setmaxnreg_dec<GEO::reg_prod<QUAD>()>();
        const int pt = tid - WARP_PROD0 * 32;       // 0..127
        const int chunk = pt & 15;                  // 16-byte chunk of the 256-byte row
        const int rsub = pt >> 4;                   // 0..7
        const uint32_t half_off = (uint32_t)(chunk >> 3);
        const uint32_t c8 = (uint32_t)(chunk & 7);
        uint32_t job = 0;                           // K/V slot fills issued so far
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, it++) {
            const int count = tile_count(P, tile);
            if (count <= 0) { it--; continue; }
            const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
            const __nv_bfloat16* kb = P.k + b * P.ks[0] + h * P.ks[1];
            const __nv_bfloat16* vb = P.v + b * P.vs[0] + h * P.vs[1];
            const int32_t* ip = P.indices + (int64_t)tile * P.idx_row_stride;
            const int nk = (count + KT - 1) / KT;
            // Lane j of producer warp w fetches the index of key row (2w + (j>>4)) + 8*(j&15) of the
            // step: one coalesced-ish load per thread per step, issued one step ahead; the 16 rows a
            // thread copies are then read from its half-warp with shuffles.
            const int my_r = rsub + 8 * (lane & 15);
            // fetch_idx returns the RAW loaded value (no arithmetic on it: that would make the warp wait for
            // the load at the fetch and defeat the one-step-ahead prefetch); fix_idx clamps it at the use site
            auto fetch_idx = [&](int kk) -> int {
                const int pos = kk * KT + my_r;
                return pos < count ? __ldg(ip + pos) : 0;
            };
            auto fix_idx = [&](int kk, int idx) -> int {
                const int pos = kk * KT + my_r;
                idx = idx < 0 ? 0 : (idx >= P.Nk ? P.Nk - 1 : idx);
                return pos < count ? idx : -1;
            };
            int idx_next = fetch_idx(0);
            for (int kk = 0; kk < nk; kk++) {
                const int idx_cur = fix_idx(kk, idx_next);
                if (kk + 1 < nk) idx_next = fetch_idx(kk + 1);
                int rowidx[KT / 8];
#pragma unroll
                for (int i = 0; i < KT / 8; i++) rowidx[i] = __shfl_sync(0xffffffffu, idx_cur, (lane & 16) + i);
#pragma unroll
                for (int op = 0; op < 2; op++) {
                    const uint32_t slot = job % NSLOT;
                    mbar_wait(&bar.kv_empty[slot], ((job / NSLOT) & 1) ^ 1);
                    const uint32_t dst0 = sKV + slot * SLOT_BYTES + half_off * (SLOT_BYTES / 2);
                    const __nv_bfloat16* base = (op == 0 ? kb : vb) + chunk * 8;
                    const int64_t rs = op == 0 ? P.ks[2] : P.vs[2];
#pragma unroll
                    for (int i = 0; i < KT / 8; i++) {
                        const int r = rsub + 8 * i;
                        const bool ok = rowidx[i] >= 0;
                        const __nv_bfloat16* src = base + (int64_t)(ok ? rowidx[i] : 0) * rs;
                        if (!CM_DBG(P, 1)) cp_async_16_zfill(dst0 + r * 128 + ((c8 ^ (r & 7)) << 4), src, ok ? 16u : 0u);
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
        setmaxnreg_dec<GEO::reg_mma<QUAD>()>();
        uint32_t job = 0, it = 0, sc0 = 0, sc1 = 0;   // slot jobs, tiles, S/P step counters per block
        const uint32_t idesc_pv0 = umma_idesc_bf16(128, D, 0, 1), idesc_pv1 = umma_idesc_bf16(64, D, 0, 1);
        const uint64_t desc_q = umma_smem_desc(sQ, 16, 1024);                    // K-major A: Q rows
        const uint64_t desc_k = umma_smem_desc(sKV, 16, 1024);                   // K-major B: gathered K rows
        const uint64_t desc_v = umma_smem_desc(sKV, SLOT_BYTES / 2, 1024);       // MN-major B: gathered V rows
        // Every MMA batch is issued under elect.sync: ptxas then emits the UTCHMMAs back to back.  With a lane test
        // (`lane == 0`) it wraps each one in an ELECT / branch loop (~25 cycles per MMA, ~800 per step) during which
        // the tensor pipe drains.
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, it++) {
            const int count = tile_count(P, tile);
            if (count <= 0) { it--; continue; }
            const int nk = (count + KT - 1) / KT;
            auto ncols = [&](int kk) { int v = count - kk * KT; v = v > KT ? KT : v; return (v + 15) & ~15; };

            // The issuing thread is the kernel's scarcest resource: descriptors are built once (their fields
            // are constant but for the 14-bit start address) and each MMA only adds a compile-time offset.
            auto issue_S = [&](int blk, uint32_t slot, int cols) {
                const uint32_t idesc = umma_idesc_bf16(blk ? 64 : 128, cols, 0, 0);
                const uint32_t d = tm + (blk ? TM_S1 : TM_S0);
                const uint64_t ad0 = desc_q + (uint64_t)(blk * ((128 * 128) >> 4));
                const uint64_t bd0 = desc_k + (uint64_t)(slot * (SLOT_BYTES >> 4));
                if (CM_DBG(P, 2)) return;
#pragma unroll
                for (int k16 = 0; k16 < D / 16; k16++) {
                    const uint64_t ad = ad0 + (uint64_t)((((k16 >> 2) * Q_HALF_BYTES) + (k16 & 3) * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)((((k16 >> 2) * (SLOT_BYTES / 2)) + (k16 & 3) * 32) >> 4);
                    umma_ss(d, ad, bd, idesc, k16 > 0);
                }
            };
            auto issue_PV = [&](int blk, uint32_t slot, int cols, bool first) {
                const uint32_t d = tm + (blk ? TM_O1 : TM_O0);
                const uint32_t a = tm + (blk ? TM_S1 : TM_S0);
                const uint64_t bd0 = desc_v + (uint64_t)(slot * (SLOT_BYTES >> 4));
                const uint32_t idesc_pv = blk ? idesc_pv1 : idesc_pv0;
                if (CM_DBG(P, 2)) return;
                if (cols == KT) {
#pragma unroll
                    for (int j = 0; j < KT / 16; j++) umma_ts(d, a + j * 8, bd0 + (uint64_t)(j * (2048 >> 4)), idesc_pv, (!first) || j > 0);
                } else {
                    for (int j = 0; j < cols / 16; j++) umma_ts(d, a + j * 8, bd0 + (uint64_t)(j * (2048 >> 4)), idesc_pv, (!first) || j > 0);
                }
            };

            mbar_wait(&bar.q_full, it & 1);
            // prologue: S0(0), S1(0)
            uint32_t slotK = job % NSLOT;
            mbar_wait(&bar.kv_full[slotK], (job / NSLOT) & 1);
            tc_fence_after_sync();
            if (elect_one()) {
                issue_S(0, slotK, ncols(0)); umma_commit(&bar.s_full[0]);
                issue_S(1, slotK, ncols(0)); umma_commit(&bar.s_full[1]);
                umma_commit(&bar.kv_empty[slotK]);
                if (nk == 1) umma_commit(&bar.q_empty);
            }
            __syncwarp();
            job++;
            for (int kk = 0; kk < nk; kk++) {
                const uint32_t slotV = job % NSLOT;
                mbar_wait(&bar.kv_full[slotV], (job / NSLOT) & 1);
                job++;
                const bool more = kk + 1 < nk;
                uint32_t slotKn = 0;
                if (more) {
                    slotKn = job % NSLOT;
                    mbar_wait(&bar.kv_full[slotKn], (job / NSLOT) & 1);
                    job++;
                }
                // block 0
                mbar_wait(&bar.p_full[0], sc0 & 1); sc0++;
                tc_fence_after_sync();
                if (elect_one()) {
                    issue_PV(0, slotV, ncols(kk), kk == 0);
                    if (more) { issue_S(0, slotKn, ncols(kk + 1)); umma_commit(&bar.s_full[0]); }
                    else umma_commit(&bar.o_full[0]);
                }
                __syncwarp();
                // block 1
                mbar_wait(&bar.p_full[1], sc1 & 1); sc1++;
                tc_fence_after_sync();
                if (elect_one()) {
                    issue_PV(1, slotV, ncols(kk), kk == 0);
                    umma_commit(&bar.kv_empty[slotV]);
                    if (more) {
                        issue_S(1, slotKn, ncols(kk + 1)); umma_commit(&bar.s_full[1]);
                        umma_commit(&bar.kv_empty[slotKn]);
                        if (kk + 2 == nk) umma_commit(&bar.q_empty);
                    } else umma_commit(&bar.o_full[1]);
                }
                __syncwarp();
            }
        }
    }
    // =========================================================================== softmax + epilogue
    else if (warp < GEO::NUM_SOFTMAX_WARPS) {
      // the role body is instantiated once per block (separate code, separate register budgets)
      auto softmax_role = [&](auto BLK) {
        // 0: rows 0-127 (M = 128), 1: rows 128-191 (M = 64); BLKC < 0: one body for both blocks (the !QUAD kernel, whose two
        // blocks run the same code: a smaller kernel)
        constexpr int BLKC = decltype(BLK)::value;
        const int blk = BLKC < 0 ? (warp >> 2) : BLKC;
        const int q4 = warp & 3;
        const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
        const uint32_t tS = tm + (blk ? TM_S1 : TM_S0) + lane_off;
        const uint32_t tO = tm + (blk ? TM_O1 : TM_O0) + lane_off;
        // block 1, softmax: lanes 0-15 of the quadrant through the 16-lane shapes, four threads per row (thread t: rows
        // t/4 and 8 + t/4).  Epilogue of both blocks: 32x32b, thread i on lane i -- block 1's rows sit on lanes 0-15.
        const int c4 = lane & 3;
        const bool active = blk == 0 || lane < 16;
        const int ri = lane & 15;
        const int r_in_tile = blk == 0 ? q4 * 32 + lane : (active ? 128 + q4 * 16 + ri : INACTIVE_ROW);
        uint32_t sc = 0, oc = 0;
        uint32_t ti = 0;                                   // tiles of this CTA so far (phases of c_full / st_full / st_free)
        const uint32_t sw = (uint32_t)(r_in_tile & 7);
        const uint32_t srow = sStage + (uint32_t)(active ? r_in_tile : 0) * 128;     // this row in a d-half of the staging tile (halves 24 KB apart)
        // the staging tile becomes writable: the cached tile has landed in it (fused add-back), or the previous tile's
        // store has read it out
        auto stage_ready = [&]() {
            if (P.cache != nullptr) mbar_wait(&bar.c_full, ti & 1);
            else if (ti > 0) mbar_wait(&bar.st_free, (ti - 1) & 1);
        };

        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ti++) {
            const int count = tile_count(P, tile);
            const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
            const int row = g * QROWS + r_in_tile;
            const bool row_ok = row < P.Nq;
            __nv_bfloat16* orow = P.o + b * P.os[0] + h * P.os[1] + (int64_t)(row_ok ? row : 0) * P.os[2];
            if (count <= 0 && P.stage) {
                // no contribution: the cached tile passes through as it is; a fresh output (or a delta to add) is zero
                stage_ready();
                if (P.cache == nullptr && active) {
#pragma unroll
                    for (int c = 0; c < 16; c++) st_shared_v4(srow + (c >> 3) * (GEO::STAGE_BYTES / 2) + ((((uint32_t)c & 7) ^ sw) << 4), 0, 0, 0, 0);
                    fence_proxy_async_smem();
                }
                if (active) mbar_arrive(&bar.st_full);
                continue;
            }
            if (count <= 0) {
                // (gather-fused epilogue) no contribution: the cached row goes out as it is
                if (row_ok) {
                    const uint4* crow = reinterpret_cast<const uint4*>(P.cache + b * P.cs[0] + h * P.cs[1] + (int64_t)row * P.cs[2]);
#pragma unroll
                    for (int c = 0; c < 16; c++) {
                        const uint4 t = __ldg(crow + c);
                        bcast_st_v4(P, reinterpret_cast<char*>(orow) + c * 16, t.x, t.y, t.z, t.w);
                    }
                }
                continue;
            }
            const int nk = (count + KT - 1) / KT;
            constexpr bool quad = QUAD && BLKC == 1;
            float m_ref = -INFINITY, l_sum = 0.f;          // one thread per row: reference max (raw score units) and row sum
            float m_ref2[2] = {-INFINITY, -INFINITY};      // four threads per row: reference maxima of the thread's two rows,
            float l_part[2] = {0.f, 0.f};                  //          this thread's share of their row sums

            for (int kk = 0; kk < nk; kk++) {
                const int valid = min(KT, count - kk * KT);
                mbar_wait(&bar.s_full[blk], sc & 1); sc++;
                tc_fence_after_sync();
                if (CM_DBG(P, 4)) { l_sum = l_part[0] = l_part[1] = 1.f; }
                else if constexpr (!quad) {
                    if (valid == KT) softmax_step<false>(tS, tO, KT, kk, m_ref, l_sum, active);
                    else if (valid <= 32) softmax_step_narrow(tS, tO, valid, kk, m_ref, l_sum, active);
                    else softmax_step<true>(tS, tO, valid, kk, m_ref, l_sum, active);
                } else {
                    if (valid == KT) softmax_step16<false>(tS, tO, KT, kk, m_ref2, l_part, c4);
                    else if (valid <= 32) softmax_step16_narrow(tS, tO, valid, kk, m_ref2, l_part, c4);
                    else softmax_step16<true>(tS, tO, valid, kk, m_ref2, l_part, c4);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(&bar.p_full[blk]);
            }
            // block 1 row sums: the four threads of a row add up their shares; the epilogue thread of row ri (one per row)
            // fetches the sum from the row's group (rows 0-7 of the window are the groups' "A" rows, 8-15 their "B" rows)
            if constexpr (quad) {
                float la = l_part[0], lb = l_part[1];
                la += __shfl_xor_sync(0xffffffffu, la, 1); lb += __shfl_xor_sync(0xffffffffu, lb, 1);
                la += __shfl_xor_sync(0xffffffffu, la, 2); lb += __shfl_xor_sync(0xffffffffu, lb, 2);
                const float ga = __shfl_sync(0xffffffffu, la, 4 * (ri & 7)), gb = __shfl_sync(0xffffffffu, lb, 4 * (ri & 7));
                l_sum = (ri & 8) ? gb : ga;
            }
            // ---- epilogue: O / l * scale (+ cached o) -> bf16
            if (P.stage) {
                // through shared memory and the TMA: no global load or store on the softmax threads.  The cached tile was
                // TMA-loaded into the staging tile while this tile was being computed; the result overwrites it in place
                // and the TMA thread stores (or reduce-adds) the tile.
                const bool fused = P.cache != nullptr;
                mbar_wait(&bar.o_full[blk], oc & 1); oc++;
                tc_fence_after_sync();
                const float inv = P.o_scale / l_sum;
                stage_ready();
#pragma unroll
                for (int hf = 0; hf < 2; hf++) {
                    uint32_t r[64];
                    tmem_ld32(tO + hf * 64, r);
                    tmem_ld32(tO + hf * 64 + 32, r + 32);
                    tmem_ld_wait();
                    const uint32_t base = srow + hf * (GEO::STAGE_BYTES / 2);
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            w[j] = pack_bf16x2(__uint_as_float(r[8 * c + 2 * j]) * inv, __uint_as_float(r[8 * c + 2 * j + 1]) * inv);
                        const uint32_t a = base + (((uint32_t)c ^ sw) << 4);
                        if (!active) continue;
                        if (fused) {
                            // o = bf16(cache + bf16(delta))
                            const uint4 cv = ld_shared_v4(a);
                            w[0] = pack_bf16x2(bf16_lo(cv.x) + bf16_lo(w[0]), bf16_hi(cv.x) + bf16_hi(w[0]));
                            w[1] = pack_bf16x2(bf16_lo(cv.y) + bf16_lo(w[1]), bf16_hi(cv.y) + bf16_hi(w[1]));
                            w[2] = pack_bf16x2(bf16_lo(cv.z) + bf16_lo(w[2]), bf16_hi(cv.z) + bf16_hi(w[2]));
                            w[3] = pack_bf16x2(bf16_lo(cv.w) + bf16_lo(w[3]), bf16_hi(cv.w) + bf16_hi(w[3]));
                        }
                        st_shared_v4(a, w[0], w[1], w[2], w[3]);
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before_sync();
                if (active) mbar_arrive(&bar.st_full);
                continue;
            }
            // multicast epilogue (cm_csp_attn_add_bcast): direct stores to the NVLS alias of the symmetric output buffer
            // fused add-back: this row of the cached output is requested before the last P.V lands
            const bool fused = P.cache != nullptr;
            const __nv_bfloat16* crow = P.cache + b * P.cs[0] + h * P.cs[1] + (int64_t)(row_ok ? row : 0) * P.cs[2];
            uint32_t cv[4][8];                         // half a row of the cache: 4 sectors of 32 bytes
            auto load_cache_half = [&](int hf) {
                if (P.wide) {
#pragma unroll
                    for (int c = 0; c < 4; c++) ld_global_nc_v8(crow + hf * 64 + c * 16, cv[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const uint4 t = ld_global_nc_v4(crow + hf * 64 + c * 8);
                        cv[c >> 1][(c & 1) * 4 + 0] = t.x; cv[c >> 1][(c & 1) * 4 + 1] = t.y;
                        cv[c >> 1][(c & 1) * 4 + 2] = t.z; cv[c >> 1][(c & 1) * 4 + 3] = t.w;
                    }
                }
            };
            if (fused) load_cache_half(0);
            mbar_wait(&bar.o_full[blk], oc & 1); oc++;
            tc_fence_after_sync();
            const float inv = P.o_scale / l_sum;
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                uint32_t r[64];
                tmem_ld32(tO + hf * 64, r);
                tmem_ld32(tO + hf * 64 + 32, r + 32);
                tmem_ld_wait();
                uint32_t w[32];
#pragma unroll
                for (int j = 0; j < 32; j++) w[j] = pack_bf16x2(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv);
                if (fused) {
                    // o = bf16(cache + bf16(delta)), written out of place: no clone of the cache, no read-modify-write
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            w[8 * c + j] = pack_bf16x2(bf16_lo(cv[c][j]) + bf16_lo(w[8 * c + j]), bf16_hi(cv[c][j]) + bf16_hi(w[8 * c + j]));
                    if (hf == 0) load_cache_half(1);
                }
                {
                    // Fused all-gather: the rows go to the NVLS multicast alias of the symmetric output buffer -- the NVSwitch
                    // replicates them into every GPU -- or, without multicast, to every peer's copy with plain stores over
                    // NVLink, while the other tiles are still being computed.  A
                    // row-per-thread store pattern would put 16-byte packets on NVLink (packet-rate-bound: measured 1.7x
                    // SLOWER than a separate NCCL all-gather at 8 GPUs), so each group of 8 lanes first transposes its
                    // 8 rows x 8 pieces of 16 bytes with three butterfly shuffle stages: store j of a group then
                    // writes 128 CONTIGUOUS bytes of row 8g + j.
                    const int t8 = lane & 7;
#pragma unroll
                    for (int m = 4; m >= 1; m >>= 1) {
                        const bool up = (t8 & m) != 0;
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            if (j & m) continue;
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const uint32_t send = up ? w[4 * j + e] : w[4 * (j | m) + e];
                                const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, m);
                                if (up) w[4 * j + e] = recv; else w[4 * (j | m) + e] = recv;
                            }
                        }
                    }
                    // w[4j .. 4j+3] now holds piece t8 of the row owned by lane (lane & ~7) + j
                    const int row0 = row - t8;                      // first row of this 8-lane group
                    char* gbase = reinterpret_cast<char*>(P.o + b * P.os[0] + h * P.os[1]) + hf * 128 + t8 * 16;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (row0 + j < P.Nq)
                            bcast_st_v4(P, gbase + (int64_t)(row0 + j) * P.os[2] * 2, w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                }
            }
            tc_fence_before_sync();
        }
      };
      if constexpr (!QUAD) { setmaxnreg_inc<GEO::REG_SOFTMAX>(); softmax_role(std::integral_constant<int, -1>{}); }
      else if (warp < 4) { setmaxnreg_inc<GEO::REG_SOFTMAX>(); softmax_role(std::integral_constant<int, 0>{}); }
      else { setmaxnreg_inc<GEO::reg_softmax1<QUAD>()>(); softmax_role(std::integral_constant<int, 1>{}); }
    }
    else {
        // =========================================================================== TMA thread (warp 9); warps 10-11 idle
        setmaxnreg_dec<GEO::reg_mma<QUAD>()>();            // warpgroup-wide: warps 8-11
        if (warp == GEO::WARP_TMA && lane == 0) {
            tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_c); tma_prefetch_desc(&tm_o);
            uint32_t qn = 0;                        // Q tiles loaded so far
            int tq = blockIdx.x;                    // first tile whose Q has not been requested yet
            // request the Q tile of the next tile that has work (tiles with count 0 never touch Q)
            auto next_q = [&]() {
                while (tq < P.num_tiles && tile_count(P, tq) <= 0) tq += gridDim.x;
                if (tq >= P.num_tiles) return;
                const int g = tq % P.G, bh = tq / P.G, h = bh % P.H, b = bh / P.H;
                if (qn > 0) mbar_wait(&bar.q_empty, (qn - 1) & 1);      // the previous Q tile has been consumed by its last S
                mbar_arrive_expect_tx(&bar.q_full, Q_BYTES);
                const TmaCoord c = tma_coords(P.pos[0], 0, g * QROWS, h, b);
                tma_load_4d(sQ, &tm_q, &bar.q_full, 0, c.c[1], c.c[2], c.c[3]);
                tma_load_4d(sQ + Q_HALF_BYTES, &tm_q, &bar.q_full, 64, c.c[1], c.c[2], c.c[3]);
                qn++;
                tq += gridDim.x;
            };
            next_q();
            uint32_t ti = 0;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ti++) {
                const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
                if (P.stage && P.cache != nullptr) {
                    // the staging tile is free (the previous store has been read out below): fetch this tile's cached rows
                    mbar_arrive_expect_tx(&bar.c_full, GEO::STAGE_BYTES);
                    const TmaCoord c = tma_coords(P.pos[1], 0, g * QROWS, h, b);
                    tma_load_4d(sStage, &tm_c, &bar.c_full, 0, c.c[1], c.c[2], c.c[3]);
                    tma_load_4d(sStage + GEO::STAGE_BYTES / 2, &tm_c, &bar.c_full, 64, c.c[1], c.c[2], c.c[3]);
                }
                if (tile_count(P, tile) > 0) next_q();                 // Q of the tile after this one, as soon as this one's last S is issued
                if (P.stage) {
                    mbar_wait(&bar.st_full, ti & 1);                    // the epilogue threads have written (and fenced) the output tile
                    const TmaCoord c = tma_coords(P.pos[2], 0, g * QROWS, h, b);
                    if (P.cache == nullptr && P.accumulate) {
                        tma_reduce_add_4d(&tm_o, sStage, 0, c.c[1], c.c[2], c.c[3]);
                        tma_reduce_add_4d(&tm_o, sStage + GEO::STAGE_BYTES / 2, 64, c.c[1], c.c[2], c.c[3]);
                    } else {
                        tma_store_4d(&tm_o, sStage, 0, c.c[1], c.c[2], c.c[3]);
                        tma_store_4d(&tm_o, sStage + GEO::STAGE_BYTES / 2, 64, c.c[1], c.c[2], c.c[3]);
                    }
                    bulk_commit();
                    bulk_wait_read<0>();
                    mbar_arrive(&bar.st_free);
                }
            }
            bulk_wait<0>();
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tm, 512);
}

}  // namespace attn
}  // namespace cm

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
using namespace cm;
using namespace cm::attn;

static int launch_attn(Params& P, cudaStream_t stream) {
    // long sequences (the video shapes) take the four-threads-per-row form of block 1, short ones the smaller kernel
    const bool quad = P.Nk >= 16384;
    static unsigned long long configured[2] = {0, 0};
    int rc = quad ? opt_in_dynamic_smem(configured[1], reinterpret_cast<const void*>(attn_kernel<true>), GEO::SMEM_BYTES)
                  : opt_in_dynamic_smem(configured[0], reinterpret_cast<const void*>(attn_kernel<false>), GEO::SMEM_BYTES);
    if (rc) return rc;
    // tensor maps (cached per argument set): Q tile in, cached tile in, output tile out; boxes of 192 rows x 64 columns
    CUtensorMap mq, mc, mo;
    rc = encode_tmap_bhnd(&mq, P.q, P.B, P.H, P.Nq, P.qs, QG, P.pos[0]);
    if (!rc) rc = encode_tmap_bhnd(&mo, P.o, P.B, P.H, P.Nq, P.os, QG, P.pos[2]);
    if (!rc && P.cache) rc = encode_tmap_bhnd(&mc, P.cache, P.B, P.H, P.Nq, P.cs, QG, P.pos[1]);
    if (rc) return rc;
    if (!P.cache) { mc = mo; for (int i = 0; i < 3; i++) P.pos[1][i] = P.pos[2][i]; }
    int grid = P.num_tiles < sm_count() ? P.num_tiles : sm_count();
    if (quad) attn_kernel<true><<<grid, GEO::NUM_THREADS, GEO::SMEM_BYTES, stream>>>(mq, mc, mo, P);
    else attn_kernel<false><<<grid, GEO::NUM_THREADS, GEO::SMEM_BYTES, stream>>>(mq, mc, mo, P);
    return (int)cudaGetLastError();
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool strides_ok(const int64_t s[3]) { return s[0] % 8 == 0 && s[1] % 8 == 0 && s[2] % 8 == 0; }

static int csp_attn_impl(const void* q, const void* k, const void* v, const void* cache, void* o, const int32_t* indices,
                         const int32_t* counts, int B, int H, int Nq, int Nk, const int64_t q_strides[3],
                         const int64_t k_strides[3], const int64_t v_strides[3], const int64_t c_strides[3],
                         const int64_t o_strides[3], int64_t idx_row_stride, int o_scale, int accumulate, void* stream,
                         int64_t mc_delta = 0, const int64_t* peer_delta = nullptr, int n_peers = 0) {
    if (B < 0 || H < 0 || Nq < 0 || Nk <= 0 || idx_row_stride <= 0) return CM_EINVAL;
    if (o_scale != 1 && o_scale != -1) return CM_EINVAL;
    if ((int64_t)B * H * Nq == 0) return CM_OK;
    if (!q || !k || !v || !o || !indices || !counts) return CM_EINVAL;
    if (!aligned16(q) || !aligned16(k) || !aligned16(v) || !aligned16(o)) return CM_EALIGN;
    if (!strides_ok(q_strides) || !strides_ok(k_strides) || !strides_ok(v_strides) || !strides_ok(o_strides))
        return CM_EALIGN;
    if (cache && (!aligned16(cache) || !strides_ok(c_strides) || cache == o)) return cache == o ? CM_EINVAL : CM_EALIGN;
    if (!is_sm100()) return CM_EARCH;
    Params P{};
    P.q = (const __nv_bfloat16*)q; P.k = (const __nv_bfloat16*)k; P.v = (const __nv_bfloat16*)v;
    P.o = (__nv_bfloat16*)o;
    P.indices = indices; P.counts = counts;
    P.cache = (const __nv_bfloat16*)cache;
    P.B = B; P.H = H; P.Nq = Nq; P.Nk = Nk; P.G = (Nq + QG - 1) / QG;
    for (int i = 0; i < 3; i++) {
        P.qs[i] = q_strides[i]; P.ks[i] = k_strides[i]; P.vs[i] = v_strides[i]; P.os[i] = o_strides[i];
        P.cs[i] = cache ? c_strides[i] : 0;
    }
    P.idx_row_stride = idx_row_stride;
    P.o_scale = (float)o_scale;
    P.accumulate = accumulate ? 1 : 0;
    auto wide_ok = [](const void* p, const int64_t st[3]) {
        return (reinterpret_cast<uintptr_t>(p) & 31) == 0 && st[0] % 16 == 0 && st[1] % 16 == 0 && st[2] % 16 == 0;
    };
    P.wide = wide_ok(o, o_strides) && (!cache || wide_ok(cache, c_strides)) ? 1 : 0;
    if (mc_delta != 0 && (!cache || (mc_delta & 15))) return CM_EINVAL;
    P.mc_delta = mc_delta;
    if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && (!cache || !peer_delta || mc_delta != 0))) return CM_EINVAL;
    P.n_peers = n_peers;
    for (int i = 0; i < n_peers; i++) { if (peer_delta[i] & 15) return CM_EINVAL; P.peer_delta[i] = peer_delta[i]; }
    P.stage = (mc_delta == 0 && n_peers == 0) ? 1 : 0;
    int64_t tiles = (int64_t)B * H * P.G;
    if (tiles > 2147483647ll) return CM_EINVAL;
    P.num_tiles = (int)tiles;
    P.dbg = debug_flags();
    return launch_attn(P, (cudaStream_t)stream);
}

extern "C" int cm_csp_attn(const void* q, const void* k, const void* v, void* o, const int32_t* indices,
                           const int32_t* counts, int B, int H, int Nq, int Nk, const int64_t q_strides[3],
                           const int64_t k_strides[3], const int64_t v_strides[3], const int64_t o_strides[3],
                           int64_t idx_row_stride, int o_scale, int accumulate, void* stream) {
    return csp_attn_impl(q, k, v, nullptr, o, indices, counts, B, H, Nq, Nk, q_strides, k_strides, v_strides, nullptr,
                         o_strides, idx_row_stride, o_scale, accumulate, stream);
}

extern "C" int cm_csp_attn_add_bcast(const void* q, const void* k, const void* v, const void* cache, void* o_local,
                                     int64_t multicast_delta_bytes, const int32_t* indices, const int32_t* counts, int B,
                                     int H, int Nq, int Nk, const int64_t q_strides[3], const int64_t k_strides[3],
                                     const int64_t v_strides[3], const int64_t cache_strides[3],
                                     const int64_t o_strides[3], int64_t idx_row_stride, int o_scale, void* stream) {
    if (!cache || !cache_strides || multicast_delta_bytes == 0) return CM_EINVAL;
    return csp_attn_impl(q, k, v, cache, o_local, indices, counts, B, H, Nq, Nk, q_strides, k_strides, v_strides,
                         cache_strides, o_strides, idx_row_stride, o_scale, 0, stream, multicast_delta_bytes);
}

extern "C" int cm_csp_attn_add_peers(const void* q, const void* k, const void* v, const void* cache, void* o_local,
                                     const int64_t* peer_delta_bytes, int n_peers, const int32_t* indices, const int32_t* counts,
                                     int B, int H, int Nq, int Nk, const int64_t q_strides[3], const int64_t k_strides[3],
                                     const int64_t v_strides[3], const int64_t cache_strides[3],
                                     const int64_t o_strides[3], int64_t idx_row_stride, int o_scale, void* stream) {
    if (!cache || !cache_strides || !peer_delta_bytes || n_peers <= 0) return CM_EINVAL;
    return csp_attn_impl(q, k, v, cache, o_local, indices, counts, B, H, Nq, Nk, q_strides, k_strides, v_strides,
                         cache_strides, o_strides, idx_row_stride, o_scale, 0, stream, 0, peer_delta_bytes, n_peers);
}

extern "C" int cm_csp_attn_add(const void* q, const void* k, const void* v, const void* cache, void* o,
                               const int32_t* indices, const int32_t* counts, int B, int H, int Nq, int Nk,
                               const int64_t q_strides[3], const int64_t k_strides[3], const int64_t v_strides[3],
                               const int64_t cache_strides[3], const int64_t o_strides[3], int64_t idx_row_stride,
                               int o_scale, void* stream) {
    if (!cache || !cache_strides) return CM_EINVAL;
    return csp_attn_impl(q, k, v, cache, o, indices, counts, B, H, Nq, Nk, q_strides, k_strides, v_strides, cache_strides,
                         o_strides, idx_row_stride, o_scale, 0, stream);
}

// cm_dense_attn (contiguous [B,H,N,128] operands) is the strided one-pass kernel of dense_attn.cu with packed strides.
extern "C" int cm_dense_attn_strided(const void* q, const void* k, const void* v, void* o, float* l, void* cs, const float* p,
                                     int B, int H, int Nq, int Nk, const int64_t q_strides[3], const int64_t k_strides[3],
                                     const int64_t v_strides[3], const int64_t o_strides[3], int64_t cs_row_stride,
                                     void* stream);

extern "C" int cm_dense_attn(const void* q, const void* k, const void* v, void* o, float* l, void* cs,
                             const float* p, int B, int H, int Nq, int Nk, int64_t cs_row_stride, void* stream) {
    const int64_t qst[3] = {(int64_t)H * Nq * D, (int64_t)Nq * D, D};
    const int64_t kst[3] = {(int64_t)H * Nk * D, (int64_t)Nk * D, D};
    return cm_dense_attn_strided(q, k, v, o, l, cs, p, B, H, Nq, Nk, qst, kst, kst, qst, cs_row_stride, stream);
}
