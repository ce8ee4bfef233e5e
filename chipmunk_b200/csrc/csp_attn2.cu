// Column-sparse delta attention on a CTA PAIR (tcgen05 cta_group::2) for sm_100a.
//
// Same operator as csp_attn.cu (`cm_csp_attn`), second-generation schedule.  One cluster of two
// CTAs (= two SMs of a TPC) owns a (b, h, 192-query group) tile:
//   * CTA r holds query rows 96r .. 96r+95 of the group on TMEM lanes 0-95 (M = 256 for the pair);
//   * the gathered K/V tile of a 128-key step is SHARED by the pair: CTA r gathers keys
//     64r..64r+63 (all 128 head dims) for S = Q K^T, whose B operand is split along N (keys), and
//     head dims 64r..64r+63 of all 128 keys for O += P V, whose B operand is split along N (dims).
//     Every SM therefore copies 32 KB per step instead of 64 KB -- the per-SM cp.async rate
//     (~30 B/clk, tests/probes/probe_gather4.cu) is what bounds the 1-CTA kernel;
//   * S is triple-buffered in TMEM (3 x 128 columns + 128 for O), so the tensor pipe runs Q K^T two
//     steps ahead of the softmax instead of ping-ponging with it.
// Roles per CTA (384 threads): warps 0-2 softmax/epilogue (one thread per query row), 4-7 gather
// producers, warp 8 = MMA issuer in the leader CTA / completion relay in the peer CTA.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/chipmunk_b200.h"
#include "attn_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace cm {
namespace attn2 {

using namespace cm::attn;

constexpr int ROWS = 96;                  // query rows per CTA
constexpr int NSLOT = 8;                  // 16 KB K/V slots per CTA
constexpr int SLOT_BYTES = 16384;         // K: 2 d-halves x [64 keys][128 B];  V: [128 keys][128 B] (this CTA's d-half)
constexpr int Q_BYTES = 2 * 128 * 128;    // 2 d-halves x [128 rows][128 B] (rows 96-127 unused)
constexpr int SMEM_BYTES = 2 * Q_BYTES + NSLOT * SLOT_BYTES + 1024;
constexpr int NUM_THREADS = 384;
constexpr int WARP_PROD0 = 4, WARP_MMA = 8, NUM_PROD = 128;
constexpr int NSB = 3;                    // S buffers in TMEM
constexpr int IDX_RING = 8;               // steps of key indices kept in shared memory by the index-prefetch warp
constexpr uint32_t TM_O = 384;

struct Params {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;
    const int32_t* indices;
    const int32_t* counts;
    int B, H, Nq, Nk, G;
    int64_t qs[3], ks[3], vs[3], os[3];
    int64_t idx_row_stride;
    float o_scale;
    int accumulate;
    int num_tiles;
    int dbg;
};

struct __align__(8) Barriers {
    uint64_t q_full[2], q_empty[2], q_peer[2];
    uint64_t kv_full[NSLOT], kv_empty[NSLOT], kv_peer[NSLOT];
    uint64_t s_full[NSB], p_full[NSB];
    uint64_t idx_full[IDX_RING], idx_empty[IDX_RING];
    uint64_t pv_done[2];      // P.V of step gs arrives on [gs & 1]; two barriers so a waiter can never alias a phase
};

// ---------------------------------------------------------------- cluster / cta_group::2 PTX
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t ncluster_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// arrive (once) on the barrier at the same shared-memory offset in BOTH CTAs when all prior MMAs are done
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the LEADER CTA's copy of `bar`
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;\n" : "=r"(ra) : "r"(smem_u32(bar)));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(ra) : "memory");
}

__device__ __forceinline__ int tile_count(const Params& P, int tile) {
    int c = __ldg(P.counts + tile);
    c = c < 0 ? 0 : c;
    return c > (int)P.idx_row_stride ? (int)P.idx_row_stride : c;
}

// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) attn2_kernel(const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) int s_idx[IDX_RING][KT];     // key indices of the next steps, prefetched IDX_AHEAD steps ahead with cp.async

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = sbase;
    const uint32_t sKV = sbase + 2 * Q_BYTES;
    const int tile0 = (int)cluster_id_x(), tile_stride = (int)ncluster_x();

    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&bar.q_full[i], NUM_PROD); mbar_init(&bar.q_empty[i], 1); mbar_init(&bar.q_peer[i], 1); }
        for (int i = 0; i < NSLOT; i++) { mbar_init(&bar.kv_full[i], NUM_PROD); mbar_init(&bar.kv_empty[i], 1); mbar_init(&bar.kv_peer[i], 1); }
        for (int i = 0; i < NSB; i++) { mbar_init(&bar.s_full[i], 1); mbar_init(&bar.p_full[i], 6); }
        mbar_init(&bar.pv_done[0], 1); mbar_init(&bar.pv_done[1], 1);
        for (int i = 0; i < IDX_RING; i++) { mbar_init(&bar.idx_full[i], 32); mbar_init(&bar.idx_empty[i], NUM_PROD); }
        fence_mbar_init();
    }
    if (warp == WARP_MMA) tmem_alloc2(&tmem_base_s, 512);
    tc_fence_before_sync();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;

    // =========================================================================== producers
    if (warp >= WARP_PROD0 && warp < WARP_PROD0 + 4) {
        setmaxnreg_dec<80>();
        const int pt = tid - WARP_PROD0 * 32;        // 0..127
        uint32_t job = 0, it = 0, gstep = 0;
        for (int tile = tile0; tile < P.num_tiles; tile += tile_stride) {
            const int count = tile_count(P, tile);
            if (count <= 0) continue;
            const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
            // ---- this CTA's 96 query rows, zero-filled past Nq
            const uint32_t qb = it & 1;
            mbar_wait(&bar.q_empty[qb], ((it >> 1) & 1) ^ 1);
            {
                const __nv_bfloat16* qbase = P.q + b * P.qs[0] + h * P.qs[1];
                const int chunk = pt & 15, rsub = pt >> 4;
                const uint32_t dq = sQ + qb * Q_BYTES + (uint32_t)(chunk >> 3) * (Q_BYTES / 2);
#pragma unroll 4
                for (int i = 0; i < ROWS / 8; i++) {
                    const int r = rsub + 8 * i;
                    const int row = g * QG + (int)rank * ROWS + r;
                    const bool ok = row < P.Nq;
                    cp_async_16_zfill(dq + r * 128 + (((chunk & 7) ^ (r & 7)) << 4),
                                      qbase + (int64_t)(ok ? row : 0) * P.qs[2] + chunk * 8, ok ? 16u : 0u);
                }
                cp_async_mbar_arrive_noinc(&bar.q_full[qb]);
            }
            it++;
            const __nv_bfloat16* kb = P.k + b * P.ks[0] + h * P.ks[1];
            const __nv_bfloat16* vb = P.v + b * P.vs[0] + h * P.vs[1];
            const int32_t* ip = P.indices + (int64_t)tile * P.idx_row_stride;
            const int nk = (count + KT - 1) / KT;
            for (int kk = 0; kk < nk; kk++) {
                const int valid = min(KT, count - kk * KT);
                const int cols = (valid + 15) & ~15;
                const int half = cols >> 1;                  // keys per CTA for S
                // key indices of this step: staged in shared memory by the index-prefetch warp
                const uint32_t ring = gstep % IDX_RING;
                mbar_wait(&bar.idx_full[ring], (gstep / IDX_RING) & 1);
                const int* si = s_idx[ring];
                int kidx[8], vidx[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int rk = (pt >> 4) + 8 * i, rv = (pt >> 3) + 16 * i;
                    const int pk = (int)rank * half + rk;
                    int a = (rk < half && kk * KT + pk < count) ? si[pk] : -1;
                    int c = (kk * KT + rv < count) ? si[rv] : -1;
                    kidx[i] = a >= P.Nk ? P.Nk - 1 : a;
                    vidx[i] = c >= P.Nk ? P.Nk - 1 : c;
                }
                mbar_arrive(&bar.idx_empty[ring]);
                gstep++;
                // ---- K: keys rank*half + r, r < half, full 256-byte rows
                {
                    const uint32_t slot = job % NSLOT;
                    mbar_wait(&bar.kv_empty[slot], ((job / NSLOT) & 1) ^ 1);
                    const int chunk = pt & 15, rsub = pt >> 4;
                    const uint32_t dst = sKV + slot * SLOT_BYTES + (uint32_t)(chunk >> 3) * (SLOT_BYTES / 2);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = rsub + 8 * i;
                        if (r < half && !(P.dbg & 1)) {
                            const int idx = kidx[i];
                            cp_async_16_zfill(dst + r * 128 + (((chunk & 7) ^ (r & 7)) << 4),
                                              kb + (int64_t)(idx < 0 ? 0 : idx) * P.ks[2] + chunk * 8, idx < 0 ? 0u : 16u);
                        }
                    }
                    cp_async_mbar_arrive_noinc(&bar.kv_full[slot]);
                    job++;
                }
                // ---- V: all keys r < cols, this CTA's 128-byte half of the row
                {
                    const uint32_t slot = job % NSLOT;
                    mbar_wait(&bar.kv_empty[slot], ((job / NSLOT) & 1) ^ 1);
                    const int chunk = pt & 7, rsub = pt >> 3;
                    const uint32_t dst = sKV + slot * SLOT_BYTES;
                    const __nv_bfloat16* vsrc = vb + (int)rank * 64 + chunk * 8;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int r = rsub + 16 * i;
                        if (r < cols && !(P.dbg & 1)) {
                            const int idx = vidx[i];
                            cp_async_16_zfill(dst + r * 128 + ((chunk ^ (r & 7)) << 4),
                                              vsrc + (int64_t)(idx < 0 ? 0 : idx) * P.vs[2], idx < 0 ? 0u : 16u);
                        }
                    }
                    cp_async_mbar_arrive_noinc(&bar.kv_full[slot]);
                    job++;
                }
            }
        }
        cp_async_wait_all();
    }
    // =========================================================================== MMA issuer (leader) / relay (peer)
    else if (warp == WARP_MMA) {
        setmaxnreg_dec<80>();
        if (rank != 0) {
            // relay: tell the leader when THIS CTA's halves of Q / K / V have landed
            uint32_t job = 0, it = 0;
            for (int tile = tile0; tile < P.num_tiles; tile += tile_stride) {
                const int count = tile_count(P, tile);
                if (count <= 0) continue;
                const uint32_t qb = it & 1;
                mbar_wait(&bar.q_full[qb], (it >> 1) & 1);
                fence_proxy_async_smem();
                if (lane == 0) mbar_arrive_leader(&bar.q_peer[qb]);
                it++;
                const int njobs = 2 * ((count + KT - 1) / KT);
                for (int j = 0; j < njobs; j++, job++) {
                    const uint32_t slot = job % NSLOT;
                    mbar_wait(&bar.kv_full[slot], (job / NSLOT) & 1);
                    fence_proxy_async_smem();
                    if (lane == 0) mbar_arrive_leader(&bar.kv_peer[slot]);
                }
            }
        } else {
            uint32_t jobbase = 0, it = 0, gs = 0;      // slot jobs before this tile, tiles, global step counter
            const uint32_t idesc_pv = umma_idesc_bf16(256, D, 0, 1);
            for (int tile = tile0; tile < P.num_tiles; tile += tile_stride) {
                const int count = tile_count(P, tile);
                if (count <= 0) continue;
                const int nk = (count + KT - 1) / KT;
                const uint32_t qb = it & 1;
                auto ncols = [&](int kk) { int v = count - kk * KT; v = v > KT ? KT : v; return (v + 15) & ~15; };
                auto wait_slot = [&](uint32_t job) {
                    const uint32_t slot = job % NSLOT, par = (job / NSLOT) & 1;
                    mbar_wait(&bar.kv_full[slot], par);
                    mbar_wait(&bar.kv_peer[slot], par);
                    return slot;
                };
                auto issue_S = [&](int kk) {
                    const uint32_t slot = wait_slot(jobbase + 2 * kk);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t idesc = umma_idesc_bf16(256, ncols(kk), 0, 0);
                        const uint32_t d = tm + ((gs + kk) % NSB) * 128;
#pragma unroll
                        for (int k16 = 0; k16 < ((P.dbg & 2) ? 0 : D / 16); k16++) {
                            const uint64_t ad = umma_smem_desc(sQ + qb * Q_BYTES + (k16 >> 2) * (Q_BYTES / 2) + (k16 & 3) * 32, 16, 1024);
                            const uint64_t bd = umma_smem_desc(sKV + slot * SLOT_BYTES + (k16 >> 2) * (SLOT_BYTES / 2) + (k16 & 3) * 32, 16, 1024);
                            umma2_ss(d, ad, bd, idesc, k16 > 0);
                        }
                        umma2_commit(&bar.s_full[(gs + kk) % NSB]);
                        umma2_commit(&bar.kv_empty[slot]);
                        if (kk == nk - 1) umma2_commit(&bar.q_empty[qb]);
                    }
                    __syncwarp();
                };
                mbar_wait(&bar.q_full[qb], (it >> 1) & 1);
                mbar_wait(&bar.q_peer[qb], (it >> 1) & 1);
                issue_S(0);
                if (nk > 1) issue_S(1);
                for (int kk = 0; kk < nk; kk++) {
                    if (kk + 2 < nk) issue_S(kk + 2);
                    const uint32_t slot = wait_slot(jobbase + 2 * kk + 1);
                    const uint32_t sb = (gs + kk) % NSB;
                    mbar_wait(&bar.p_full[sb], ((gs + kk) / NSB) & 1);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t a = tm + sb * 128, d = tm + TM_O;
                        const int steps = (P.dbg & 2) ? 0 : ncols(kk) / 16;
                        for (int j = 0; j < steps; j++) {
                            const uint64_t bd = umma_smem_desc(sKV + slot * SLOT_BYTES + j * 2048, SLOT_BYTES, 1024);
                            umma2_ts(d, a + j * 8, bd, idesc_pv, (kk | j) != 0);
                        }
                        umma2_commit(&bar.kv_empty[slot]);
                        umma2_commit(&bar.pv_done[(gs + kk) & 1]);
                    }
                    __syncwarp();
                }
                gs += nk;
                jobbase += 2 * nk;
                it++;
            }
        }
    }
    // =========================================================================== softmax + epilogue
    else if (warp < 3) {
        setmaxnreg_inc<208>();
        const int r_in_cta = warp * 32 + lane;                       // 0..95
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        const uint32_t tO = tm + TM_O + lane_off;
        uint32_t gs = 0;
        for (int tile = tile0; tile < P.num_tiles; tile += tile_stride) {
            const int count = tile_count(P, tile);
            const int g = tile % P.G, bh = tile / P.G, h = bh % P.H, b = bh / P.H;
            const int row = g * QG + (int)rank * ROWS + r_in_cta;
            const bool row_ok = row < P.Nq;
            __nv_bfloat16* orow = P.o + b * P.os[0] + h * P.os[1] + (int64_t)(row_ok ? row : 0) * P.os[2];
            if (count <= 0) {
                if (!P.accumulate && row_ok) {
#pragma unroll
                    for (int c = 0; c < 16; c++) reinterpret_cast<uint4*>(orow)[c] = make_uint4(0, 0, 0, 0);
                }
                continue;
            }
            const int nk = (count + KT - 1) / KT;
            float m_ref = -INFINITY, l_sum = 0.f;
            for (int kk = 0; kk < nk; kk++, gs++) {
                const int valid = min(KT, count - kk * KT);
                const uint32_t sb = gs % NSB;
                mbar_wait(&bar.s_full[sb], (gs / NSB) & 1);
                tc_fence_after_sync();
                const uint32_t tS = tm + sb * 128 + lane_off;
                // P.V of the previous step (needed only if O must be rescaled): barrier [(gs-1)&1], its ((gs-1)>>1)-th phase
                uint64_t* pvb = &bar.pv_done[(gs - 1) & 1];
                const uint32_t pvp = ((gs - 1) >> 1) & 1;
                if (P.dbg & 4) l_sum = 1.f;
                else if (valid == KT) softmax_step<false>(tS, tO, KT, kk, m_ref, l_sum, pvb, pvp);
                else softmax_step<true>(tS, tO, valid, kk, m_ref, l_sum, pvb, pvp);
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive(&bar.p_full[sb]);
                    else mbar_arrive_leader(&bar.p_full[sb]);
                }
            }
            // ---- epilogue: wait for the last P.V, then O / l * scale (+ cached o) -> bf16
            mbar_wait(&bar.pv_done[(gs - 1) & 1], ((gs - 1) >> 1) & 1);
            tc_fence_after_sync();
            const float inv = P.o_scale / l_sum;
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
                        // the delta add-back: o = bf16(o + delta) as a 16-byte reduction at the L2 (the reference
                        // uses a TMA reduce-add, csp_attn.cu:300); plain store for csp_128_attn / dense
                        if (P.accumulate) red_add_bf16x8(dst, w[0], w[1], w[2], w[3]);
                        else *dst = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
            }
            tc_fence_before_sync();
        }
    } else if (warp == 3) {
        setmaxnreg_inc<208>();      // (warpgroup-wide) -- this warp owns no query rows
        // ======================================================================= index prefetch
        // Streams the tile's key indices into the shared-memory ring, up to IDX_RING steps ahead of the
        // gather warps and decoupled from their cp.async completion tracking.
        uint32_t gstep = 0;
        for (int tile = tile0; tile < P.num_tiles; tile += tile_stride) {
            const int count = tile_count(P, tile);
            if (count <= 0) continue;
            const int32_t* ip = P.indices + (int64_t)tile * P.idx_row_stride;
            const bool vec = ((reinterpret_cast<uintptr_t>(ip) & 15) == 0) && (P.idx_row_stride % 4 == 0);
            const int nk = (count + KT - 1) / KT;
            for (int kk = 0; kk < nk; kk++, gstep++) {
                const uint32_t ring = gstep % IDX_RING;
                mbar_wait(&bar.idx_empty[ring], ((gstep / IDX_RING) & 1) ^ 1);
                const int pos = kk * KT + lane * 4;
                const uint32_t dst = smem_u32(&s_idx[ring][lane * 4]);
                if (vec && pos + 4 <= (int)P.idx_row_stride) {
                    cp_async_16(dst, ip + pos);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (pos + j < (int)P.idx_row_stride)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst + 4 * j), "l"(ip + pos + j) : "memory");
                }
                cp_async_mbar_arrive_noinc(&bar.idx_full[ring]);
            }
        }
        cp_async_wait_all();
    } else {
        setmaxnreg_dec<80>();       // warps 9-11: idle
    }

    tc_fence_before_sync();
    cluster_sync_all();             // the peer's shared memory must outlive the leader's last MMA
    if (warp == WARP_MMA) tmem_dealloc2(tm, 512);
}

}  // namespace attn2
}  // namespace cm

// ------------------------------------------------------------------------------------------
namespace cm { namespace attn2 {
int launch(const void* q, const void* k, const void* v, void* o, const int32_t* indices, const int32_t* counts, int B,
           int H, int Nq, int Nk, const int64_t qs[3], const int64_t ks[3], const int64_t vs[3], const int64_t os[3],
           int64_t idx_row_stride, int o_scale, int accumulate, cudaStream_t stream) {
    static unsigned long long configured = 0;
    if (int rc0 = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(attn2_kernel), SMEM_BYTES)) return rc0;
    Params P{};
    P.q = (const __nv_bfloat16*)q; P.k = (const __nv_bfloat16*)k; P.v = (const __nv_bfloat16*)v; P.o = (__nv_bfloat16*)o;
    P.indices = indices; P.counts = counts;
    P.B = B; P.H = H; P.Nq = Nq; P.Nk = Nk; P.G = (Nq + QG - 1) / QG;
    for (int i = 0; i < 3; i++) { P.qs[i] = qs[i]; P.ks[i] = ks[i]; P.vs[i] = vs[i]; P.os[i] = os[i]; }
    P.idx_row_stride = idx_row_stride;
    P.o_scale = (float)o_scale;
    P.accumulate = accumulate ? 1 : 0;
    const int64_t tiles = (int64_t)B * H * P.G;
    if (tiles > 2147483647ll) return CM_EINVAL;
    P.num_tiles = (int)tiles;
    P.dbg = debug_flags();
    const int max_clusters = sm_count() / 2;
    const int clusters = P.num_tiles < max_clusters ? P.num_tiles : max_clusters;
    attn2_kernel<<<2 * clusters, NUM_THREADS, SMEM_BYTES, stream>>>(P);
    return (int)cudaGetLastError();
}
} }
