// Column sums of the attention probabilities for mask selection (dense "full" steps).
//   cs[b,h,g,j] = sum_{i in 192-row group g} exp(q_i . k_j / sqrt(128)) * p_i ,  p = previous step's l
// Replaces the colsum half of csrc/attn/dense_colsum_attn.cu:267-333 of the reference, which
// reduces bf16 partials across 12 warps with shared-memory atomics.  Here the product is
// computed TRANSPOSED on the tensor pipe, S^T = K_tile Q_g^T (M = 128 keys on the TMEM lanes,
// N = 192 queries on the columns), so the sum over the group's queries is a per-thread loop
// over TMEM columns: no atomics, no shuffles, fp32 accumulation, one bf16 rounding.
//
//   warp 5      TMA: Q group tile once per (b,h,g) (double-buffered), K tiles through a 3-stage ring
//   warp 4      MMA issuer: tcgen05.mma M=128 N=192 K=128 into one of two TMEM accumulators
//   warps 0-3   one thread per key, query columns 0-95:  sum_i exp2(s*c + log2 p_i), adds the partner's partial, bf16 store
//   warps 8-11  the same keys, query columns 96-191 (two exp warps per SM sub-partition: a single warp can issue
//               a MUFU.EX2 only every ~8.8 cycles and nothing else meanwhile; two interleave and fill the MUFU pipe)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace cm {
namespace attn {

namespace cs {
#ifndef CM_COLSUM_POLY
#define CM_COLSUM_POLY 2      // every CM_COLSUM_POLY-th group of 4 scores is exponentiated on the FMA pipe (0: all on the MUFU)
#endif
constexpr int D = 128, QG = 192, KT = 128;

// exp2 of two fp32 on the FMA / ALU pipes: x = n + f (round to nearest with the 1.5*2^23 trick), 2^f by a degree-3
// Chebyshev fit on [-0.5, 0.5] (relative error 1.0e-4; the sums feed a top-k selection and are rounded to bf16),
// 2^n added into the exponent field.  This kernel IS MUFU-throughput-bound (two exp warps per sub-partition and
// nothing else on the FMA pipe), unlike the attention softmax where the polynomial costs what the MUFU it replaces costs.
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
constexpr int Q_BYTES = 2 * QG * 128;       // 49152: two 64-wide d-halves
constexpr int K_BYTES = 2 * KT * 128;       // 32768
constexpr int KSTAGES = 3;
constexpr int SMEM_BYTES = 2 * Q_BYTES + KSTAGES * K_BYTES + 1024;
constexpr int NUM_THREADS = 384;    // warps 0-3 exp A | 4 MMA | 5 TMA | 6-7 idle | 8-11 exp B
constexpr float SCALE_LOG2 = 0.08838834764f * 1.44269504089f;

struct Params {
    const float* p;            // [B*H, Nq]
    __nv_bfloat16* cs;         // [B*H*G, cs_stride]
    int BH, Nq, Nk, G;
    int64_t cs_stride;
    int num_tiles, n_kt;
};

struct __align__(8) Barriers {
    uint64_t q_full[2], q_empty[2];
    uint64_t k_full[KSTAGES], k_empty[KSTAGES];
    uint64_t acc_full[2], acc_empty[2];
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
colsum_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float s_lp[2][QG];     // log2(p_i) of the tile's query rows
    __shared__ float s_part[2][KT];                 // partial sums of the second exp group

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sQ = sbase, sK = sbase + 2 * Q_BYTES;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&bar.q_full[i], 1); mbar_init(&bar.q_empty[i], 1);
            mbar_init(&bar.acc_full[i], 1); mbar_init(&bar.acc_empty[i], 256);
        }
        for (int i = 0; i < KSTAGES; i++) { mbar_init(&bar.k_full[i], 1); mbar_init(&bar.k_empty[i], 1); }
        fence_mbar_init();
    }
    if (warp == 4) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (warp == 5 && lane == 0) { tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;

    if (warp == 5) {
        // ======================================================================= TMA loader
        if (lane == 0) {
            uint32_t tcount = 0, kc = 0;
            for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, tcount++) {
                const int g = tile % P.G, bh = tile / P.G;
                const uint32_t qb = tcount & 1;
                mbar_wait(&bar.q_empty[qb], ((tcount >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar.q_full[qb], Q_BYTES);
                const int qrow = bh * P.Nq + g * QG;
                tma_load_2d(sQ + qb * Q_BYTES, &tmap_q, &bar.q_full[qb], 0, qrow);
                tma_load_2d(sQ + qb * Q_BYTES + Q_BYTES / 2, &tmap_q, &bar.q_full[qb], 64, qrow);
                for (int kt = 0; kt < P.n_kt; kt++, kc++) {
                    const uint32_t s = kc % KSTAGES;
                    mbar_wait(&bar.k_empty[s], ((kc / KSTAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar.k_full[s], K_BYTES);
                    const int krow = bh * P.Nk + kt * KT;
                    tma_load_2d(sK + s * K_BYTES, &tmap_k, &bar.k_full[s], 0, krow);
                    tma_load_2d(sK + s * K_BYTES + K_BYTES / 2, &tmap_k, &bar.k_full[s], 64, krow);
                }
            }
        }
    } else if (warp == 4) {
        // ======================================================================= MMA issuer
        uint32_t tcount = 0, kc = 0;
        const uint32_t idesc = umma_idesc_bf16(KT, QG, 0, 0);
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, tcount++) {
            const uint32_t qb = tcount & 1;
            mbar_wait(&bar.q_full[qb], (tcount >> 1) & 1);
            for (int kt = 0; kt < P.n_kt; kt++, kc++) {
                const uint32_t s = kc % KSTAGES, ab = kc & 1;
                mbar_wait(&bar.k_full[s], (kc / KSTAGES) & 1);
                mbar_wait(&bar.acc_empty[ab], ((kc >> 1) & 1) ^ 1);
                tc_fence_after_sync();
                if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < D / 16; k16++) {
                        const uint64_t ad = umma_smem_desc(sK + s * K_BYTES + (k16 >> 2) * (K_BYTES / 2) + (k16 & 3) * 32, 16, 1024);
                        const uint64_t bd = umma_smem_desc(sQ + qb * Q_BYTES + (k16 >> 2) * (Q_BYTES / 2) + (k16 & 3) * 32, 16, 1024);
                        umma_ss(tm + ab * 256, ad, bd, idesc, k16 > 0);
                    }
                    umma_commit(&bar.k_empty[s]);
                    umma_commit(&bar.acc_full[ab]);
                    if (kt == P.n_kt - 1) umma_commit(&bar.q_empty[qb]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 4 || warp >= 8) {
        // ======================================================================= exp + column sum
        const int grp = warp >> 3;                        // 0: query columns 0-95, 1: columns 96-191
        const int wq = warp & 3;                          // TMEM lane quadrant = SM sub-partition
        const int et = grp * 128 + wq * 32 + lane;        // 0..255 over the two exp groups
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        const int key_in_tile = wq * 32 + lane;
        const int cbeg = grp * (QG / 2);
        uint32_t tcount = 0, kc = 0;
        const uint64_t c2 = pack_f32x2(SCALE_LOG2, SCALE_LOG2);
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, tcount++) {
            const int g = tile % P.G, bh = tile / P.G;
            const uint32_t lb = tcount & 1;
            if (et < QG) {
                const int row = g * QG + et;
                const float pv = row < P.Nq ? __ldg(P.p + (int64_t)bh * P.Nq + row) : 0.f;
                s_lp[lb][et] = pv > 0.f ? __log2f(pv) : -INFINITY;
            }
            named_bar_sync(1, 256);
            __nv_bfloat16* crow = P.cs + (int64_t)tile * P.cs_stride;
            for (int kt = 0; kt < P.n_kt; kt++, kc++) {
                const uint32_t ab = kc & 1;
                mbar_wait(&bar.acc_full[ab], (kc >> 1) & 1);
                tc_fence_after_sync();
                const uint32_t tacc = tm + ab * 256 + lane_off + cbeg;
                uint64_t acc[2] = {0ull, 0ull};
#pragma unroll 1
                for (int c0 = 0; c0 < QG / 2; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tacc + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 lp = *reinterpret_cast<const float4*>(&s_lp[lb][cbeg + c0 + j]);
                        const uint64_t xa = ffma2(pack_f32x2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), c2, pack_f32x2(lp.x, lp.y));
                        const uint64_t xb = ffma2(pack_f32x2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), c2, pack_f32x2(lp.z, lp.w));
                        float a0, a1, b0, b1;
                        unpack_f32x2(xa, a0, a1);
                        unpack_f32x2(xb, b0, b1);
                        acc[0] = fadd2(acc[0], pack_f32x2(fast_exp2(a0), fast_exp2(a1)));
                        if (CM_COLSUM_POLY > 0 && ((j >> 2) % (CM_COLSUM_POLY > 0 ? CM_COLSUM_POLY : 1)) == 0) acc[1] = fadd2(acc[1], exp2_poly2(xb));
                        else acc[1] = fadd2(acc[1], pack_f32x2(fast_exp2(b0), fast_exp2(b1)));
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&bar.acc_empty[ab]);
                float s0, s1, s2, s3;
                unpack_f32x2(acc[0], s0, s1);
                unpack_f32x2(acc[1], s2, s3);
                const float part = (s0 + s1) + (s2 + s3);
                // the two threads of a key meet on a 64-thread named barrier per lane quadrant
                if (grp == 1) {
                    s_part[ab][key_in_tile] = part;
                    named_bar_sync(2 + wq, 64);
                } else {
                    named_bar_sync(2 + wq, 64);
                    const int key = kt * KT + key_in_tile;
                    if (key < P.Nk) crow[key] = __float2bfloat16(part + s_part[ab][key_in_tile]);
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tm, 512);
}
}  // namespace cs

int launch_colsum(const __nv_bfloat16* q, const __nv_bfloat16* k, const float* p, __nv_bfloat16* cs_out, int B, int H,
                  int Nq, int Nk, int64_t cs_row_stride, cudaStream_t stream) {
    using namespace cs;
    CUtensorMap tq, tk;
    int rc = encode_tmap_2d_bf16_sw128(&tq, q, (uint64_t)B * H * Nq, D, D * 2, QG);
    if (rc) return rc;
    rc = encode_tmap_2d_bf16_sw128(&tk, k, (uint64_t)B * H * Nk, D, D * 2, KT);
    if (rc) return rc;
    static unsigned long long configured = 0;
    rc = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(colsum_kernel), SMEM_BYTES);
    if (rc) return rc;
    Params P{};
    P.p = p; P.cs = cs_out; P.BH = B * H; P.Nq = Nq; P.Nk = Nk; P.G = (Nq + QG - 1) / QG;
    P.cs_stride = cs_row_stride;
    const int64_t tiles = (int64_t)B * H * P.G;
    if (tiles > 2147483647ll) return CM_EINVAL;
    P.num_tiles = (int)tiles;
    P.n_kt = (Nk + KT - 1) / KT;
    const int grid = P.num_tiles < sm_count() ? P.num_tiles : sm_count();
    colsum_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, P);
    return (int)cudaGetLastError();
}

}  // namespace attn
}  // namespace cm
