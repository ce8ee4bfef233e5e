// Column-sparse MLP GEMMs for sm_100a.
//
// Replaces csrc/mlp/csp_mlp_mm1.cu (Hopper wgmma), the Triton mm2 kernel and the CUDA-graph
// launcher csrc/mlp/csp_mlp_mm2_and_scatter_add.cu, and csrc/indexed_io/scatter_add.cu.
//
// Both GEMMs share one persistent, warp-specialised skeleton (one CTA per SM, 512 threads):
//     warps 0-7   epilogue (2 warpgroups x 128 columns): TMEM accumulator -> registers -> fused elementwise -> global,
//                 with the cache / output tile prefetched one 32-column chunk ahead
//     warp  12    MMA issuer (one thread): tcgen05.mma M=128, N<=256, fp32 accumulators in TMEM,
//                 two 256-column accumulator buffers so tile i+1's main loop overlaps tile i's epilogue
//     warp  13    TMA: the dense operand tile (tokens x 64 of K) per stage
//     warps 8-11  gather producers: the ACTIVE weight rows of this 128-token block, copied by index
//                 with 16-byte cp.async into 128B-swizzled shared memory, i.e. the column-sparse
//                 operand becomes a dense tensor-core tile
// mm1:  C[m, j]   = gelu(A[m,:] . W1[idx[j],:] + b[idx[j]]) - PA[idx[j], m]      gathered rows = N side (K-major B)
// mm2:  O[m, :]  += P[m, 0:cnt] @ W2T[idx[0:cnt], :]                              gathered rows = K side (MN-major B)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include "tma.cuh"

namespace cm {
namespace mlp {

constexpr int BM = 128;            // token rows per tile (= MLP index group)
constexpr int BN = 256;            // output columns per tile
constexpr int BK = 64;             // K elements per stage (128 bytes)
constexpr int MAX_STAGES = 4;
constexpr int A_BYTES = BM * 128;  // 16 KB
constexpr int B_BYTES = BN * 128;  // 32 KB (mm1: 256 rows x 128 B; mm2: 4 n-chunks x 64 rows x 128 B)
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OUT_STAGE_BYTES = 32 * 128;            // mm2 epilogue: one 32-row x 64-column bf16 box per epilogue warp
constexpr int EPI1_BYTES = 2048;                     // mm1 epilogue, per warp: one 32x32 bf16 box of C, one 32-neuron x 32-token slice of the cache
constexpr int SMEM_MM1 = 4 * STAGE_BYTES + 16 * EPI1_BYTES + 1024;
constexpr int SMEM_MM2 = 4 * STAGE_BYTES + 8 * OUT_STAGE_BYTES + 1024;
constexpr int NUM_THREADS = 512;   // warps 0-7 epilogue (two warpgroups, 128 columns each) | 8-11 gather | 12 MMA | 13 TMA | 14-15 idle
constexpr int WARP_MMA = 12, WARP_TMA = 13, WARP_PROD0 = 8, NUM_EPI = 256;

struct Params {
    // mm1: a = tokens [M,K] (TMA), w = W1 [F,K], out = C [M,F] packed, bias [F], pa_T [F,M]
    // mm2: a = packed [M,F] (TMA), w = W2T [F,N], out = O [M,N]
    const __nv_bfloat16* w;
    __nv_bfloat16* out;
    const __nv_bfloat16* bias;
    __nv_bfloat16* pa_T;
    const int32_t* indices;
    const int32_t* counts;
    int M, K, F, N;            // mm1: K = reduction, F = #neurons;  mm2: F = #neurons (gathered K), N = out cols
    int64_t idx_stride;
    int update_pa;
    int n_mb, n_nb;            // tile grid
    int dbg;                   // stage-isolation timing switches: only read in -DCM_DEBUG_STAGES builds (CM_DBG)
};

struct __align__(8) Barriers {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    uint64_t acc_full[2], acc_empty[2];
};

__device__ __forceinline__ float gelu_tanh(float x) {
    // reference csrc/common/elementwise/gelu.cuh:29-31
    return x * 0.5f * (1.0f + fast_tanh(x * 0.79788456f * (1.0f + x * x * 0.044715f)));
}

// ------------------------------------------------------------------------------------------
// IS_MM2 = false: mm1, IS_MM2 = true: mm2
// ------------------------------------------------------------------------------------------
template <bool IS_MM2>
__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                             const __grid_constant__ CUtensorMap tmap_out, const Params P) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Barriers bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ __nv_bfloat16 s_bias[IS_MM2 ? 1 : 2][IS_MM2 ? 1 : BN];       // mm1 epilogue: bias of each packed column of the tile

    constexpr int STAGES = 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;

    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(&bar.full[i], 128 + 1); mbar_init(&bar.empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&bar.acc_full[i], 1); mbar_init(&bar.acc_empty[i], NUM_EPI); }
        fence_mbar_init();
    }
    if (warp == WARP_MMA) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    if (warp == WARP_TMA && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_out); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const int total_tiles = P.n_mb * P.n_nb;

    // work size of a tile along its gathered dimension
    //   mm1: ncols = packed columns of this tile (<= 256), ksteps = K / 64
    //   mm2: ncols = 256, ksteps = ceil(cnt / 64)
    auto tile_shape = [&](int tile, int& mb, int& nb, int& ncols, int& ksteps, int& klast) {
        mb = tile / P.n_nb;
        nb = tile % P.n_nb;
        int cnt = __ldg(P.counts + mb);
        cnt = cnt < 0 ? 0 : (cnt > P.F ? P.F : cnt);
        cnt &= ~15;
        if (!IS_MM2) {
            ncols = min(BN, cnt - nb * BN);
            ksteps = P.K / BK;
            klast = BK;
        } else {
            ncols = cnt > 0 ? BN : 0;
            ksteps = (cnt + BK - 1) / BK;
            klast = cnt - (ksteps - 1) * BK;
        }
    };

    // =========================================================================== gather producers
    if (warp >= WARP_PROD0 && warp < WARP_PROD0 + 4) {
        setmaxnreg_dec<88>();
        const int pt = tid - WARP_PROD0 * 32;
        uint32_t it = 0;       // global stage counter
        // the weight matrix (75 MB at FLUX sizes) is the only operand with reuse across token blocks: keep it in
        // L2 against the streaming token / cache / output traffic
        const uint64_t pol_w = l2_policy_evict_last();
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int mb, nb, ncols, ksteps, klast;
            tile_shape(tile, mb, nb, ncols, ksteps, klast);
            if (ncols <= 0) continue;
            const int32_t* ip = P.indices + (int64_t)mb * P.idx_stride;
            if (!IS_MM2) {
                // rows (pt>>3) + 16 i of the B tile = packed columns nb*256 + row; fixed for the whole K loop
                const int chunk = pt & 7, r0 = pt >> 3;
                const __nv_bfloat16* src[16];
                uint32_t okm = 0;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int r = r0 + 16 * i;
                    const bool ok = r < ncols;
                    int f = ok ? __ldg(ip + nb * BN + r) : 0;
                    f = f < 0 ? 0 : (f >= P.F ? P.F - 1 : f);
                    src[i] = P.w + (int64_t)f * P.K + chunk * 8;
                    okm |= (ok ? 1u : 0u) << i;
                }
                for (int ks = 0; ks < ksteps; ks++, it++) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&bar.empty[s], ((it / STAGES) & 1) ^ 1);
                    const uint32_t dst = sbase + s * STAGE_BYTES + A_BYTES;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int r = r0 + 16 * i;
                        if (((okm >> i) & 1u) && !CM_DBG(P, 1))
                            cp_async_16_hint(dst + r * 128 + ((chunk ^ (r & 7)) << 4), src[i] + ks * BK, pol_w);
                    }
                    cp_async_mbar_arrive_noinc(&bar.full[s]);
                }
            } else {
                // stage ks gathers W2T rows idx[ks*64 + r], r < 64, columns nb*256 .. +256 (512 B each):
                // warp w copies rows w + 4 i, lane = 16-byte chunk of the row
                const int chunk = pt & 31, w = pt >> 5;
                const __nv_bfloat16* wb = P.w + (int64_t)nb * BN + chunk * 8;
                const uint32_t dcol = (uint32_t)(chunk >> 3) * (BK * 128) ;
                auto fetch = [&](int ks) -> int {
                    const int pos = ks * BK + w + 4 * (lane & 15);
                    const int cnt_all = (ksteps - 1) * BK + klast;
                    // the raw value is returned untouched: any arithmetic on it here would make the warp wait
                    // for the load right away and defeat the prefetch (the clamp happens at the use site)
                    return pos < cnt_all ? __ldg(ip + pos) : -1;
                };
                // index loads run IDX_AHEAD stages ahead of their use (statically rotated registers): a stage
                // (~0.5 us) is shorter than a global-load round trip
                constexpr int IDX_AHEAD = 4;
                int fq[IDX_AHEAD];
#pragma unroll
                for (int j = 0; j < IDX_AHEAD; j++) fq[j] = j < ksteps ? fetch(j) : -1;
                for (int ks0 = 0; ks0 < ksteps; ks0 += IDX_AHEAD) {
#pragma unroll
                    for (int j = 0; j < IDX_AHEAD; j++) {
                        const int ks = ks0 + j;
                        if (ks >= ksteps) break;
                        const int f_cur = fq[j];
                        if (ks + IDX_AHEAD < ksteps) fq[j] = fetch(ks + IDX_AHEAD);
                        const uint32_t s = it % STAGES;
                        mbar_wait(&bar.empty[s], ((it / STAGES) & 1) ^ 1);
                        const uint32_t dst = sbase + s * STAGE_BYTES + A_BYTES + dcol;
#pragma unroll
                        for (int i = 0; i < 16; i++) {
                            const int r = w + 4 * i;
                            int f = __shfl_sync(0xffffffffu, f_cur, i);
                            f = f >= P.F ? P.F - 1 : f;
                            if (f >= 0 && !CM_DBG(P, 1)) cp_async_16_hint(dst + r * 128 + (((chunk & 7) ^ (r & 7)) << 4), wb + (int64_t)f * P.N, pol_w);
                        }
                        cp_async_mbar_arrive_noinc(&bar.full[s]);
                        it++;
                    }
                }
            }
        }
        cp_async_wait_all();
    }
    // =========================================================================== TMA (dense operand)
    else if (warp == WARP_TMA) {
        setmaxnreg_dec<88>();
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int mb, nb, ncols, ksteps, klast;
                tile_shape(tile, mb, nb, ncols, ksteps, klast);
                if (ncols <= 0) continue;
                for (int ks = 0; ks < ksteps; ks++, it++) {
                    const uint32_t s = it % STAGES;
                    mbar_wait(&bar.empty[s], ((it / STAGES) & 1) ^ 1);
                    if (CM_DBG(P, 2)) { mbar_arrive(&bar.full[s]); continue; }
                    mbar_arrive_expect_tx(&bar.full[s], A_BYTES);
                    tma_load_2d(sbase + s * STAGE_BYTES, &tmap_a, &bar.full[s], ks * BK, mb * BM);
                }
            }
        }
    }
    // =========================================================================== MMA issuer
    else if (warp == WARP_MMA) {
        setmaxnreg_dec<88>();
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int mb, nb, ncols, ksteps, klast;
            tile_shape(tile, mb, nb, ncols, ksteps, klast);
            if (ncols <= 0) continue;
            const uint32_t buf = tcount & 1;
            mbar_wait(&bar.acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d = tm + buf * BN;
            const uint32_t idesc = umma_idesc_bf16(BM, ncols, 0, IS_MM2 ? 1 : 0);
            for (int ks = 0; ks < ksteps; ks++, it++) {
                const uint32_t s = it % STAGES;
                mbar_wait(&bar.full[s], (it / STAGES) & 1);
                tc_fence_after_sync();
                if (elect_one()) {      // elect.sync: ptxas emits the UTCHMMAs back to back (a lane test costs a per-instruction ELECT loop)
                    const uint32_t sa = sbase + s * STAGE_BYTES, sb = sa + A_BYTES;
                    const int k16s = (ks == ksteps - 1 ? klast : BK) / 16;
                    for (int k16 = 0; k16 < (CM_DBG(P, 4) ? 0 : k16s); k16++) {
                        const uint64_t ad = umma_smem_desc(sa + k16 * 32, 16, 1024);
                        const uint64_t bd = IS_MM2 ? umma_smem_desc(sb + k16 * 2048, BK * 128, 1024)
                                                   : umma_smem_desc(sb + k16 * 32, 16, 1024);
                        umma_ss(d, ad, bd, idesc, (ks | k16) != 0);
                    }
                    umma_commit(&bar.empty[s]);
                    if (ks == ksteps - 1) umma_commit(&bar.acc_full[buf]);
                }
                __syncwarp();
            }
            tcount++;
        }
    }
    // =========================================================================== epilogue
    else if (warp < 8) {
        setmaxnreg_inc<152>();
        const int wq = warp & 3, half = warp >> 2;           // TMEM lane quadrant, column half of the tile
        const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int mb, nb, ncols, ksteps, klast;
            tile_shape(tile, mb, nb, ncols, ksteps, klast);
            if (ncols <= 0) continue;
            const uint32_t buf = tcount & 1;
            const int m = mb * BM + wq * 32 + lane;
            const int cbeg = half * (BN / 2);
            const int cend = min(ncols, cbeg + BN / 2);
            if constexpr (!IS_MM2) {
                // stage this tile's biases (256 epilogue threads, one packed column each)
                const int32_t* ipt = P.indices + (int64_t)mb * P.idx_stride + nb * BN;
                int f = tid < ncols ? __ldg(ipt + tid) : 0;
                f = f < 0 ? 0 : (f >= P.F ? P.F - 1 : f);
                s_bias[buf][tid] = P.bias[f];
                named_bar_sync(1, NUM_EPI);
            }
            const uint32_t tacc = tm + buf * BN + lane_off;
            if (CM_DBG(P, 8)) {
                mbar_wait(&bar.acc_full[buf], (tcount >> 1) & 1);
            } else if constexpr (!IS_MM2) {
                // The weight gather saturates the LSU / L1 path, so the epilogue keeps off it: per 32-column chunk a
                // warp (32 tokens) pulls its 32-neuron x 32-token slice of the cache into shared memory with 16-byte
                // cp.async (one chunk ahead), and hands its 32 x 32 block of C to the TMA as one box store.
                const int32_t* ip = P.indices + (int64_t)mb * P.idx_stride + nb * BN;
                const uint32_t sY = sbase + STAGES * STAGE_BYTES + warp * EPI1_BYTES;      // C box, SWIZZLE_64B
                const uint32_t sX = sY + 8 * EPI1_BYTES;                                   // cache slice [32 neurons][32 tokens]
                const int m0 = mb * BM + wq * 32;
                __nv_bfloat16* crow = P.out + (int64_t)m * P.F + nb * BN;
                // lane j holds the neuron id of packed column c0 + j
                auto load_f = [&](int c0) -> int {
                    int f = (c0 + lane < cend) ? __ldg(ip + c0 + lane) : 0;
                    return f < 0 ? 0 : (f >= P.F ? P.F - 1 : f);
                };
                auto fetch_pa = [&](int c0, int f_lane) {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int row = 8 * i + (lane >> 2), piece = lane & 3;
                        const int f = __shfl_sync(0xffffffffu, f_lane, row);
                        if (c0 + row < cend) cp_async_16(sX + row * 64 + piece * 16, P.pa_T + (int64_t)f * P.M + m0 + piece * 8);
                    }
                    cp_async_commit();
                };
                int f_nxt = 0;
                if (cbeg < cend) { f_nxt = load_f(cbeg); fetch_pa(cbeg, f_nxt); }
                mbar_wait(&bar.acc_full[buf], (tcount >> 1) & 1);
                tc_fence_after_sync();
                for (int c0 = cbeg; c0 < cend; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tacc + c0, r);
                    const int f_cur = f_nxt;
                    cp_async_wait_all();
                    __syncwarp();
                    uint32_t pa[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) pa[j] = ld_shared_u16(sX + j * 64 + lane * 2) << 16;
                    __syncwarp();
                    if (c0 + 32 < cend) { f_nxt = load_f(c0 + 32); fetch_pa(c0 + 32, f_nxt); }
                    tmem_ld_wait();
                    const int nvalid = min(32, cend - c0);     // multiple of 16
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float2 bb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&s_bias[buf][c0 + j]));
                        const float g0 = gelu_tanh(__uint_as_float(r[j]) + bb.x) - __uint_as_float(pa[j]);
                        const float g1 = gelu_tanh(__uint_as_float(r[j + 1]) + bb.y) - __uint_as_float(pa[j + 1]);
                        pk[j >> 1] = pack_bf16x2(g0, g1);
                    }
                    if (P.update_pa) {
                        unsigned short* pa_out = reinterpret_cast<unsigned short*>(P.pa_T) + m;
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const int fj = __shfl_sync(0xffffffffu, f_cur, j);
                            if (j < nvalid) {
                                const float d = (j & 1) ? bf16_hi(pk[j >> 1]) : bf16_lo(pk[j >> 1]);
                                pa_out[(int64_t)fj * P.M] = __bfloat16_as_ushort(__float2bfloat16(__uint_as_float(pa[j]) + d));
                            }
                        }
                    }
                    if (CM_DBG(P, 16)) continue;
                    if (nvalid == 32) {
                        if (lane == 0) bulk_wait_read<0>();          // the previous box has left the staging buffer
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            st_shared_v4(sY + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) { tma_store_2d(&tmap_out, sY, nb * BN + c0, m0); bulk_commit(); }
                    } else {
                        // ragged tail of the block's column list (count % 32 == 16): columns past `count` stay untouched
                        *reinterpret_cast<uint4*>(crow + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4*>(crow + c0 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    }
                }
            } else {
                // O[m, tile columns] += bf16(acc): each warp packs its 32 rows x 64 columns into a 128B-swizzled
                // 4 KB box in shared memory and ONE thread hands it to the TMA as a bf16 reduce-add on `out`
                // (bf16(acc) first, then the bf16 add at the L2: triton/csp_mlp_mm2.py:100-101).  No global loads
                // or stores go through the LSU, which the weight gather saturates.
                const uint32_t stg = sbase + STAGES * STAGE_BYTES + warp * OUT_STAGE_BYTES;
                mbar_wait(&bar.acc_full[buf], (tcount >> 1) & 1);
                tc_fence_after_sync();
#pragma unroll 1
                for (int ch = 0; ch < 2; ch++) {
                    const int c0 = cbeg + ch * 64;
                    uint32_t r[64];
                    tmem_ld32(tacc + c0, r);
                    tmem_ld32(tacc + c0 + 32, r + 32);
                    tmem_ld_wait();
                    if (lane == 0) bulk_wait_read<0>();      // the previous box has been read out of the staging buffer
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        uint32_t w[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            w[j] = pack_bf16x2(__uint_as_float(r[c * 8 + 2 * j]), __uint_as_float(r[c * 8 + 2 * j + 1]));
                        st_shared_v4(stg + sw128_off(lane, c), w[0], w[1], w[2], w[3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && !CM_DBG(P, 16)) {
                        tma_reduce_add_2d(&tmap_out, stg, nb * BN + c0, mb * BM + wq * 32);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bar.acc_empty[buf]);
            tcount++;
        }
    } else {
        setmaxnreg_dec<88>();     // warps 14-15: idle, setmaxnreg is warpgroup-wide
    }

    if (warp < 8 && lane == 0) bulk_wait<0>();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------
// scatter-add: pa_T[idx[mb,c], mb*128 + r] += packed[mb*128 + r, c]   (bf16 + bf16 -> bf16)
// One CTA per (token block, 64 packed columns): the 128 x 64 packed tile is transposed through
// shared memory so that both the read of `packed` and the read-modify-write of pa_T rows are
// coalesced.  Every (neuron, token) pair receives exactly one add, so no atomics are needed.
// ------------------------------------------------------------------------------------------
constexpr int SC_COLS = 64;
__global__ void __launch_bounds__(256) scatter_add_kernel(const __nv_bfloat16* __restrict__ packed, __nv_bfloat16* __restrict__ pa_T,
                                                          const int32_t* __restrict__ indices, const int32_t* __restrict__ counts,
                                                          int M, int F, int64_t idx_stride) {
    __shared__ __nv_bfloat16 tile[BM][SC_COLS + 2];
    const int mb = blockIdx.y, c0 = blockIdx.x * SC_COLS;
    int cnt = counts[mb];
    cnt = (cnt < 0 ? 0 : (cnt > F ? F : cnt)) & ~15;      // the same clamp + 16-column granularity as mlp_kernel::tile_shape
    if (c0 >= cnt) return;
    const int ncol = min(SC_COLS, cnt - c0);
    const int tid = threadIdx.x;
    // load: 32 lanes cover one row's 64 columns (2 per lane)
    for (int r = tid >> 5; r < BM; r += 8) {
        const int c = (tid & 31) * 2;
        const __nv_bfloat16* src = packed + (int64_t)(mb * BM + r) * F + c0 + c;
        __nv_bfloat162 v = c < ncol ? *reinterpret_cast<const __nv_bfloat162*>(src) : __nv_bfloat162();
        tile[r][c] = v.x;
        tile[r][c + 1] = v.y;
    }
    __syncthreads();
    // each warp owns columns warp, warp+8, ...; lane handles 4 consecutive token rows (8 bytes)
    for (int c = tid >> 5; c < ncol; c += 8) {
        int f = indices[(int64_t)mb * idx_stride + c0 + c];
        if ((unsigned)f >= (unsigned)F) continue;
        const int r = (tid & 31) * 4;
        __nv_bfloat16* dst = pa_T + (int64_t)f * M + mb * BM + r;
        uint2 old = *reinterpret_cast<uint2*>(dst);
        const uint32_t ov[2] = {old.x, old.y};
        uint32_t w[2];
#pragma unroll
        for (int j = 0; j < 2; j++)
            w[j] = pack_bf16x2(bf16_lo(ov[j]) + __bfloat162float(tile[r + 2 * j][c]),
                               bf16_hi(ov[j]) + __bfloat162float(tile[r + 2 * j + 1][c]));
        *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
    }
}

}  // namespace mlp
}  // namespace cm

// ------------------------------------------------------------------------------------------
using namespace cm;
using namespace cm::mlp;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool IS_MM2>
static int launch_mlp(const CUtensorMap& tmap, const CUtensorMap& tmap_out, Params& P, cudaStream_t stream) {
    static unsigned long long configured = 0;          // one per template instance
    auto kern = mlp_kernel<IS_MM2>;
    const int rc = opt_in_dynamic_smem(configured, reinterpret_cast<const void*>(kern), IS_MM2 ? SMEM_MM2 : SMEM_MM1);
    if (rc) return rc;
    const int tiles = P.n_mb * P.n_nb;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, NUM_THREADS, IS_MM2 ? SMEM_MM2 : SMEM_MM1, stream>>>(tmap, tmap_out, P);
    return (int)cudaGetLastError();
}

extern "C" int cm_csp_mlp_mm1(const void* a, const void* w1, void* c, const void* bias, void* pa_T,
                              const int32_t* indices, const int32_t* counts, int M, int K, int F,
                              int64_t idx_stride, int update_pa, void* stream) {
    if (M <= 0 || K <= 0 || F <= 0 || M % BM || K % BK || F % 16 || idx_stride < F) return CM_EINVAL;
    if (!a || !w1 || !c || !bias || !pa_T || !indices || !counts) return CM_EINVAL;
    if (!al16(a) || !al16(w1) || !al16(c) || !al16(pa_T)) return CM_EALIGN;
    if (!is_sm100()) return CM_EARCH;
    CUtensorMap tmap;
    int rc = encode_tmap_2d_bf16_sw128(&tmap, a, (uint64_t)M, (uint64_t)K, (uint64_t)K * 2, BM);
    if (rc) return rc;
    Params P{};
    P.w = (const __nv_bfloat16*)w1; P.out = (__nv_bfloat16*)c; P.bias = (const __nv_bfloat16*)bias;
    P.pa_T = (__nv_bfloat16*)pa_T; P.indices = indices; P.counts = counts;
    P.M = M; P.K = K; P.F = F; P.N = F; P.idx_stride = idx_stride; P.update_pa = update_pa ? 1 : 0;
    P.n_mb = M / BM; P.n_nb = (F + BN - 1) / BN;
    P.dbg = debug_flags();
    CUtensorMap tmap_c;        // C [M, F] in boxes of 32 rows x 32 columns (the epilogue's TMA stores)
    rc = encode_tmap_2d_bf16(&tmap_c, c, (uint64_t)M, (uint64_t)F, (uint64_t)F * 2, 32, 32, 64);
    if (rc) return rc;
    return launch_mlp<false>(tmap, tmap_c, P, (cudaStream_t)stream);
}

extern "C" int cm_csp_scatter_add(const void* packed, void* pa_T, const int32_t* indices, const int32_t* counts,
                                  int M, int F, int64_t idx_stride, void* stream) {
    if (M <= 0 || F <= 0 || M % BM || F % 2 || idx_stride < F) return CM_EINVAL;
    if (!packed || !pa_T || !indices || !counts) return CM_EINVAL;
    if (!al16(packed) || !al16(pa_T)) return CM_EALIGN;
    dim3 grid((F + SC_COLS - 1) / SC_COLS, M / BM);
    scatter_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)packed, (__nv_bfloat16*)pa_T, indices,
                                                               counts, M, F, idx_stride);
    return (int)cudaGetLastError();
}

extern "C" int cm_csp_mlp_mm2(const void* packed, const void* w2_T, void* out, void* pa_T, const int32_t* indices,
                              const int32_t* counts, int M, int F, int N, int64_t idx_stride, int do_scatter,
                              void* stream) {
    if (M <= 0 || F <= 0 || N <= 0 || M % BM || N % BN || F % BK || idx_stride < F) return CM_EINVAL;
    if (!packed || !w2_T || !out || !indices || !counts || (do_scatter && !pa_T)) return CM_EINVAL;
    if (!al16(packed) || !al16(w2_T) || !al16(out)) return CM_EALIGN;
    if (!is_sm100()) return CM_EARCH;
    if (do_scatter) {
        int rc = cm_csp_scatter_add(packed, pa_T, indices, counts, M, F, idx_stride, stream);
        if (rc) return rc;
    }
    CUtensorMap tmap;
    int rc = encode_tmap_2d_bf16_sw128(&tmap, packed, (uint64_t)M, (uint64_t)F, (uint64_t)F * 2, BM);
    if (rc) return rc;
    Params P{};
    P.w = (const __nv_bfloat16*)w2_T; P.out = (__nv_bfloat16*)out; P.bias = nullptr; P.pa_T = nullptr;
    P.indices = indices; P.counts = counts;
    P.M = M; P.K = F; P.F = F; P.N = N; P.idx_stride = idx_stride; P.update_pa = 0;
    P.n_mb = M / BM; P.n_nb = N / BN;
    P.dbg = debug_flags();
    CUtensorMap tmap_out;      // out [M, N] in boxes of 32 rows x 64 columns (the epilogue's TMA reduce-add)
    rc = encode_tmap_2d_bf16_sw128(&tmap_out, out, (uint64_t)M, (uint64_t)N, (uint64_t)N * 2, 32);
    if (rc) return rc;
    return launch_mlp<true>(tmap, tmap_out, P, (cudaStream_t)stream);
}
