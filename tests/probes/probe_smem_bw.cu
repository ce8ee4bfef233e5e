// Hardware probe: does a tcgen05.mma whose operands both come from shared memory (SS) keep its nominal rate when the
// operand addresses walk real tiles (not one cached 4 KB block) and while bulk copies stream new tiles into shared memory?
// One "step" = what an attention step issues: 8 x (M128 N128 K16) S-type MMAs (A = Q tile, B = K slot) + 8 x P.V-type
// MMAs (A from TMEM, B = V slot, MN-major).  Variants:
//   ss      S-type MMAs read A and B from shared memory          (the kernels as they are)
//   ts      S-type MMAs read A from TMEM                         (Q kept in TMEM)
//   +copy   a second thread streams 64 KB per step into the ring with cp.async.bulk (the TMA write path)
// Prints cycles per step on one SM, and with all 148 SMs running the same loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_smem_bw tests/probes/probe_smem_bw.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "../../chipmunk_b200/csrc/ptx.cuh"
using namespace cm;

constexpr int TILE = 32768, NSLOT = 4, STEPS = 256;

template <int TS, int NS, int NPV>
__global__ void __launch_bounds__(128, 1) probe(long long* out, const uint8_t* src, int with_copy) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t done, landed;
    __shared__ volatile int issued;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t sQ = sbase, sKV = sbase + TILE;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&done, 1); mbar_init(&landed, 1); issued = -1; fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    for (int i = tid; i < (TILE * (NSLOT + 1)) / 4; i += 128) reinterpret_cast<uint32_t*>(smem + (sbase - smem_u32(smem)))[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    if (warp == 0) {
        // every MMA batch is issued under elect.sync (back-to-back UTCHMMAs), fully unrolled with constant offsets
        const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0), idesc_pv = umma_idesc_bf16(128, 128, 0, 1);
        const uint64_t desc_q = umma_smem_desc(sQ, 16, 1024), desc_k = umma_smem_desc(sKV, 16, 1024);
        const uint64_t desc_v = umma_smem_desc(sKV, TILE / 2, 1024);
        long long t0 = clock64();
        for (int st = 0; st < STEPS; st++) {
            const uint32_t slot_k = (2 * st) % NSLOT, slot_v = (2 * st + 1) % NSLOT;
            if (elect_one()) {
#pragma unroll
                for (int k16 = 0; k16 < NS; k16++) {
                    const uint64_t off = (uint64_t)(((((k16 & 7) >> 2) * (TILE / 2)) + (k16 & 3) * 32) >> 4);
                    const uint64_t bd = desc_k + (uint64_t)(slot_k * (TILE >> 4)) + off;
                    if (TS) umma_ts(tm + (st & 1) * 128, tm + 384 + (k16 & 7) * 8, bd, idesc_s, k16 > 0);
                    else umma_ss(tm + (st & 1) * 128, desc_q + off, bd, idesc_s, k16 > 0);
                }
#pragma unroll
                for (int j = 0; j < NPV; j++)
                    umma_ts(tm + 256, tm + (st & 1) * 128 + (j & 7) * 8, desc_v + (uint64_t)(slot_v * (TILE >> 4)) + (uint64_t)((j & 7) * (2048 >> 4)), idesc_pv, 1);
                if (with_copy) issued = st;
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&done);
        __syncwarp();
        mbar_wait(&done, 0);
        if (lane_id() == 0) out[blockIdx.x] = (clock64() - t0) / STEPS;
    } else if (tid == 32 && with_copy) {
        // 64 KB per step into the ring, in 16 KB bulk copies from an L2-resident source
        uint32_t ph = 0;
        for (int st = 0; st < STEPS; st++) {
            while (issued < st) { }
            mbar_arrive_expect_tx(&landed, 2 * TILE);
            for (int c = 0; c < 4; c++)
                bulk_g2s(sKV + ((2 * st + 2) % NSLOT) * TILE + c * 16384, src + ((st * 4 + c) % 64) * 16384 + (size_t)blockIdx.x * 0, 16384, &landed);
            mbar_wait(&landed, ph); ph ^= 1;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 256 * 8); cudaMemset(d, 0, 256 * 8);
    uint8_t* src; cudaMalloc(&src, 64 * 16384); cudaMemset(src, 0, 64 * 16384);
    const int smem = TILE * (NSLOT + 1) + 1024;
    struct V { const char* name; void (*fn)(long long*, const uint8_t*, int); int copy; };
    V v[] = {
        {"S(ss) + PV(ts)            ", probe<0, 8, 8>, 0}, {"S(ss) + PV(ts) + copy     ", probe<0, 8, 8>, 1},
        {"S(ts) + PV(ts)            ", probe<1, 8, 8>, 0}, {"S(ts) + PV(ts) + copy     ", probe<1, 8, 8>, 1},
        {"S(ss) x 16                ", probe<0, 16, 0>, 0}, {"S(ss) x 16 + copy         ", probe<0, 16, 0>, 1},
        {"S(ts) x 16                ", probe<1, 16, 0>, 0}, {"S(ts) x 16 + copy         ", probe<1, 16, 0>, 1},
        {"PV(ts) x 16               ", probe<0, 0, 16>, 0}, {"PV(ts) x 16 + copy        ", probe<0, 0, 16>, 1},
    };
    for (auto& c : v) cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int grid : {1, 148}) {
        for (auto& c : v) {
            c.fn<<<grid, 128, smem>>>(d, src, c.copy);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("fail %s\n", cudaGetErrorString(e)); return 1; }
            long long h[256]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0, mn = 1 << 30;
            for (int i = 0; i < grid; i++) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
            printf("grid %3d  %s %5lld .. %5lld cycles per step (nominal 1024)\n", grid, c.name, mn, mx);
        }
    }
    return 0;
}
