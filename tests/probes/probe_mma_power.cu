// Hardware probe: how far does the power cap pull the clock down under a sustained tcgen05.mma stream, and does an
// M = 64 MMA (or an M = 128 MMA whose upper 64 A rows are zero -- block 1 of the 192-row attention tile) cost less
// power than a full M = 128 one?  Every SM runs the attention step's MMA mix (8 x S-type SS + 8 x P.V-type TS,
// M x 128 x 16 each) on pseudo-random bf16 operands for about a second per variant; reported: cycles per step (the
// pipe's own pace), wall time per step, and the effective SM clock = cycles / time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mma_power tests/probes/probe_mma_power.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "../../chipmunk_b200/csrc/ptx.cuh"
using namespace cm;

constexpr int TILE = 32768, NSLOT = 4;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// two bf16 in [-1, 1) from a hash
__device__ __forceinline__ uint32_t rnd_bf16x2(uint32_t i) {
    const uint32_t h = hash32(i);
    const float a = (float)(h & 0xffff) / 32768.f - 1.f, b = (float)(h >> 16) / 32768.f - 1.f;
    return pack_bf16x2(a, b);
}

template <int M, int ZERO_UPPER>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int steps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t done;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t sQ = sbase, sKV = sbase + TILE;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&done, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    uint32_t* sm = reinterpret_cast<uint32_t*>(smem + (sbase - smem_u32(smem)));
    for (int i = tid; i < (TILE * (NSLOT + 1)) / 4; i += 128) {
        uint32_t v = rnd_bf16x2(i * 2654435761u + blockIdx.x);
        // Q tile: two 64-column halves of [128 rows x 128 B]; row r of a half starts at r * 128 bytes
        if (ZERO_UPPER && i < TILE / 4 && ((i * 4) % (TILE / 2)) / 128 >= 64) v = 0;
        sm[i] = v;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    {   // P operands (bf16 pairs) in the first 64 columns of both S buffers; zero on lanes 64-127 for the half-empty block
        uint32_t r[32];
        for (int b = 0; b < 2; b++)
            for (int c0 = 0; c0 < 64; c0 += 32) {
                for (int j = 0; j < 32; j++) r[j] = (ZERO_UPPER && tid >= 64) ? 0u : rnd_bf16x2((tid * 64 + c0 + j) * 40503u + b);
                tmem_st_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + b * 128 + c0, r);
            }
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0) {
        const uint32_t idesc_s = umma_idesc_bf16(M, 128, 0, 0), idesc_pv = umma_idesc_bf16(M, 128, 0, 1);
        const uint64_t desc_q = umma_smem_desc(sQ, 16, 1024), desc_k = umma_smem_desc(sKV, 16, 1024);
        const uint64_t desc_v = umma_smem_desc(sKV, TILE / 2, 1024);
        long long t0 = clock64();
        for (int st = 0; st < steps; st++) {
            const uint32_t slot_k = (2 * st) % NSLOT, slot_v = (2 * st + 1) % NSLOT;
            if (elect_one()) {
                // S goes to columns [256, 384) so that the P operands in [0, 128) stay what they are
#pragma unroll
                for (int k16 = 0; k16 < 8; k16++) {
                    const uint64_t off = (uint64_t)((((k16 >> 2) * (TILE / 2)) + (k16 & 3) * 32) >> 4);
                    umma_ss(tm + 256, desc_q + off, desc_k + (uint64_t)(slot_k * (TILE >> 4)) + off, idesc_s, k16 > 0);
                }
#pragma unroll
                for (int j = 0; j < 8; j++)
                    umma_ts(tm + 384, tm + (st & 1) * 128 + j * 8, desc_v + (uint64_t)(slot_v * (TILE >> 4)) + (uint64_t)(j * (2048 >> 4)), idesc_pv, j > 0);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(&done);
        __syncwarp();
        mbar_wait(&done, 0);
        if (lane_id() == 0) out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

template <int M, int Z>
static void run(const char* name, long long* d, int steps) {
    const int smem = TILE * (NSLOT + 1) + 1024;
    cudaFuncSetAttribute(probe<M, Z>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<M, Z><<<148, 128, smem>>>(d, steps / 20);          // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<M, Z><<<148, 128, smem>>>(d, steps);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("fail %s\n", cudaGetErrorString(e)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 148; i++) cyc += (double)h[i] / 148;
    const double flops = 148.0 * steps * 16 * 2.0 * 128 * 128 * 16;      // counted at M = 128 for every variant
    printf("%-44s %7.1f cycles/step  %7.3f us/step  clock %.3f GHz  (%6.1f TFLOP/s at M=128-equivalent issue)\n", name, cyc / steps,
           ms * 1e3 / steps, cyc / (ms * 1e6), flops / (ms * 1e-3) / 1e12);
}

int main() {
    long long* d; cudaMalloc(&d, 256 * 8);
    const int steps = 1200000;           // ~0.7 s per variant at 1024 cycles per step
    for (int rep = 0; rep < 2; rep++) {
        run<128, 0>("M=128, all rows random", d, steps);
        run<128, 1>("M=128, upper 64 rows of A zero", d, steps);
        run<64, 0>("M=64", d, steps);
    }
    return 0;
}
