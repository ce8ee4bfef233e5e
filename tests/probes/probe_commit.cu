// Hardware probe: latency of tcgen05.commit -> mbarrier completion, alone and after MMAs of known length,
// and of a plain mbarrier arrive -> waiter wake-up between two warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_commit tests/probes/probe_commit.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "../../chipmunk_b200/csrc/ptx.cuh"
using namespace cm;

__global__ void __launch_bounds__(64, 1) probe(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    for (int i = tid; i < 65536 / 4; i += 64) reinterpret_cast<uint32_t*>(smem + (sbase - smem_u32(smem)))[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    uint32_t ph = 0;
    if (tid == 0) {
        // (a) commit with nothing in flight
        for (int rep = 0; rep < 4; rep++) {
            long long t0 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1;
            out[rep] = clock64() - t0;
        }
        // (b) n MMAs (M128 N128 K16, 64 cycles each nominal) then commit
        const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
        const uint64_t ad = umma_smem_desc(sbase, 16, 1024), bd = umma_smem_desc(sbase + 32768, 16, 1024);
        int ns[6] = {1, 8, 16, 32, 64, 128};
        for (int c = 0; c < 6; c++) {
            long long t0 = clock64();
            for (int i = 0; i < ns[c]; i++) umma_ss(tm, ad, bd, idesc, i > 0);
            long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1;
            out[4 + 2 * c] = t1 - t0;              // issue time
            out[5 + 2 * c] = clock64() - t0;       // issue + execute + commit + wake
        }
    }
    __syncthreads();
    // (c) ping-pong between warp 0 lane 0 and warp 1 lane 0 with plain arrives: round trip
    if (tid == 0) {
        long long t0 = clock64();
        uint32_t p = 0;
        for (int i = 0; i < 100; i++) { mbar_arrive(&bar2); mbar_wait(&bar3, p); p ^= 1; }
        out[20] = (clock64() - t0) / 100;
    } else if (tid == 32) {
        uint32_t p = 0;
        for (int i = 0; i < 100; i++) { mbar_wait(&bar2, p); p ^= 1; mbar_arrive(&bar3); }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 32 * 8); cudaMemset(d, 0, 32 * 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
    probe<<<1, 64, 70 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("fail %s\n", cudaGetErrorString(e)); return 1; }
    long long h[32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("commit with empty pipe -> waiter wake: %lld %lld %lld %lld cycles\n", h[0], h[1], h[2], h[3]);
    int ns[6] = {1, 8, 16, 32, 64, 128};
    for (int c = 0; c < 6; c++)
        printf("%3d x MMA(M128 N128 K16): issue %lld cycles, issue+exec+commit+wake %lld cycles (nominal exec %d)\n", ns[c], h[4 + 2 * c], h[5 + 2 * c], ns[c] * 64);
    printf("mbarrier arrive/wait ping-pong between two warps: %lld cycles per round trip\n", h[20]);
    return 0;
}
