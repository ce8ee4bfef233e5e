// Hardware probe: MUFU.EX2 issue cost per warp instruction with 32 vs 16 active lanes, and with one vs two
// warps on the same SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_mufu ...
#include <cstdio>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__global__ void k(float* out, long long* cyc, int active_lanes, int nwarps_active) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a[8];
    for (int i = 0; i < 8; i++) a[i] = -0.001f * (lane + i);
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps_active && lane < active_lanes) {
        for (int it = 0; it < 256; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = ex2(a[i]) - 1.0f;     // 8 independent chains
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[threadIdx.x] = s;
    if (lane == 0) cyc[warp] = t1 - t0;
}
int main() {
    float* o; long long* c; cudaMalloc(&o, 1024 * 4); cudaMalloc(&c, 32 * 8);
    // warps w and w+4 share an SM sub-partition
    struct { int threads, lanes, nw; const char* name; } cfg[] = {
        {128, 32, 4, "4 warps (1 per SMSP), 32 lanes"}, {128, 16, 4, "4 warps (1 per SMSP), 16 lanes"},
        {256, 32, 8, "8 warps (2 per SMSP), 32 lanes"}, {256, 16, 8, "8 warps (2 per SMSP), 16 lanes"},
        {256, 32, 5, "warps 0-4 (SMSP0 has 2), 32 lanes"}};
    for (auto& f : cfg) {
        k<<<1, f.threads>>>(o, c, f.lanes, f.nw); cudaDeviceSynchronize();
        k<<<1, f.threads>>>(o, c, f.lanes, f.nw); cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-40s warp0: %.2f cycles per MUFU warp-instr (2048 instr)", f.name, h[0] / 2048.0);
        if (f.nw > 4) printf("   warp4: %.2f", h[4] / 2048.0);
        printf("\n");
    }
    return 0;
}
