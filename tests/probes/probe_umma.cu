// Hardware probe (test infrastructure, not product): verifies on a real B200 the operand /
// accumulator layouts the product kernels assume for tcgen05.mma (SS and TS, M=128 and M=64,
// K-major and MN-major B) and discovers the thread<->TMEM mapping of the 16x256b / 16x128b
// tcgen05.ld/st shapes. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma
// tests/probes/probe_umma.cu ; run on the GPU box; prints PASS/FAIL lines and mapping tables.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../../chipmunk_b200/csrc/ptx.cuh"

using namespace cm;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

struct Cfg {
    int M, N, K;       // K multiple of 64 (<=128), N <= 256
    int a_tmem;        // 0: A from smem (K-major SW128), 1: A from TMEM (packed bf16x2)
    int b_mn;          // 0: B K-major [N x K], 1: B MN-major stored as [K x N]
};

// smem: A at 0 (up to 128x128x2 = 32 KB), B at 32 KB (up to 256 x 128 x 2 = 64 KB)
__global__ void __launch_bounds__(128, 1)
probe_mma(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
          float* __restrict__ out /*[128][N]*/, Cfg c) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t sA = sbase, sB = sbase + 32768;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 512); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;

    // ---- stage A (K-major) into smem: chunk kc (64 k) at sA + kc*(M*128)
    if (!c.a_tmem) {
        for (int i = tid; i < c.M * (c.K / 8); i += 128) {
            int r = i / (c.K / 8), c16g = i % (c.K / 8);
            int kc = c16g / 8, c16 = c16g % 8;
            uint4 v = *reinterpret_cast<const uint4*>(A + (size_t)r * c.K + c16g * 8);
            uint32_t dst = sA + kc * (c.M * 128) + sw128_off(r, c16);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
        }
    }
    // ---- stage B
    if (!c.b_mn) {  // K-major: rows n, chunk kc at sB + kc*(N*128)
        for (int i = tid; i < c.N * (c.K / 8); i += 128) {
            int r = i / (c.K / 8), c16g = i % (c.K / 8);
            int kc = c16g / 8, c16 = c16g % 8;
            uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)r * c.K + c16g * 8);
            uint32_t dst = sB + kc * (c.N * 128) + sw128_off(r, c16);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
        }
    } else {  // MN-major: global B is [N x K]; smem rows are k, 64-wide n chunks at sB + nc*(K*128)
        for (int i = tid; i < c.K * c.N; i += 128) {
            int k = i / c.N, n = i % c.N;
            int nc = n / 64, nn = n % 64;
            uint32_t dst = sB + nc * (c.K * 128) + sw128_off(k, nn / 8) + (nn % 8) * 2;
            uint16_t v = reinterpret_cast<const uint16_t*>(B)[(size_t)n * c.K + k];
            asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(dst), "h"(v));
        }
    }
    // ---- A into TMEM (packed bf16x2), columns [256, 256+K/2)
    if (c.a_tmem) {
        int row = -1;
        if (c.M == 128) row = tid;
        else if (lane < 16) row = warp * 16 + lane;   // M=64: lanes (m%16)+32*(m/16)
        for (int cb = 0; cb < c.K / 2; cb += 16) {
            uint32_t r[16];
            for (int j = 0; j < 16; j++) {
                uint32_t v = 0;
                if (row >= 0) v = reinterpret_cast<const uint32_t*>(A + (size_t)row * c.K)[cb + j];
                r[j] = v;
            }
            tmem_st_32x32b_x16(tm + ((uint32_t)(warp * 32) << 16) + 256 + cb, r);
        }
        tmem_st_wait();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

    if (tid == 0) {
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_bf16(c.M, c.N, 0, c.b_mn);
        for (int k16 = 0; k16 < c.K / 16; k16++) {
            int kc = k16 / 4, kk = k16 % 4;
            uint64_t bd;
            if (!c.b_mn) bd = umma_smem_desc(sB + kc * (c.N * 128) + kk * 32, 16, 1024);
            else         bd = umma_smem_desc(sB + k16 * 2048, c.K * 128, 1024);
            if (!c.a_tmem) {
                uint64_t ad = umma_smem_desc(sA + kc * (c.M * 128) + kk * 32, 16, 1024);
                umma_ss(tm, ad, bd, idesc, k16 > 0);
            } else {
                umma_ts(tm, tm + 256 + k16 * 8, bd, idesc, k16 > 0);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int cb = 0; cb < c.N; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + cb, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) out[(size_t)tid * c.N + cb + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- shape discovery: fill TMEM[lane][col] = lane*1024+col via 32x32b, read with other shapes
__global__ void __launch_bounds__(128, 1) probe_shapes(uint32_t* __restrict__ out) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const uint32_t wbase = tm + ((uint32_t)(warp * 32) << 16);
    uint32_t r[32];
    for (int j = 0; j < 32; j++) r[j] = tid * 1024 + j;
    tmem_st_32x32b_x32(wbase, r);
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    // (a) 16x256b.x2 at lane offset 0 -> 8 regs
    uint32_t a[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                 : "r"(wbase));
    tmem_ld_wait();
    for (int j = 0; j < 8; j++) out[0 * 128 * 8 + tid * 8 + j] = a[j];
    // (b) 16x256b.x2 at lane offset 16
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7])
                 : "r"(wbase + (16u << 16)));
    tmem_ld_wait();
    for (int j = 0; j < 8; j++) out[1 * 128 * 8 + tid * 8 + j] = a[j];
    // (c) 16x128b.x2 -> 4 regs
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(wbase));
    tmem_ld_wait();
    for (int j = 0; j < 8; j++) out[2 * 128 * 8 + tid * 8 + j] = j < 4 ? a[j] : 0xffffffffu;
    // (d) 16x64b.x4 -> 4 regs
    asm volatile("tcgen05.ld.sync.aligned.16x64b.x4.b32 {%0,%1,%2,%3}, [%4];\n"
                 : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(wbase));
    tmem_ld_wait();
    for (int j = 0; j < 8; j++) out[3 * 128 * 8 + tid * 8 + j] = j < 4 ? a[j] : 0xffffffffu;
    // (e) st 16x128b.x2 of a recognisable pattern into cols 32.., read back with 32x32b
    uint32_t s[4];
    for (int j = 0; j < 4; j++) s[j] = 0x80000000u | (tid << 8) | j;
    tc_fence_before_sync();
    __syncthreads();
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(wbase + 32), "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]));
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    uint32_t b[16];
    tmem_ld_32x32b_x16(wbase + 32, b);
    tmem_ld_wait();
    for (int j = 0; j < 8; j++) out[4 * 128 * 8 + tid * 8 + j] = b[j];
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static bool run_cfg(const Cfg& c, const char* name) {
    std::vector<__nv_bfloat16> hA((size_t)c.M * c.K), hB((size_t)c.N * c.K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(1234);
    for (size_t i = 0; i < hA.size(); i++) { float v = bf((rand() % 17 - 8) / 8.0f); hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); i++) { float v = bf((rand() % 17 - 8) / 8.0f); hB[i] = __float2bfloat16(v); fB[i] = v; }
    std::vector<float> ref((size_t)c.M * c.N);
    for (int m = 0; m < c.M; m++)
        for (int n = 0; n < c.N; n++) {
            float s = 0;
            for (int k = 0; k < c.K; k++) s += fA[(size_t)m * c.K + k] * fB[(size_t)n * c.K + k];
            ref[(size_t)m * c.N + n] = s;
        }
    __nv_bfloat16 *dA, *dB; float* dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, 128 * c.N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xff, 128 * c.N * 4));
    CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    probe_mma<<<1, 128, 100 * 1024>>>(dA, dB, dO, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-34s LAUNCH FAIL: %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> out((size_t)128 * c.N);
    CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
    // hypothesised row->lane map
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < c.M; m++) {
        int lane = c.M == 128 ? m : (m % 16) + 32 * (m / 16);
        for (int n = 0; n < c.N; n++) {
            double d = fabs(out[(size_t)lane * c.N + n] - ref[(size_t)m * c.N + n]);
            if (d > maxerr) maxerr = d;
            if (d > 1e-3) bad++;
        }
    }
    printf("%-34s %s (bad=%d maxerr=%g)\n", name, bad ? "FAIL" : "PASS", bad, maxerr);
    if (bad) {
        // discovery: for each logical row find the lane whose N-vector matches
        printf("   row->lane discovery (first 70 rows):");
        for (int m = 0; m < c.M && m < 70; m++) {
            int found = -1;
            for (int l = 0; l < 128; l++) {
                bool ok = true;
                for (int n = 0; n < c.N && ok; n++) ok = fabs(out[(size_t)l * c.N + n] - ref[(size_t)m * c.N + n]) < 1e-3;
                if (ok) { found = l; break; }
            }
            printf(" %d", found);
        }
        printf("\n   sample out[lane0][0..7]:");
        for (int n = 0; n < 8; n++) printf(" %g", out[n]);
        printf("\n   ref[0][0..7]:");
        for (int n = 0; n < 8; n++) printf(" %g", ref[n]);
        printf("\n");
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return bad == 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device: %s sm_%d%d SMs=%d smem/block optin=%zu\n", p.name, p.major, p.minor, p.multiProcessorCount, p.sharedMemPerBlockOptin);
    int ok = 1;
    ok &= run_cfg({128, 128, 128, 0, 0}, "SS  M128 N128 K128 B:K-major");
    ok &= run_cfg({128, 128, 128, 0, 1}, "SS  M128 N128 K128 B:MN-major");
    ok &= run_cfg({128, 128, 128, 1, 1}, "TS  M128 N128 K128 B:MN-major");
    ok &= run_cfg({128, 256, 64, 0, 0},  "SS  M128 N256 K64  B:K-major");
    ok &= run_cfg({128, 256, 128, 0, 1}, "SS  M128 N256 K128 B:MN-major");
    ok &= run_cfg({64, 128, 128, 0, 0},  "SS  M64  N128 K128 B:K-major");
    ok &= run_cfg({64, 128, 128, 0, 1},  "SS  M64  N128 K128 B:MN-major");
    ok &= run_cfg({64, 128, 128, 1, 1},  "TS  M64  N128 K128 B:MN-major");
    ok &= run_cfg({128, 192, 128, 0, 0}, "SS  M128 N192 K128 B:K-major");

    uint32_t* dS; CK(cudaMalloc(&dS, 5 * 128 * 8 * 4));
    probe_shapes<<<1, 128>>>(dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe_shapes LAUNCH FAIL: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<uint32_t> hs(5 * 128 * 8);
    CK(cudaMemcpy(hs.data(), dS, hs.size() * 4, cudaMemcpyDeviceToHost));
    const char* names[5] = {"ld 16x256b.x2 @lane0", "ld 16x256b.x2 @lane16", "ld 16x128b.x2 @lane0", "ld 16x64b.x4 @lane0", "st 16x128b.x2 -> 32x32b readback"};
    for (int s = 0; s < 5; s++) {
        printf("== %s : thread t reg j -> (lane,col) [warp 0 and warp 1 thread 0..3]\n", names[s]);
        for (int t = 0; t < 36; t++) {
            if (t >= 32 && s != 0) break;
            printf("  t%-3d:", t);
            for (int j = 0; j < 8; j++) {
                uint32_t v = hs[(size_t)s * 1024 + t * 8 + j];
                if (v == 0xffffffffu) continue;
                if (s < 4) printf(" (%u,%u)", v / 1024, v % 1024);
                else if (v & 0x80000000u) printf(" [t%u r%u]", (v >> 8) & 0xff, v & 0xff);
                else printf(" (%u,%u)", v / 1024, v % 1024);
            }
            printf("\n");
        }
    }
    printf("PROBE %s\n", ok ? "ALL-PASS" : "HAS-FAIL");
    return 0;
}
