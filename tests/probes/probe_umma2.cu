// Hardware probe (test infrastructure, not product): tcgen05.mma.cta_group::2 on a CTA pair.
// Verifies the operand split the 2-CTA attention kernel assumes:
//   D[256 x N] = A[256 x K] * B[N x K]^T,  A rows 128r..128r+127 in CTA r (smem K-major, or TMEM),
//   B split along N: CTA r holds N/2 rows (K-major case) or N/2 columns (MN-major case),
//   D rows 128r.. in CTA r's TMEM.  Also exercises commit multicast and the cluster barrier.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma2 tests/probes/probe_umma2.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "../../chipmunk_b200/csrc/ptx.cuh"

using namespace cm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
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
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

struct Cfg { int N, K, a_tmem, b_mn; };   // M = 256

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe2(const __nv_bfloat16* __restrict__ A /*[256][K]*/, const __nv_bfloat16* __restrict__ B /*[N][K]*/, float* __restrict__ out /*[256][N]*/, Cfg c) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t sbase = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t sA = sbase, sB = sbase + 32768;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int Nh = c.N / 2;

    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc2(&tmem_base_s, 512);
    tc_fence_before_sync();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tm = tmem_base_s;

    // ---- A: this CTA's 128 rows
    const __nv_bfloat16* Ar = A + (size_t)rank * 128 * c.K;
    if (!c.a_tmem) {
        for (int i = tid; i < 128 * (c.K / 8); i += 128) {
            int r = i / (c.K / 8), c16g = i % (c.K / 8), kc = c16g / 8, c16 = c16g % 8;
            uint4 v = *reinterpret_cast<const uint4*>(Ar + (size_t)r * c.K + c16g * 8);
            uint32_t dst = sA + kc * (128 * 128) + sw128_off(r, c16);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
        }
    } else {
        for (int cb = 0; cb < c.K / 2; cb += 16) {
            uint32_t r[16];
            for (int j = 0; j < 16; j++) r[j] = reinterpret_cast<const uint32_t*>(Ar + (size_t)tid * c.K)[cb + j];
            tmem_st_32x32b_x16(tm + ((uint32_t)(warp * 32) << 16) + 256 + cb, r);
        }
        tmem_st_wait();
    }
    // ---- B: this CTA's half
    if (!c.b_mn) {      // K-major: rows n = rank*Nh + r, r < Nh; chunk kc at sB + kc*(Nh*128)
        for (int i = tid; i < Nh * (c.K / 8); i += 128) {
            int r = i / (c.K / 8), c16g = i % (c.K / 8), kc = c16g / 8, c16 = c16g % 8;
            uint4 v = *reinterpret_cast<const uint4*>(B + (size_t)(rank * Nh + r) * c.K + c16g * 8);
            uint32_t dst = sB + kc * (Nh * 128) + sw128_off(r, c16);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
        }
    } else {            // MN-major: smem rows are k (all K of them), this CTA's Nh columns n = rank*Nh + nn, 64-wide chunks
        for (int i = tid; i < c.K * Nh; i += 128) {
            int k = i / Nh, nn = i % Nh, nc = nn / 64, n64 = nn % 64;
            uint32_t dst = sB + nc * (c.K * 128) + sw128_off(k, n64 / 8) + (n64 % 8) * 2;
            uint16_t v = reinterpret_cast<const uint16_t*>(B)[(size_t)(rank * Nh + nn) * c.K + k];
            asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(dst), "h"(v));
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    cluster_sync_all();          // both CTAs' operands are in place

    if (rank == 0 && tid == 0) {
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_bf16(256, c.N, 0, c.b_mn);
        for (int k16 = 0; k16 < c.K / 16; k16++) {
            int kc = k16 / 4, kk = k16 % 4;
            uint64_t bd = !c.b_mn ? umma_smem_desc(sB + kc * (Nh * 128) + kk * 32, 16, 1024)
                                  : umma_smem_desc(sB + k16 * 2048, c.K * 128, 1024);
            if (!c.a_tmem) umma2_ss(tm, umma_smem_desc(sA + kc * (128 * 128) + kk * 32, 16, 1024), bd, idesc, k16 > 0);
            else umma2_ts(tm, tm + 256 + k16 * 8, bd, idesc, k16 > 0);
        }
        umma2_commit_mc(&bar, 0x3);        // arrive on `bar` in both CTAs
    }
    mbar_wait(&bar, 0);
    tc_fence_after_sync();
    for (int cb = 0; cb < c.N; cb += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + cb, r);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++) out[(size_t)(rank * 128 + tid) * c.N + cb + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync();
    cluster_sync_all();
    if (warp == 0) tmem_dealloc2(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static bool run_cfg(const Cfg& c, const char* name) {
    const int M = 256;
    std::vector<__nv_bfloat16> hA((size_t)M * c.K), hB((size_t)c.N * c.K);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(4321);
    for (size_t i = 0; i < hA.size(); i++) { float v = bf((rand() % 17 - 8) / 8.0f); hA[i] = __float2bfloat16(v); fA[i] = v; }
    for (size_t i = 0; i < hB.size(); i++) { float v = bf((rand() % 17 - 8) / 8.0f); hB[i] = __float2bfloat16(v); fB[i] = v; }
    __nv_bfloat16 *dA, *dB; float* dO;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dO, (size_t)M * c.N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dO, 0xff, (size_t)M * c.N * 4));
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    probe2<<<2, 128, 100 * 1024>>>(dA, dB, dO, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-40s LAUNCH FAIL: %s\n", name, cudaGetErrorString(e)); exit(2); }
    std::vector<float> out((size_t)M * c.N);
    CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < c.N; n++) {
            float s = 0;
            for (int k = 0; k < c.K; k++) s += fA[(size_t)m * c.K + k] * fB[(size_t)n * c.K + k];
            double d = fabs(out[(size_t)m * c.N + n] - s);
            if (d > maxerr) maxerr = d;
            if (!(d <= 1e-3)) bad++;
        }
    printf("%-40s %s (bad=%d maxerr=%g)\n", name, bad ? "FAIL" : "PASS", bad, maxerr);
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
    return bad == 0;
}

int main() {
    int ok = 1;
    ok &= run_cfg({128, 128, 0, 0}, "2CTA SS M256 N128 K128 B:K-major (S=QK^T)");
    ok &= run_cfg({128, 128, 1, 1}, "2CTA TS M256 N128 K128 B:MN-major (O=PV)");
    ok &= run_cfg({128, 128, 0, 1}, "2CTA SS M256 N128 K128 B:MN-major");
    ok &= run_cfg({256, 64, 0, 0},  "2CTA SS M256 N256 K64  B:K-major");
    ok &= run_cfg({256, 128, 1, 1}, "2CTA TS M256 N256 K128 B:MN-major");
    printf("PROBE2 %s\n", ok ? "ALL-PASS" : "HAS-FAIL");
    return 0;
}
