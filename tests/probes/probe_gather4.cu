// Hardware probe (test infrastructure, not product): TMA tile::gather4 on sm_100a.
//  (1) layout: gather 4 arbitrary rows x 64 bf16 (128 B) into a SWIZZLE_128B smem tile and check
//      that row i of the gather lands at dst + i*128 with chunk c at ((c ^ (row&7)) << 4), i.e. the
//      same layout the cp.async producers write by hand;
//  (2) throughput: every SM fills 32 KB tiles (128 rows x 256 B) from random rows of an L2-resident
//      matrix, (a) with 16-byte cp.async from 128 threads, (b) with gather4 from 1 or 4 threads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_gather4 tests/probes/probe_gather4.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../chipmunk_b200/csrc/ptx.cuh"

using namespace cm;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    return (EncodeFn)fn;
}

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n" ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
        "r"(smem_u32(bar)) : "memory");
}

__global__ void layout_probe(const __grid_constant__ CUtensorMap map, const int* rows, uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, 8 * 4 * 128 * 2);          // 8 gathers x 4 rows x 128 B, two column halves
        for (int g = 0; g < 8; g++)
            for (int h = 0; h < 2; h++)
                tma_gather4(base + h * 4096 + g * 512, &map, &bar, h * 64, rows[4 * g], rows[4 * g + 1], rows[4 * g + 2], rows[4 * g + 3]);
    }
    mbar_wait(&bar, 0);
    const uint8_t* s = smem + (base - smem_u32(smem));
    for (int i = threadIdx.x; i < 8192 / 2; i += blockDim.x) out[i] = reinterpret_cast<const uint16_t*>(s)[i];
}

constexpr int TILE_ROWS = 128;
template <int MODE>   // 0: cp.async, 1: gather4 from 1 thread, 4: gather4 from 4 threads, 5: two 2-D TMA tile loads (64 cols x 128 rows)
__global__ void __launch_bounds__(160) bw_probe(const __grid_constant__ CUtensorMap map, const __nv_bfloat16* mat,
                                                const int* rows, int nrows_list, int iters, unsigned long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[4];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) mbar_init(&full[i], (MODE == 0 || MODE == 6 || MODE == 7) ? 128 : (MODE == 4 ? 4 : 1)); fence_mbar_init(); }
    __syncthreads();
    const int tid = threadIdx.x;
    unsigned long long t0 = clock64();
    if (tid >= 32) {   // 128 producer threads
        const int pt = tid - 32;
        for (int it = 0; it < iters; it++) {
            const int slot = it & 3;
            if (it >= 4) mbar_wait(&full[slot], ((it >> 2) - 1) & 1);    // previous fill of this slot landed
            const int* rl = rows + (((size_t)(blockIdx.x * 132 + it * 128) % (size_t)(nrows_list - 256)) & ~(size_t)3);
            const uint32_t dst = base + slot * 32768;
            if (MODE == 0) {
                const int chunk = pt & 15, rsub = pt >> 4;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    int r = rsub + 8 * i;
                    const __nv_bfloat16* src = mat + (size_t)rl[r] * 128 + chunk * 8;
                    cp_async_16(dst + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4), src);
                }
                cp_async_mbar_arrive_noinc(&full[slot]);
            } else if (MODE == 6 || MODE == 7) {
                // 6: all 16 loads of a thread in flight, then 16 stores; 7: same but 256-bit loads (8 per thread)
                const int chunk = pt & 15, rsub = pt >> 4;
                if (MODE == 6) {
                    uint4 v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        int r = rsub + 8 * i;
                        v[i] = __ldg(reinterpret_cast<const uint4*>(mat + (size_t)rl[r] * 128 + chunk * 8));
                    }
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        int r = rsub + 8 * i;
                        uint32_t a = dst + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "r"(v[i].x), "r"(v[i].y), "r"(v[i].z), "r"(v[i].w) : "memory");
                    }
                } else {
                    const int c32 = pt & 7, rs = pt >> 3;      // 8 x 32-byte pieces per row, 16 rows per pass
                    uint32_t v[8][8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        int r = rs + 16 * i;
                        const void* src = mat + (size_t)rl[r] * 128 + c32 * 16;
                        asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                                     : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]), "=r"(v[i][4]), "=r"(v[i][5]), "=r"(v[i][6]), "=r"(v[i][7]) : "l"(src));
                    }
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        int r = rs + 16 * i;
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {
                            int chunk = c32 * 2 + hh;
                            uint32_t a = dst + (chunk >> 3) * 16384 + r * 128 + (((chunk & 7) ^ (r & 7)) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "r"(v[i][4*hh]), "r"(v[i][4*hh+1]), "r"(v[i][4*hh+2]), "r"(v[i][4*hh+3]) : "memory");
                        }
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(&full[slot]);
            } else if (MODE == 5) {
                if (pt == 0) {
                    mbar_arrive_expect_tx(&full[slot], 32768);
                    const int row0 = (rl[0] & ~127);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(dst), "l"(&map), "r"(0), "r"(row0), "r"(smem_u32(&full[slot])) : "memory");
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3}], [%4];\n"
                                 ::"r"(dst + 16384), "l"(&map), "r"(64), "r"(row0), "r"(smem_u32(&full[slot])) : "memory");
                }
            } else {
                const int nthr = MODE == 1 ? 1 : 4;
                if (pt < nthr) {
                    mbar_arrive_expect_tx(&full[slot], 32768 / nthr);
                    for (int g = pt; g < 32; g += nthr) {
                        int4 r4 = *reinterpret_cast<const int4*>(rl + 4 * g);
                        tma_gather4(dst + g * 512, &map, &full[slot], 0, r4.x, r4.y, r4.z, r4.w);
                        tma_gather4(dst + 16384 + g * 512, &map, &full[slot], 64, r4.x, r4.y, r4.z, r4.w);
                    }
                }
            }
        }
        if (MODE == 0) cp_async_wait_all();
    }
    __syncthreads();
    if (MODE != 0 && MODE != 6 && MODE != 7 && tid == 32) {   // wait for the final fills
        for (int it = max(0, iters - 4); it < iters; it++) mbar_wait(&full[it & 3], (it >> 2) & 1);
    }
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    EncodeFn encode = get_encode();
    const int R = 65536;   // 65536 rows x 256 B = 16 MB: L2 resident
    std::vector<__nv_bfloat16> h((size_t)R * 128);
    for (size_t i = 0; i < h.size(); i++) h[i] = __float2bfloat16((float)((i * 7) % 251));
    __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap map;
    cuuint64_t gdim[2] = {128, (cuuint64_t)R};
    cuuint64_t gstr[1] = {256};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;

    // ---- (1) layout
    std::vector<int> rows(32);
    srand(5);
    for (int i = 0; i < 32; i++) rows[i] = rand() % R;
    int* drows; CK(cudaMalloc(&drows, 32 * 4)); CK(cudaMemcpy(drows, rows.data(), 32 * 4, cudaMemcpyHostToDevice));
    uint16_t* dout; CK(cudaMalloc(&dout, 8192));
    CK(cudaFuncSetAttribute(layout_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    layout_probe<<<1, 128, 16384>>>(map, drows, dout);
    CK(cudaDeviceSynchronize());
    std::vector<uint16_t> out(4096);
    CK(cudaMemcpy(out.data(), dout, 8192, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int row = 0; row < 32; row++)
        for (int half = 0; half < 2; half++)
            for (int c = 0; c < 8; c++)
                for (int e = 0; e < 8; e++) {
                    size_t soff = half * 4096 + row * 128 + ((c ^ (row & 7)) << 4) + e * 2;
                    uint16_t got = out[soff / 2];
                    __nv_bfloat16 want = h[(size_t)rows[row] * 128 + half * 64 + c * 8 + e];
                    if (got != *reinterpret_cast<uint16_t*>(&want)) bad++;
                }
    printf("gather4 layout (row i at +i*128 B, chunk ^ (row&7)): %s (bad=%d)\n", bad ? "FAIL" : "PASS", bad);

    // ---- (2) throughput
    const int NL = 1 << 20;
    std::vector<int> rl(NL);
    for (int i = 0; i < NL; i++) rl[i] = rand() % R;
    int* drl; CK(cudaMalloc(&drl, NL * 4)); CK(cudaMemcpy(drl, rl.data(), NL * 4, cudaMemcpyHostToDevice));
    unsigned long long* dcy; CK(cudaMalloc(&dcy, 148 * 8));
    const int iters = 2000;
    auto run = [&](auto kern, const char* name, const CUtensorMap& map, int grid) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768 + 1024));
        kern<<<grid, 160, 4 * 32768 + 1024>>>(map, d, drl, NL, iters, dcy);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        kern<<<grid, 160, 4 * 32768 + 1024>>>(map, d, drl, NL, iters, dcy);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> cy(grid);
        CK(cudaMemcpy(cy.data(), dcy, grid * 8, cudaMemcpyDeviceToHost));
        double avg = 0; for (auto c : cy) avg += c; avg /= grid;
        double bytes = (double)iters * 32768 * grid;
        printf("%-28s grid=%3d %.3f ms  %.0f GB/s aggregate  %.1f B/clk/SM  (%.0f cycles per 32 KB tile)\n", name, grid, ms, bytes / ms / 1e6,
               (double)iters * 32768 / avg, avg / iters);
    };
    CUtensorMap map_tile;
    cuuint32_t box_tile[2] = {64, 128};
    encode(&map_tile, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstr, box_tile, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int grid : {8, 37, 74, 148}) run(bw_probe<0>, "cp.async 16B x128 threads", map, grid);
    for (int grid : {8, 148}) run(bw_probe<6>, "LDG.128 -> STS.128 x128 thr", map, grid);
    for (int grid : {8, 148}) run(bw_probe<7>, "LDG.256 -> STS.128 x128 thr", map, grid);
    for (int grid : {8, 37, 74, 148}) run(bw_probe<5>, "TMA 2-D tiles 2 x 16 KB", map_tile, grid);
    run(bw_probe<1>, "gather4 x1 thread", map, 148);
    run(bw_probe<4>, "gather4 x4 threads", map, 148);
    return 0;
}
