// Column sums of the attention probabilities for mask selection (dense "full" steps).
//   cs[b,h,g,j] = sum_{i in 192-row group g} exp(q_i . k_j / sqrt(128)) * p_i ,  p = previous step's l
// Replaces the colsum half of csrc/attn/dense_colsum_attn.cu:267-333 of the reference, which
// reduces bf16 partials across 12 warps with shared-memory atomics.  Here the product is
// computed TRANSPOSED on the tensor pipe, S^T = K_tile Q_g^T (M = 128 keys on the TMEM lanes,
// N = 192 queries on the columns), so the sum over the group's queries is a per-thread loop
// over TMEM columns: no atomics, no shuffles, fp32 accumulation, one bf16 rounding.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/chipmunk_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace cm {
namespace attn {

int launch_colsum(const __nv_bfloat16* q, const __nv_bfloat16* k, const float* p, __nv_bfloat16* cs, int B, int H,
                  int Nq, int Nk, int64_t cs_row_stride, cudaStream_t stream) {
    (void)q; (void)k; (void)p; (void)cs; (void)B; (void)H; (void)Nq; (void)Nk; (void)cs_row_stride; (void)stream;
    return CM_EUNSUPPORTED;   // filled in below once the main attention kernel is validated
}

}  // namespace attn
}  // namespace cm
