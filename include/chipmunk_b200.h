/* chipmunk_b200 — C ABI of the sm_100a column-sparse DiT kernels.
 *
 * One entry point per operator the reference registers in csrc/chipmunk.cpp:47-60
 * (`TORCH_LIBRARY(chipmunk, m)`), plus the bit-mask codec the reference implements in
 * src/chipmunk/ops/bitpack.py.  Plain pointers, sizes and element strides only: no torch
 * types cross this boundary.  Every function
 *   - is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default),
 *   - never allocates, never synchronises, never throws,
 *   - returns 0 on success, a positive cudaError_t for a launch/driver error, or a negative
 *     CM_E* code for an argument the kernel cannot serve (cm_strerror() explains both).
 * All paths cited below are relative to the reference tree (/root/reference).
 */
#ifndef CHIPMUNK_B200_H
#define CHIPMUNK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_ABI_VERSION 2

enum {
    CM_OK = 0,
    CM_EINVAL = -1,   /* bad shape / multiple / null pointer */
    CM_EALIGN = -2,   /* pointer or stride not 16-byte aligned where the kernel needs it */
    CM_EUNSUPPORTED = -3, /* dtype / head_dim the kernels do not implement */
    CM_EARCH = -4     /* device is not sm_100 */
};

/* dtype tags for the kernels that accept several element types */
enum { CM_BF16 = 0, CM_F16 = 1, CM_F32 = 2 };

int cm_abi_version(void);
const char* cm_strerror(int code);
/* number of SMs of the current device (grid sizing is internal; exported for benches) */
int cm_sm_count(void);

/* ---------------------------------------------------------------------------------------
 * Column-sparse "delta" attention.
 * Replaces chipmunk::csp_attn      (csrc/attn/csp_attn.cu:315-422, schema chipmunk.cpp:51)
 *      and chipmunk::csp_128_attn  (csrc/attn/csp_128_attn.cu:355-460, schema chipmunk.cpp:52).
 * For every (b, h, group g of 192 query rows):
 *     delta = bf16( o_scale * softmax(Q_g K[idx]^T / sqrt(128)) V[idx] ),
 *     idx   = indices[b,h,g, 0:counts[b,h,g]]
 *     accumulate != 0 : o[b,h,rows of g] = bf16(o + delta)      (csp_attn, o_scale = +1/-1)
 *     accumulate == 0 : o[b,h,rows of g] = delta                (csp_128_attn)
 * q,o: [B,H,Nq,128] bf16; k,v: [B,H,Nk,128] bf16; last-dim stride 1, the other three strides
 * given in ELEMENTS as {batch, head, row} and multiples of 8 (16 bytes).
 * indices: int32 [B,H,ceil(Nq/192), idx_row_stride] contiguous; counts: int32 [B,H,ceil(Nq/192)].
 * counts may be any value in [0, idx_row_stride]; 0 means "no contribution".
 */
int cm_csp_attn(const void* q, const void* k, const void* v, void* o,
                const int32_t* indices, const int32_t* counts,
                int B, int H, int Nq, int Nk,
                const int64_t q_strides[3], const int64_t k_strides[3],
                const int64_t v_strides[3], const int64_t o_strides[3],
                int64_t idx_row_stride, int o_scale, int accumulate, void* stream);

/* The same kernel with the delta add-back written OUT OF PLACE:
 *     o[b,h,rows of g] = bf16(cache + delta),   cache != o, cache is not modified.
 * Replaces the `o = o_cache.clone(); csp_attn(q, k, v, o, ...)` pair of the reference's sparse step
 * (src/chipmunk/modules/attn.py:186-190) and the `o_cache = o - csp(...)` of its full step (:165-170, o_scale = -1):
 * one pass over the cache instead of clone + read-modify-write.  Numerically identical to cm_csp_attn(accumulate=1)
 * on a copy of the cache (delta is rounded to bf16 first, then one bf16 add).
 */
int cm_csp_attn_add(const void* q, const void* k, const void* v, const void* cache, void* o,
                    const int32_t* indices, const int32_t* counts,
                    int B, int H, int Nq, int Nk,
                    const int64_t q_strides[3], const int64_t k_strides[3],
                    const int64_t v_strides[3], const int64_t cache_strides[3], const int64_t o_strides[3],
                    int64_t idx_row_stride, int o_scale, void* stream);

/* cm_csp_attn_add fused with the head-parallel all-gather of O (replaces the two all_to_all + all_gather of
 * examples/hunyuan/hyvideo/modules/head_parallel.py:42-115 for the sparse steps).  `o_local` is this GPU's slice of a
 * SYMMETRIC buffer (same layout on every GPU of one NVSwitch domain) and `multicast_delta_bytes` the distance from
 * that buffer's local address to its NVLS multicast alias: every output row is written once, with multimem.st, and
 * the switch replicates it into all GPUs' copies while the remaining tiles are still being computed.  The caller
 * owns the cross-GPU barrier after the kernel (and before the buffer is overwritten again).
 */
int cm_csp_attn_add_bcast(const void* q, const void* k, const void* v, const void* cache, void* o_local,
                          int64_t multicast_delta_bytes,
                          const int32_t* indices, const int32_t* counts,
                          int B, int H, int Nq, int Nk,
                          const int64_t q_strides[3], const int64_t k_strides[3],
                          const int64_t v_strides[3], const int64_t cache_strides[3], const int64_t o_strides[3],
                          int64_t idx_row_stride, int o_scale, void* stream);

/* The same fused gather for an NVLink domain WITHOUT a multicast object (or when NVLS is disabled): every output row is
 * stored with plain 16-byte stores into each peer's copy of the symmetric buffer, `peer_delta_bytes[p]` = (address of GPU
 * p's buffer in this process' address space) - (address of the local buffer), own copy included (delta 0), n_peers <= 8.
 * n_peers x the store traffic of the multicast variant on this GPU's NVLink egress, still hidden behind the remaining tiles. */
int cm_csp_attn_add_peers(const void* q, const void* k, const void* v, const void* cache, void* o_local,
                          const int64_t* peer_delta_bytes, int n_peers,
                          const int32_t* indices, const int32_t* counts,
                          int B, int H, int Nq, int Nk,
                          const int64_t q_strides[3], const int64_t k_strides[3],
                          const int64_t v_strides[3], const int64_t cache_strides[3], const int64_t o_strides[3],
                          int64_t idx_row_stride, int o_scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense attention with the statistics the sparse steps need.
 * Replaces chipmunk::dense_attn        (csrc/attn/dense_attn.cu:246-371, schema chipmunk.cpp:54)
 *      and chipmunk::dense_colsum_attn (csrc/attn/dense_colsum_attn.cu:521-668, chipmunk.cpp:55).
 *   o = softmax(QK^T/sqrt(128)) V                      [B,H,Nq,128] bf16 (contiguous)
 *   l[b,h,i] = 1 / sum_j exp(s_ij/sqrt(128))           [B,H,Nq] fp32
 *   if cs != NULL (then p != NULL, p = previous step's l, [B,H,Nq] fp32):
 *   cs[b,h,g,j] = sum_{i in group g} exp(s_ij/sqrt(128)) * p_i   [B,H,ceil(Nq/192),cs_row_stride] bf16
 * q,k,v contiguous [B,H,N,128] bf16.
 */
int cm_dense_attn(const void* q, const void* k, const void* v, void* o, float* l,
                  void* cs, const float* p, int B, int H, int Nq, int Nk,
                  int64_t cs_row_stride, void* stream);
/* The same operators for strided q/k/v/o views ({batch, head, row} element strides, multiples of 8; FLUX hands
 * q, k, v as views of one fused projection buffer, examples/flux layers.py:298), computed in ONE pass: o, l and the
 * column sums come out of a single walk over Q K^T, as in the reference's dense_colsum_attn.cu:205-341.
 * cs (if not NULL) is zero-initialised here; cs_row_stride >= Nk and a multiple of 8. */
int cm_dense_attn_strided(const void* q, const void* k, const void* v, void* o, float* l, void* cs, const float* p,
                          int B, int H, int Nq, int Nk,
                          const int64_t q_strides[3], const int64_t k_strides[3],
                          const int64_t v_strides[3], const int64_t o_strides[3],
                          int64_t cs_row_stride, void* stream);

/* ---------------------------------------------------------------------------------------
 * Column-sparse MLP, first GEMM.
 * Replaces chipmunk::csp_mlp_mm1 (csrc/mlp/csp_mlp_mm1.cu:625-702, schema chipmunk.cpp:47).
 * For token block mb (128 rows) and packed column j < counts[mb], f = indices[mb*idx_stride+j]:
 *     c[m, j] = bf16( gelu_tanh(a[m,:] . w1[f,:] + bias[f]) - pa_T[f, m] )
 * a [M,K], w1 [F,K], c [M,F] (packed), bias [F], pa_T [F,M], all bf16 contiguous.
 * M % 128 == 0, K % 64 == 0, counts[mb] % 16 == 0.
 * update_pa != 0 additionally performs the reference's later scatter step in the same
 * epilogue: pa_T[f, m] = bf16(pa_T[f, m] + c[m, j])   (csrc/indexed_io/scatter_add.cu:50-64).
 */
int cm_csp_mlp_mm1(const void* a, const void* w1, void* c, const void* bias, void* pa_T,
                   const int32_t* indices, const int32_t* counts,
                   int M, int K, int F, int64_t idx_stride, int update_pa, void* stream);

/* Column-sparse MLP, second GEMM (+ optional scatter).
 * Replaces chipmunk::csp_mlp_mm2_and_scatter_add
 * (csrc/mlp/csp_mlp_mm2_and_scatter_add.cu:96-259, schema chipmunk.cpp:49) and the Triton
 * kernel it launches (src/chipmunk/triton/csp_mlp_mm2.py:24-109).
 *     out[m, :] = bf16( bf16(packed[m, 0:cnt] @ w2_T[idx[mb,0:cnt], :]) + out[m, :] )
 *     do_scatter: pa_T[idx[mb,c], m] = bf16(pa_T[idx[mb,c], m] + packed[m, c]),  c < cnt
 * packed [M,F], w2_T [F,N], out [M,N], pa_T [F,M] bf16 contiguous; N % 256 == 0.
 */
int cm_csp_mlp_mm2(const void* packed, const void* w2_T, void* out, void* pa_T,
                   const int32_t* indices, const int32_t* counts,
                   int M, int F, int N, int64_t idx_stride, int do_scatter, void* stream);

/* Replaces chipmunk::csp_scatter_add (csrc/indexed_io/scatter_add.cu:74-154, chipmunk.cpp:59). */
int cm_csp_scatter_add(const void* packed, void* pa_T, const int32_t* indices,
                       const int32_t* counts, int M, int F, int64_t idx_stride, void* stream);

/* ---------------------------------------------------------------------------------------
 * Index selection / packing (integer path, bit-exact).
 */
/* Replaces chipmunk::mask_to_indices (csrc/indexed_io/mask_to_indices.cu:91-143, chipmunk.cpp:60).
 * mask: bool bytes [rows, n]; indices: int32 [rows, pad_n] (caller-allocated, entries past
 * counts[r] untouched); counts: int32 [rows] = popcount rounded up to `multiple_of`.
 * Emission order identical to the reference: set columns sorted by (col % 32, col), then the
 * first unset columns ascending as padding. */
int cm_mask_to_indices(const uint8_t* mask, int32_t* indices, int32_t* counts,
                       int64_t rows, int n, int pad_n, int multiple_of, void* stream);
/* Same result from the bit-packed mask (little-endian bit i of the flat mask, as produced by
 * cm_bitpack / src/chipmunk/ops/bitpack.py:4-41): fuses bitunpack + mask_to_indices
 * (src/chipmunk/modules/attn.py:173-176). */
int cm_bitmask_to_indices(const uint8_t* packed, int32_t* indices, int32_t* counts,
                          int64_t rows, int n, int pad_n, int multiple_of, void* stream);

/* Column selection of a full attention step in one kernel.  Replaces `random_and_topk` + `bitpack` +
 * `mask_to_indices` (src/chipmunk/modules/attn.py:76-84,132-139) and the `torch.topk` of the uncompressed path
 * (:141-150).  For every row r of cs [rows, cs_row_stride] bf16 (rows = B*H*G, the column sums of one query group):
 *     keep[c]  = c among the k largest of cs[r, 0:n]  (exactly k: equal values at the threshold are taken lowest
 *                column first; torch.topk leaves that choice unspecified)
 *              | hash(seed, r, c) < random_prob                      (the reference draws torch.randint(0,100)==0)
 *     keep     = (keep & group_is_sparse[g]) | static_words[g]       g = r % static_rows   (either may be NULL)
 * static_words: [static_rows, static_stride_words] uint32, bit c%32 of word c/32 = column c (rows padded to words).
 * Outputs (each optional, at least one): packed_out = the flat little-endian bit mask bitpack() would produce
 * (4-byte aligned, room for ceil(rows*n/32) words; zero-initialised here when rows share words);
 * indices [rows, pad_n] / counts [rows] exactly as cm_mask_to_indices would emit them from that mask. */
int cm_select_columns(const void* cs, int64_t cs_row_stride, int64_t rows, int n, int k, float random_prob,
                      uint64_t seed, const uint32_t* static_words, int64_t static_stride_words, int static_rows,
                      const uint8_t* group_is_sparse, uint8_t* packed_out, int32_t* indices, int32_t* counts,
                      int pad_n, int multiple_of, void* stream);

/* Replaces chipmunk::topk_indices (csrc/indexed_io/topk_indices.cu:145-222, chipmunk.cpp:58).
 * act: [B*R, C] of `dtype`; indices int32 [B*R, C]; counts int32 [B*R].
 * Threshold = ascending-sorted first min(C,1024) values at position int(1024*sparsity);
 * keep act >= threshold or curand_uniform < random_amount (same XORWOW seeding and draw
 * order as the reference); kept columns are emitted in ASCENDING order (the reference's
 * order is nondeterministic), padded to `multiple_of` with the first rejected columns. */
int cm_topk_indices(const void* act, int dtype, int32_t* indices, int32_t* counts,
                    int B, int R, int C, float sparsity, int multiple_of,
                    float random_amount, void* stream);

/* Replaces chipmunk::copy_indices (csrc/indexed_io/copy_indices.cu:82-163, chipmunk.cpp:57).
 * dst[b, r, idx] = src[b, r, idx] for idx in indices[b, r / Rr, 0:counts[b, r / Rr]];
 * src/dst [B, M*Rr, F] of elem_size bytes (2 or 4); indices [B,M,F]; counts [B,M]. */
int cm_copy_indices(const void* src, void* dst, int elem_size, const int32_t* indices,
                    const int32_t* counts, int B, int M, int Rr, int F, void* stream);

/* Token reordering: dst[o, i, :] = src[o, perm[i], :] for o < outer, i < n_dst; rows of row_bytes bytes, src has n_src
 * rows per outer index.  One gather replaces the chains of einops rearranges of the reference's 2-level patchify
 * (src/chipmunk/ops/patch.py:7-80) and 3-D voxel chunking (src/chipmunk/ops/voxel.py:9-99); the permutation itself is
 * built once per shape on the host side (chipmunk_b200/ops/{patch,voxel}.py). */
int cm_gather_rows(const void* src, void* dst, const int32_t* perm, int64_t outer, int64_t n_src, int64_t n_dst,
                   int64_t row_bytes, void* stream);

/* Replace bitpack / bitunpack (src/chipmunk/ops/bitpack.py:4-69): n mask bytes <-> ceil(n/8)
 * packed bytes, little-endian bit order. */
int cm_bitpack(const uint8_t* mask, uint8_t* packed, int64_t n, void* stream);
int cm_bitunpack(const uint8_t* packed, uint8_t* mask, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHIPMUNK_B200_H */
