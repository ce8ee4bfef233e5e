// The ten operator entry points of the reference's binding TU, on top of the chipmunk_b200 C ABI.
//
// The reference's csrc/chipmunk.cpp:27-43 DECLARES `chipmunk::csp_attn(at::Tensor ...)` and nine more as `extern`,
// registers them (`TORCH_LIBRARY(chipmunk, m)` :45-61, `TORCH_LIBRARY_IMPL(chipmunk, CUDA, m)` :64-80) and exposes the
// empty `chipmunk.cuda` module (:9-25); the definitions live in csrc/attn, csrc/mlp and csrc/indexed_io (ThunderKittens,
// sm_90a).  A maintainer who keeps that binding TU replaces those three directories by THIS file and links
// libchipmunk_b200.so:
//
//     g++ -shared -fPIC -std=c++17  csrc/chipmunk.cpp  bindings/chipmunk_ops_b200.cpp  -Iinclude
//         <torch include and library flags>  -Lchipmunk_b200 -lchipmunk_b200  -o chipmunk/cuda.so      (bindings/build_binding.py)
//
// `import chipmunk.cuda` then registers `torch.ops.chipmunk.*` exactly as before, and every call lands on the sm_100a
// kernels.  tests/test_cpp_binding_cpu.py compiles this file, and -- where /root/reference is mounted -- builds that
// extension from the reference's own, unmodified csrc/chipmunk.cpp and loads it in a fresh interpreter.  The shipped Python
// package registers the same schemas from Python instead (chipmunk_b200/torch_ops.py: ctypes -> the same C ABI); the
// checks and the pointer / size mapping below mirror that file function by function.
//
// Every kernel runs on the CURRENT stream of the tensors' device, nothing here synchronises.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include <vector>

#include "chipmunk_b200.h"

namespace chipmunk {
namespace {

constexpr int64_t QG = 192;      // query rows per index group
constexpr int64_t MLP_BM = 128;  // token rows per MLP index group

void check(int rc, const char* what) { TORCH_CHECK(rc == 0, what, ": ", cm_strerror(rc), " (code ", rc, ")"); }

void* stream_of(const at::Tensor& t) { return at::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

struct Strides3 {
    int64_t s[3];
    explicit Strides3(const at::Tensor& t) : s{t.stride(0), t.stride(1), t.stride(2)} {}
};

void check_bhnd(const at::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda(), "chipmunk_b200 kernels need CUDA tensors; there is no CPU path");
    TORCH_CHECK(t.dim() == 4 && t.scalar_type() == at::kBFloat16, name, " must be bfloat16 [B,H,N,D]");
    TORCH_CHECK(t.size(3) == 128, "Head dimension must be 128");
    TORCH_CHECK(t.stride(3) == 1, name, ".stride(3) must be 1");
    TORCH_CHECK(t.stride(0) % 8 == 0 && t.stride(1) % 8 == 0 && t.stride(2) % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(t.data_ptr()) % 16 == 0,
                name, " must be 16-byte aligned in every stride");
}

// shared by csp_attn / csp_128_attn: the reference's TORCH_CHECKs (csrc/attn/csp_attn.cu:325-364)
void check_attn(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& indices,
                const at::Tensor& counts) {
    check_bhnd(q, "q"); check_bhnd(k, "k"); check_bhnd(v, "v");
    TORCH_CHECK(k.size(0) == q.size(0) && v.size(0) == q.size(0), "K/V batch dimension - idx 0 - must match for all inputs");
    TORCH_CHECK(k.size(1) == q.size(1) && v.size(1) == q.size(1), "QO heads must be equal to KV heads");
    TORCH_CHECK(k.size(2) == v.size(2), "V sequence length dimension - idx 2 - must match for all inputs");
    const int64_t G = (q.size(2) + QG - 1) / QG;
    TORCH_CHECK(indices.is_cuda() && counts.is_cuda(), "indices / counts must be CUDA tensors");
    TORCH_CHECK(indices.dim() == 4, "Indices must be a 4D tensor");
    TORCH_CHECK(counts.dim() == 3, "Indices counts must be a 3D tensor");
    TORCH_CHECK(indices.scalar_type() == at::kInt, "Indices must be a 32-bit integer tensor");
    TORCH_CHECK(counts.scalar_type() == at::kInt, "Indices counts must be a 32-bit integer tensor");
    TORCH_CHECK(indices.is_contiguous() && counts.is_contiguous(), "Indices and counts must be contiguous");
    TORCH_CHECK(indices.size(0) == q.size(0) && indices.size(1) == q.size(1) && indices.size(2) == G,
                "Indices [batch, head, query group] dimensions must match q");
    TORCH_CHECK(counts.size(0) == q.size(0) && counts.size(1) == q.size(1) && counts.size(2) == G,
                "Indices counts [batch, head, query group] dimensions must match q");
}

void launch_csp_attn(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, at::Tensor& o,
                     const at::Tensor& indices, const at::Tensor& counts, int o_scale, int accumulate) {
    check_attn(q, k, v, indices, counts);
    check_bhnd(o, "o");
    TORCH_CHECK(o.sizes() == q.sizes(), "O must match Q");
    if (q.numel() == 0) return;
    const c10::cuda::CUDAGuard guard(q.device());
    const Strides3 qs(q), ks(k), vs(v), os(o);
    check(cm_csp_attn(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), indices.data_ptr<int32_t>(),
                      counts.data_ptr<int32_t>(), (int)q.size(0), (int)q.size(1), (int)q.size(2), (int)k.size(2), qs.s,
                      ks.s, vs.s, os.s, indices.size(3), o_scale, accumulate, stream_of(q)),
          "csp_attn");
}

std::vector<at::Tensor> launch_dense(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor* p) {
    check_bhnd(q, "q"); check_bhnd(k, "k"); check_bhnd(v, "v");
    TORCH_CHECK(k.sizes() == v.sizes() && k.size(0) == q.size(0) && k.size(1) == q.size(1),
                "dense_attn: K/V shapes must match Q's batch and heads");
    const int64_t B = q.size(0), H = q.size(1), Nq = q.size(2), Nk = k.size(2), G = (Nq + QG - 1) / QG;
    at::Tensor o = at::empty(q.sizes(), q.options());
    at::Tensor l = at::empty({B, H, Nq, 1}, q.options().dtype(at::kFloat));
    at::Tensor cs, pc;
    const int64_t cs_stride = (Nk + 7) / 8 * 8;
    if (p) {
        TORCH_CHECK(p->is_cuda() && p->scalar_type() == at::kFloat && p->numel() >= B * H * Nq && p->size(0) == B &&
                        p->size(1) == H,
                    "dense_colsum_attn: p must be fp32 [B,H,N,1]");
        pc = p->reshape({B, H, -1}).slice(2, 0, Nq).contiguous();
        cs = at::empty({B, H, G, cs_stride}, q.options());
    }
    if (q.numel() != 0) {
        const c10::cuda::CUDAGuard guard(q.device());
        const Strides3 qs(q), ks(k), vs(v), os(o);
        check(cm_dense_attn_strided(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), l.data_ptr<float>(),
                                    p ? cs.data_ptr() : nullptr, p ? pc.data_ptr<float>() : nullptr, (int)B, (int)H,
                                    (int)Nq, (int)Nk, qs.s, ks.s, vs.s, os.s, cs_stride, stream_of(q)),
              "dense_attn");
    }
    if (p) return {o, cs_stride == Nk ? cs : cs.slice(3, 0, Nk), l};
    return {o, l};
}

void check_mlp_indices(const at::Tensor& indices, const at::Tensor& counts, int64_t M, const char* what) {
    TORCH_CHECK(indices.is_cuda() && counts.is_cuda(), what, ": indices / counts must be CUDA tensors");
    TORCH_CHECK(indices.scalar_type() == at::kInt && counts.scalar_type() == at::kInt, what, ": indices/counts must be int32");
    TORCH_CHECK(indices.is_contiguous() && counts.is_contiguous(), what, ": indices/counts must be contiguous");
    TORCH_CHECK(M % MLP_BM == 0, what, ": M must be a multiple of 128");
    TORCH_CHECK(indices.numel() == (M / MLP_BM) * indices.size(-1) && counts.numel() == M / MLP_BM, what,
                ": indices must be [M/128, F] and counts [M/128]");
}

void check_bf16_contiguous(const at::Tensor& t, const char* what) {
    TORCH_CHECK(t.is_cuda(), "chipmunk_b200 kernels need CUDA tensors; there is no CPU path");
    TORCH_CHECK(t.scalar_type() == at::kBFloat16 && t.is_contiguous(), what, ": tensors must be contiguous bfloat16");
}

int dtype_tag(at::ScalarType t) {
    switch (t) {
        case at::kBFloat16: return CM_BF16;
        case at::kHalf: return CM_F16;
        case at::kFloat: return CM_F32;
        default: TORCH_CHECK(false, "Unsupported dtype for activation tensor");
    }
    return -1;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ attention
void csp_attn(at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor o, at::Tensor indices, at::Tensor indices_counts,
              int64_t o_scale) {
    TORCH_CHECK(o_scale == 1 || o_scale == -1, "o_scale must be 1 or -1");
    launch_csp_attn(q, k, v, o, indices, indices_counts, (int)o_scale, /*accumulate=*/1);
}

at::Tensor csp_128_attn(at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor indices, at::Tensor indices_counts) {
    at::Tensor o = at::empty(q.sizes(), q.options());
    launch_csp_attn(q, k, v, o, indices, indices_counts, 1, /*accumulate=*/0);
    return o;
}

std::vector<at::Tensor> dense_attn(at::Tensor q, at::Tensor k, at::Tensor v) { return launch_dense(q, k, v, nullptr); }

std::vector<at::Tensor> dense_colsum_attn(at::Tensor q, at::Tensor k, at::Tensor v, at::Tensor p) {
    return launch_dense(q, k, v, &p);
}

// ------------------------------------------------------------------------------------------------------ MLP
void csp_mlp_mm1(at::Tensor a, at::Tensor b_colmajor, at::Tensor c, at::Tensor bias, at::Tensor pa_cache_colmajor,
                 at::Tensor indices, at::Tensor indices_counts) {
    for (const at::Tensor* t : {&a, &b_colmajor, &c, &bias, &pa_cache_colmajor}) check_bf16_contiguous(*t, "csp_mlp_mm1");
    TORCH_CHECK(a.dim() == 2 && b_colmajor.dim() == 2 && c.dim() == 2, "csp_mlp_mm1: a [M,K], b_colmajor [F,K], c [M,F]");
    const int64_t M = a.size(0), K = a.size(1), F = b_colmajor.size(0);
    TORCH_CHECK(b_colmajor.size(1) == K, "csp_mlp_mm1: K must match");
    TORCH_CHECK(c.size(0) == M && c.size(1) == F, "csp_mlp_mm1: c must be [M,F]");
    TORCH_CHECK(bias.numel() == F, "csp_mlp_mm1: bias must be [F]");
    TORCH_CHECK(pa_cache_colmajor.dim() == 2 && pa_cache_colmajor.size(0) == F && pa_cache_colmajor.size(1) == M,
                "csp_mlp_mm1: pa_cache_colmajor must be [F,M]");
    TORCH_CHECK(K % 64 == 0, "csp_mlp_mm1: K must be a multiple of 64");
    check_mlp_indices(indices, indices_counts, M, "csp_mlp_mm1");
    const c10::cuda::CUDAGuard guard(a.device());
    check(cm_csp_mlp_mm1(a.data_ptr(), b_colmajor.data_ptr(), c.data_ptr(), bias.data_ptr(), pa_cache_colmajor.data_ptr(),
                         indices.data_ptr<int32_t>(), indices_counts.data_ptr<int32_t>(), (int)M, (int)K, (int)F,
                         indices.size(-1), /*update_pa=*/0, stream_of(a)),
          "csp_mlp_mm1");
}

// `matmul_kernel` was the raw CUfunction of the reference's Triton mm2 and `num_sms_scatter_add` its SM split
// (csrc/mlp/csp_mlp_mm2_and_scatter_add.cu:181-256); one sm_100a kernel does the GEMM and the scatter, both are ignored.
void csp_mlp_mm2_and_scatter_add(at::Tensor packed, at::Tensor unpacked_colmajor, at::Tensor sp_inds, at::Tensor sp_counts,
                                 at::Tensor mma_a, at::Tensor mma_b, at::Tensor mma_c, int64_t /*num_sms_scatter_add*/,
                                 int64_t /*matmul_kernel*/) {
    TORCH_CHECK(packed.dim() == 3 && packed.size(0) == 1, "csp_mlp_mm2_and_scatter_add: batch must be 1");
    TORCH_CHECK(mma_a.data_ptr() == packed.data_ptr(), "csp_mlp_mm2_and_scatter_add: mma_a must alias packed");
    for (const at::Tensor* t : {&packed, &unpacked_colmajor, &mma_b, &mma_c}) check_bf16_contiguous(*t, "csp_mlp_mm2");
    const int64_t M = packed.size(-2), F = packed.size(-1), N = mma_b.size(-1);
    TORCH_CHECK(mma_b.size(-2) == F && mma_c.size(-2) == M && mma_c.size(-1) == N,
                "csp_mlp_mm2: shapes must be packed [M,F], w2_T [F,N], out [M,N]");
    TORCH_CHECK(N % 256 == 0, "csp_mlp_mm2: N must be a multiple of 256");
    TORCH_CHECK(unpacked_colmajor.size(-2) == F && unpacked_colmajor.size(-1) == M,
                "csp_mlp_mm2: unpacked_colmajor must be contiguous bf16 [F,M]");
    check_mlp_indices(sp_inds, sp_counts, M, "csp_mlp_mm2");
    const c10::cuda::CUDAGuard guard(packed.device());
    check(cm_csp_mlp_mm2(packed.data_ptr(), mma_b.data_ptr(), mma_c.data_ptr(), unpacked_colmajor.data_ptr(),
                         sp_inds.data_ptr<int32_t>(), sp_counts.data_ptr<int32_t>(), (int)M, (int)F, (int)N,
                         sp_inds.size(-1), /*do_scatter=*/1, stream_of(packed)),
          "csp_mlp_mm2");
}

void csp_scatter_add(at::Tensor packed, at::Tensor unpacked_colmajor, at::Tensor sp_inds, at::Tensor sp_counts,
                     int64_t /*num_sms*/) {
    TORCH_CHECK(packed.dim() == 3 && packed.size(0) == 1, "csp_scatter_add: batch must be 1");
    check_bf16_contiguous(packed, "csp_scatter_add"); check_bf16_contiguous(unpacked_colmajor, "csp_scatter_add");
    const int64_t M = packed.size(-2), F = packed.size(-1);
    TORCH_CHECK(unpacked_colmajor.size(-2) == F && unpacked_colmajor.size(-1) == M,
                "csp_scatter_add: unpacked_colmajor must be [1,F,M]");
    check_mlp_indices(sp_inds, sp_counts, M, "csp_scatter_add");
    const c10::cuda::CUDAGuard guard(packed.device());
    check(cm_csp_scatter_add(packed.data_ptr(), unpacked_colmajor.data_ptr(), sp_inds.data_ptr<int32_t>(),
                             sp_counts.data_ptr<int32_t>(), (int)M, (int)F, sp_inds.size(-1), stream_of(packed)),
          "csp_scatter_add");
}

// ----------------------------------------------------------------------------------------------- indexed IO
void copy_indices(at::Tensor bmfc1, at::Tensor bm_mid_cache, at::Tensor sp_inds, at::Tensor sp_counts) {
    TORCH_CHECK(bmfc1.is_cuda() && bm_mid_cache.is_cuda() && sp_inds.is_cuda() && sp_counts.is_cuda(),
                "chipmunk_b200 kernels need CUDA tensors; there is no CPU path");
    TORCH_CHECK(sp_inds.scalar_type() == at::kInt, "sp_inds must be int32");
    TORCH_CHECK(sp_counts.scalar_type() == at::kInt, "sp_counts must be int32");
    TORCH_CHECK(bmfc1.scalar_type() == bm_mid_cache.scalar_type() && bmfc1.sizes() == bm_mid_cache.sizes(),
                "copy_indices: src/dst must match");
    TORCH_CHECK(bmfc1.scalar_type() == at::kBFloat16 || bmfc1.scalar_type() == at::kHalf || bmfc1.scalar_type() == at::kFloat,
                "Unsupported tensor type");
    TORCH_CHECK(bmfc1.is_contiguous() && bm_mid_cache.is_contiguous() && sp_inds.is_contiguous() && sp_counts.is_contiguous(),
                "copy_indices: tensors must be contiguous");
    TORCH_CHECK(bmfc1.dim() == 3 && sp_inds.dim() == 3, "copy_indices: shapes must be [B,M*R,F] / [B,M,F]");
    const int64_t B = bmfc1.size(0), M = sp_inds.size(1), F = sp_inds.size(2);
    TORCH_CHECK(bmfc1.size(2) == F && bm_mid_cache.size(1) % M == 0, "copy_indices: shapes must be [B,M*R,F] / [B,M,F]");
    const c10::cuda::CUDAGuard guard(bmfc1.device());
    check(cm_copy_indices(bmfc1.data_ptr(), bm_mid_cache.data_ptr(), (int)bmfc1.element_size(), sp_inds.data_ptr<int32_t>(),
                          sp_counts.data_ptr<int32_t>(), (int)B, (int)M, (int)(bm_mid_cache.size(1) / M), (int)F,
                          stream_of(bmfc1)),
          "copy_indices");
}

void topk_indices(at::Tensor activation, at::Tensor indices, at::Tensor counts, double sparsity_amount, int64_t multiple_of,
                  double random_amount) {
    TORCH_CHECK(activation.is_cuda() && indices.is_cuda() && counts.is_cuda(),
                "chipmunk_b200 kernels need CUDA tensors; there is no CPU path");
    TORCH_CHECK(activation.dim() == 3, "activation must be 3-dimensional [batch, rows, cols]");
    TORCH_CHECK(indices.dim() == 3, "indices must be 3-dimensional [batch, rows, cols]");
    TORCH_CHECK(counts.dim() == 2, "counts must be 2-dimensional [batch, rows]");
    TORCH_CHECK(sparsity_amount >= 0 && sparsity_amount <= 1, "sparsity_amount must be between 0 and 1");
    TORCH_CHECK(indices.scalar_type() == at::kInt && counts.scalar_type() == at::kInt, "indices/counts must be int32");
    TORCH_CHECK(activation.is_contiguous() && indices.is_contiguous() && counts.is_contiguous(),
                "topk_indices: tensors must be contiguous");
    TORCH_CHECK(indices.sizes() == activation.sizes() && counts.size(0) == activation.size(0) &&
                    counts.size(1) == activation.size(1),
                "topk_indices: indices/counts shapes must match activation");
    const c10::cuda::CUDAGuard guard(activation.device());
    check(cm_topk_indices(activation.data_ptr(), dtype_tag(activation.scalar_type()), indices.data_ptr<int32_t>(),
                          counts.data_ptr<int32_t>(), (int)activation.size(0), (int)activation.size(1),
                          (int)activation.size(2), (float)sparsity_amount, (int)multiple_of, (float)random_amount,
                          stream_of(activation)),
          "topk_indices");
}

std::vector<at::Tensor> mask_to_indices(at::Tensor mask, int64_t multiple_of, int64_t pad_to_multiple_of) {
    TORCH_CHECK(mask.is_cuda(), "chipmunk_b200 kernels need CUDA tensors; there is no CPU path");
    TORCH_CHECK(mask.dim() == 4, "mask must be 4-dimensional [b, h, m, n]");
    TORCH_CHECK(mask.scalar_type() == at::kBool, "mask must be bool type");
    mask = mask.contiguous();
    const int64_t b = mask.size(0), h = mask.size(1), m = mask.size(2), n = mask.size(3);
    const int64_t pad_n = (n + pad_to_multiple_of - 1) / pad_to_multiple_of * pad_to_multiple_of;
    at::Tensor indices = at::empty({b, h, m, pad_n}, mask.options().dtype(at::kInt));
    at::Tensor counts = at::empty({b, h, m}, mask.options().dtype(at::kInt));
    const c10::cuda::CUDAGuard guard(mask.device());
    check(cm_mask_to_indices(reinterpret_cast<const uint8_t*>(mask.data_ptr<bool>()), indices.data_ptr<int32_t>(),
                             counts.data_ptr<int32_t>(), b * h * m, (int)n, (int)pad_n, (int)multiple_of, stream_of(mask)),
          "mask_to_indices");
    return {indices, counts};
}

}  // namespace chipmunk
