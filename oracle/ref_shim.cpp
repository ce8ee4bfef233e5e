// Test infrastructure only (see oracle/build_ref.py).  Registers the reference's own arch-generic
// index kernels — compiled unmodified from /root/reference/csrc/indexed_io/{mask_to_indices,
// topk_indices,copy_indices,scatter_add}.cu — under a SEPARATE operator namespace, `torch.ops.chipmunk_ref.*`,
// so the GPU parity tests can run the real reference kernels next to ours on the same inputs.
// The schemas are the reference's (csrc/chipmunk.cpp:56-59); nothing in the product path loads this.
#include <torch/extension.h>
#include <ATen/ATen.h>
#include <vector>

namespace chipmunk {
extern void copy_indices(at::Tensor bmfc1, at::Tensor bm_mid_cache, at::Tensor sp_inds, at::Tensor sp_counts);
extern void topk_indices(at::Tensor activation, at::Tensor indices, at::Tensor counts, double sparsity_amount,
                         int64_t multiple_of, double random_amount);
extern std::vector<at::Tensor> mask_to_indices(at::Tensor mask, int64_t multiple_of, int64_t pad_to_multiple_of);
extern void csp_scatter_add(at::Tensor packed, at::Tensor unpacked_colmajor, at::Tensor sp_inds, at::Tensor sp_counts, int64_t num_sms);
}  // namespace chipmunk

TORCH_LIBRARY(chipmunk_ref, m) {
    m.def("copy_indices(Tensor bmfc1, Tensor(bm_mid_cache!) bm_mid_cache, Tensor sp_inds, Tensor sp_counts) -> ()");
    m.def("topk_indices(Tensor activation, Tensor(indices!) indices, Tensor(counts!) counts, float sparsity_amount, int multiple_of, float random_amount) -> ()");
    m.def("mask_to_indices(Tensor mask, int multiple_of, int pad_to_multiple_of) -> Tensor[]");
    m.def("csp_scatter_add(Tensor packed, Tensor(unpacked_colmajor!) unpacked_colmajor, Tensor sp_inds, Tensor sp_counts, int num_sms) -> ()");
}

TORCH_LIBRARY_IMPL(chipmunk_ref, CUDA, m) {
    m.impl("copy_indices", &chipmunk::copy_indices);
    m.impl("topk_indices", &chipmunk::topk_indices);
    m.impl("mask_to_indices", &chipmunk::mask_to_indices);
    m.impl("csp_scatter_add", &chipmunk::csp_scatter_add);
}
