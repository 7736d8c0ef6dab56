// TORCH_LIBRARY registration of the forward lookups over the C ABI (include/recemb_b200.h), so that
// modules containing them survive torch.jit.script / torch.jit.save / torch.jit.load -- the reference
// exports its embedding module that way (embedding_module_gen.py:186-196: ModelWrapper(model, mask_model)
// -> torch.jit.script -> jit.save) and loads it back in the LTHM encoder (models/lthm/sequence/
// encoder.py:25-29).  A ctypes call cannot be scripted; a registered operator can.
//
// Inference-side operators (the exported module is consumed detached, product_tower.py:47): no autograd
// formula is registered, calling them on tensors that require grad raises in the dispatcher's default
// fallback.  CPU tensors raise: there is no CPU fallback.  No torch type crosses into the kernels -- this
// file only unwraps pointers, sizes and the current stream and calls the plain C entry points.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "../../include/recemb_b200.h"

namespace {

int dtype_code(const at::Tensor& t) {
  if (t.scalar_type() == at::kFloat) return RECEMB_F32;
  if (t.scalar_type() == at::kBFloat16) return RECEMB_BF16;
  TORCH_CHECK(false, "recemb_b200: table dtype must be float32 or bfloat16, got ", t.scalar_type());
}

void check_call(int rc, const char* what) {
  TORCH_CHECK(rc == RECEMB_OK, what, " failed (rc=", rc, "): ", recemb_last_error());
}

void check_inputs(const at::Tensor& table, const at::Tensor& ids) {
  TORCH_CHECK(table.is_cuda() && ids.is_cuda(), "recemb_b200 runs on CUDA (sm_100a) only; there is no CPU fallback");
  TORCH_CHECK(table.dim() == 2 && table.is_contiguous(), "table must be a contiguous [rows, dim] tensor");
  TORCH_CHECK(ids.scalar_type() == at::kLong, "ids must be int64");  // models/lthm/sequence/wrapper.py:52
  TORCH_CHECK(table.get_device() == ids.get_device(), "table and ids on different devices");
}

// KShiftEmbedding.forward (commons/layers.py:152-172): one fused kernel
at::Tensor kshift_fwd(const at::Tensor& table, const at::Tensor& ids, int64_t num_shifts, bool normalize_output,
                      int64_t flip_len) {
  check_inputs(table, ids);
  c10::cuda::CUDAGuard guard(table.device());
  const at::Tensor flat = ids.contiguous();
  auto sizes = flat.sizes().vec();
  sizes.push_back(table.size(1));
  at::Tensor out = at::empty(sizes, table.options());
  auto stream = c10::cuda::getCurrentCUDAStream(table.get_device());
  check_call(recemb_kshift_fwd(table.data_ptr(), table.size(0), (int32_t)table.size(1), dtype_code(table),
                               flat.data_ptr<int64_t>(), flat.numel(), (int32_t)num_shifts,
                               normalize_output ? RECEMB_EPI_L2NORM : RECEMB_EPI_RSQRT_K, (int32_t)flip_len,
                               out.data_ptr(), nullptr, table.get_device(), stream.stream()),
             "recemb_kshift_fwd");
  return out;
}

// FlatEmbedding.forward (commons/layers.py:56-61) with the optional fused pad mask / flip
at::Tensor gather_fwd(const at::Tensor& table, const at::Tensor& ids, bool normalize_output, bool zero_pad,
                      int64_t flip_len) {
  check_inputs(table, ids);
  c10::cuda::CUDAGuard guard(table.device());
  const at::Tensor flat = ids.contiguous();
  auto sizes = flat.sizes().vec();
  sizes.push_back(table.size(1));
  at::Tensor out = at::empty(sizes, table.options());
  recemb_layout layout = {0, 0, 1, 0, (int32_t)flip_len};
  auto stream = c10::cuda::getCurrentCUDAStream(table.get_device());
  check_call(recemb_gather_fwd(table.data_ptr(), table.size(0), nullptr, 0, (int32_t)table.size(1), dtype_code(table),
                               flat.data_ptr<int64_t>(), flat.numel(), flip_len > 0 ? &layout : nullptr,
                               RECEMB_HASH_FLOORMOD, 0, 0, normalize_output ? RECEMB_EPI_L2NORM : RECEMB_EPI_NONE,
                               zero_pad ? 1 : 0, 0, out.data_ptr(), nullptr, table.get_device(), stream.stream()),
             "recemb_gather_fwd");
  return out;
}

// nn.EmbeddingBag(mode='sum' | 'mean') on [num_bags, bag_size] ids (commons/transformers/layers.py:457, :469)
at::Tensor pool_fwd(const at::Tensor& table, const at::Tensor& ids, bool hash_ids, bool mean) {
  check_inputs(table, ids);
  TORCH_CHECK(ids.dim() == 2, "pooled bags take ids of shape [num_bags, bag_size]");
  c10::cuda::CUDAGuard guard(table.device());
  const at::Tensor flat = ids.contiguous();
  at::Tensor out = at::empty({flat.size(0), table.size(1)}, table.options());
  auto stream = c10::cuda::getCurrentCUDAStream(table.get_device());
  check_call(recemb_pool_fwd(table.data_ptr(), table.size(0), (int32_t)table.size(1), dtype_code(table),
                             flat.data_ptr<int64_t>(), flat.size(0), (int32_t)flat.size(1), nullptr, 0, nullptr,
                             hash_ids ? RECEMB_HASH_FLOORMOD : RECEMB_HASH_IDENTITY, 0,
                             mean ? RECEMB_POOL_MEAN : RECEMB_POOL_SUM, 0, 0, nullptr, out.data_ptr(),
                             table.get_device(), stream.stream()),
             "recemb_pool_fwd");
  return out;
}

}  // namespace

TORCH_LIBRARY(recemb_b200, m) {
  m.def("kshift_fwd(Tensor table, Tensor ids, int num_shifts, bool normalize_output, int flip_len) -> Tensor");
  m.def("gather_fwd(Tensor table, Tensor ids, bool normalize_output, bool zero_pad, int flip_len) -> Tensor");
  m.def("pool_fwd(Tensor table, Tensor ids, bool hash_ids, bool mean) -> Tensor");
}

TORCH_LIBRARY_IMPL(recemb_b200, CUDA, m) {
  m.impl("kshift_fwd", kshift_fwd);
  m.impl("gather_fwd", gather_fwd);
  m.impl("pool_fwd", pool_fwd);
}
