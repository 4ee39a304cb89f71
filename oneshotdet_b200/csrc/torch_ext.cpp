// Thin torch C++ extension over the C ABI (include/osd_b200.h): the pybind11 module `oneshotdet_b200._C_torch`
// exports `nms` with exactly the signature and return contract of the reference's
// `maskrcnn_benchmark._C.nms` (csrc/vision.cpp:8, csrc/nms.h:10-28).  No kernels live here -- every call forwards to
// libosd_b200.so on the current CUDA stream; tensors come from PyTorch's caching allocator.
#include <torch/extension.h>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "osd_b200.h"

namespace {

at::Tensor nms(const at::Tensor& dets, const at::Tensor& scores, const double threshold) {
  TORCH_CHECK(dets.is_cuda(), "oneshotdet_b200._C_torch.nms: dets must be a CUDA tensor; this package has no CPU path");
  TORCH_CHECK(scores.is_cuda(), "oneshotdet_b200._C_torch.nms: scores must be a CUDA tensor");
  TORCH_CHECK(dets.scalar_type() == scores.scalar_type(), "dets should have the same type as scores");  // nms_cpu.cpp:11
  if (dets.numel() == 0)  // csrc/nms.h:17-18: empty int64 tensor on the CPU
    return at::empty({0}, dets.options().dtype(at::kLong).device(at::kCPU));
  TORCH_CHECK(dets.scalar_type() == at::kFloat, "oneshotdet_b200._C_torch.nms: only float32 is supported (csrc/cuda/nms.cu:71)");
  TORCH_CHECK(dets.dim() == 2 && dets.size(1) == 4 && scores.dim() == 1 && scores.size(0) == dets.size(0),
              "expected dets [N,4] and scores [N]");
  c10::cuda::CUDAGuard guard(dets.device());
  TORCH_CHECK(osd_check_device() == OSD_OK, osd_last_error());
  const auto d = dets.contiguous();
  const auto s = scores.contiguous();
  const int64_t n = d.size(0);
  osd_nms_plan plan;
  TORCH_CHECK(osd_batched_nms_plan(1, n, &plan) == OSD_OK, osd_last_error());
  auto ws = at::empty({(int64_t)plan.workspace_bytes}, d.options().dtype(at::kByte));
  auto seg = at::empty({2}, d.options().dtype(at::kLong));
  seg.select(0, 0).fill_(0);
  seg.select(0, 1).fill_(n);
  auto keep = at::empty({n}, d.options().dtype(at::kLong));
  auto cnt = at::zeros({1}, d.options().dtype(at::kInt));
  const int rc = osd_batched_nms(d.data_ptr<float>(), s.data_ptr<float>(), seg.data_ptr<int64_t>(), 1, n, (float)threshold,
                                 /*strict=*/0, ws.data_ptr(), plan.workspace_bytes, keep.data_ptr<int64_t>(),
                                 cnt.data_ptr<int32_t>(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == OSD_OK, osd_last_error());
  return keep.narrow(0, 0, cnt.item<int32_t>());
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("nms", &nms, "non-maximum suppression (libosd_b200, sm_100a)");
  m.def("version", []() { return osd_version(); });
}
