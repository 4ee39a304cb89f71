// Thin torch C++ extension over the C ABI (include/osd_b200.h): the pybind11 module `oneshotdet_b200._C_torch`, the
// stand-in for the reference's `maskrcnn_benchmark._C` (csrc/vision.cpp:7-15) on this path.  It exports
//   * `nms` with exactly the signature and return contract of `_C.nms` (csrc/vision.cpp:8, csrc/nms.h:10-28);
//   * `match_forward(features, supp_pooled, batch_size, mode)` -- the matching module's forward
//     (modeling/detector/generalized_rcnn.py:100-104, 306-311; concat: box_head.py:127,144,147);
//   * `fcos_postprocess(box_cls, box_regression, centerness, image_sizes, strides, ...)` -- the tensor-in / tensor-out
//     core of FCOSPostProcessor.forward (modeling/rpn/fcos/inference.py:251-323): padded [B,K,4] / [B,K] / [B,K] / [B].
// No kernels live here -- every call forwards to libosd_b200.so on the current CUDA stream; outputs and workspaces come
// from PyTorch's caching allocator; errors surface as RuntimeError (TORCH_CHECK), as the reference's AT_ASSERTM do.
#include <torch/extension.h>

#include <cstring>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

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

std::vector<at::Tensor> match_forward(const std::vector<at::Tensor>& features, const std::vector<at::Tensor>& supp_pooled,
                                      int64_t batch_size, const std::string& mode) {
  const size_t nl = features.size();
  TORCH_CHECK(nl >= 1 && nl <= OSD_MAX_LEVELS && supp_pooled.size() == nl, "match_forward: need 1..8 levels and one support tensor per level");
  osd_match_desc d;
  memset(&d, 0, sizeof(d));
  if (mode == "product") d.mode = OSD_MATCH_PRODUCT;
  else if (mode == "concat") d.mode = OSD_MATCH_CONCAT;
  else if (mode == "concat_reversed") d.mode = OSD_MATCH_CONCAT_REVERSED;
  else TORCH_CHECK(false, "match_forward: unknown mode '", mode, "' (product, concat, concat_reversed; the fusion mode owns parameters: use MatchingModule)");
  const auto& f0 = features[0];
  TORCH_CHECK(f0.is_cuda(), "match_forward: features must be CUDA tensors; this package has no CPU path");
  TORCH_CHECK(f0.dim() == 4 && f0.size(0) == batch_size, "match_forward: features must be [B,C,H,W] with B = batch_size");
  TORCH_CHECK(f0.scalar_type() == at::kFloat || f0.scalar_type() == at::kBFloat16, "match_forward: float32 or bfloat16");
  c10::cuda::CUDAGuard guard(f0.device());
  TORCH_CHECK(osd_check_device() == OSD_OK, osd_last_error());
  const bool nhwc = f0.is_contiguous(at::MemoryFormat::ChannelsLast) && !f0.is_contiguous();
  const int64_t B = batch_size, C = f0.size(1);
  d.num_levels = (int32_t)nl; d.batch = (int32_t)B; d.channels = (int32_t)C;
  d.layout = nhwc ? OSD_LAYOUT_NHWC : OSD_LAYOUT_NCHW;
  d.dtype = f0.scalar_type() == at::kFloat ? OSD_DTYPE_F32 : OSD_DTYPE_BF16;
  std::vector<at::Tensor> keep, outs;
  int64_t shots = -1;
  for (size_t l = 0; l < nl; ++l) {
    TORCH_CHECK(features[l].device() == f0.device() && supp_pooled[l].device() == f0.device(), "match_forward: all tensors on one device");
    TORCH_CHECK(features[l].scalar_type() == f0.scalar_type() && supp_pooled[l].scalar_type() == f0.scalar_type(), "match_forward: one dtype");
    TORCH_CHECK(features[l].dim() == 4 && features[l].size(0) == B && features[l].size(1) == C, "match_forward: level ", l, " shape");
    auto f = nhwc ? features[l].contiguous(at::MemoryFormat::ChannelsLast) : features[l].contiguous();
    TORCH_CHECK(supp_pooled[l].numel() > 0 && supp_pooled[l].numel() % (B * C) == 0, "match_forward: support must be [B*S,C,1,1]");
    const int64_t sl = supp_pooled[l].numel() / (B * C);
    if (shots < 0) shots = sl;
    TORCH_CHECK(sl == shots, "match_forward: every level must carry the same number of shots");
    auto s = supp_pooled[l].reshape({B * sl, C}).contiguous();
    const int64_t H = f.size(2), W = f.size(3);
    const int64_t cout = d.mode == OSD_MATCH_PRODUCT ? C : 2 * C;
    auto o = nhwc ? at::empty({B, cout, H, W}, f.options().memory_format(at::MemoryFormat::ChannelsLast))
                  : at::empty({B, cout, H, W}, f.options());
    d.hw[l] = (int32_t)(H * W);
    d.feat[l] = f.data_ptr(); d.supp[l] = s.data_ptr(); d.out[l] = o.data_ptr();
    keep.push_back(f); keep.push_back(s);
    outs.push_back(o);
  }
  d.shots = (int32_t)shots;
  TORCH_CHECK(osd_match_forward(&d, at::cuda::getCurrentCUDAStream().stream()) == OSD_OK, osd_last_error());
  return outs;
}

// -> (boxes [B,K,4], scores [B,K], index int32 [B,K], count int32 [B]); rows >= count[b] are unspecified
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor> fcos_postprocess(
    const std::vector<at::Tensor>& box_cls, const std::vector<at::Tensor>& box_regression, const std::vector<at::Tensor>& centerness,
    const std::vector<std::pair<int64_t, int64_t>>& image_sizes, const std::vector<int64_t>& strides, double pre_nms_thresh,
    int64_t pre_nms_top_n, double nms_thresh, int64_t fpn_post_nms_top_n, double min_size, bool strict, bool early_exit) {
  const size_t nl = box_cls.size();
  TORCH_CHECK(nl >= 1 && nl <= OSD_MAX_LEVELS && box_regression.size() == nl && centerness.size() == nl && strides.size() == nl,
              "fcos_postprocess: need 1..8 levels with cls / regression / centerness / stride each");
  const auto& c0 = box_cls[0];
  TORCH_CHECK(c0.is_cuda(), "fcos_postprocess: inputs must be CUDA tensors; this package has no CPU path");
  c10::cuda::CUDAGuard guard(c0.device());
  TORCH_CHECK(osd_check_device() == OSD_OK, osd_last_error());
  const int64_t B = c0.size(0);
  TORCH_CHECK((int64_t)image_sizes.size() == B, "fcos_postprocess: one (h, w) per episode");
  osd_fcos_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.num_levels = (int32_t)nl; cfg.batch = (int32_t)B;
  cfg.pre_nms_thresh = (float)pre_nms_thresh; cfg.pre_nms_top_n = (int32_t)pre_nms_top_n; cfg.nms_thresh = (float)nms_thresh;
  cfg.post_nms_top_n = (int32_t)fpn_post_nms_top_n; cfg.min_size = (float)min_size; cfg.strict = strict ? 1 : 0;
  cfg.early_exit = early_exit ? 1 : 0; cfg.reg_transform = OSD_REG_DISTANCES;
  std::vector<at::Tensor> keep;
  const float* pc[OSD_MAX_LEVELS]; const float* pr[OSD_MAX_LEVELS]; const float* pt[OSD_MAX_LEVELS];
  for (size_t l = 0; l < nl; ++l) {
    const auto& c = box_cls[l]; const auto& r = box_regression[l]; const auto& t = centerness[l];
    TORCH_CHECK(c.scalar_type() == at::kFloat && r.scalar_type() == at::kFloat && t.scalar_type() == at::kFloat, "fcos_postprocess: float32 inputs");
    TORCH_CHECK(c.dim() == 4 && c.size(0) == B && c.size(1) == 1, "fcos_postprocess: box_cls[", l, "] must be [B,1,H,W] (one foreground class)");
    const int64_t H = c.size(2), W = c.size(3);
    TORCH_CHECK(r.dim() == 4 && r.size(0) == B && r.size(1) == 4 && r.size(2) == H && r.size(3) == W, "fcos_postprocess: box_regression[", l, "] must be [B,4,H,W]");
    TORCH_CHECK(t.dim() == 4 && t.size(0) == B && t.size(1) == 1 && t.size(2) == H && t.size(3) == W, "fcos_postprocess: centerness[", l, "] must be [B,1,H,W]");
    auto cc = c.contiguous(), rc = r.contiguous(), tc = t.contiguous();
    cfg.height[l] = (int32_t)H; cfg.width[l] = (int32_t)W; cfg.stride[l] = (int32_t)strides[l];
    pc[l] = cc.data_ptr<float>(); pr[l] = rc.data_ptr<float>(); pt[l] = tc.data_ptr<float>();
    keep.push_back(cc); keep.push_back(rc); keep.push_back(tc);
  }
  osd_fcos_plan plan;
  TORCH_CHECK(osd_fcos_postprocess_plan(&cfg, &plan) == OSD_OK, osd_last_error());
  auto hw_host = at::empty({B, 2}, at::TensorOptions().dtype(at::kInt).pinned_memory(true));
  for (int64_t b = 0; b < B; ++b) {
    hw_host[b][0] = (int32_t)image_sizes[b].first;
    hw_host[b][1] = (int32_t)image_sizes[b].second;
  }
  auto hw = hw_host.to(c0.device(), /*non_blocking=*/true);
  auto opt = c0.options();
  auto ws = at::empty({(int64_t)std::max<size_t>(plan.workspace_bytes, 256)}, opt.dtype(at::kByte));
  const int64_t K = plan.out_capacity;
  auto boxes = at::empty({B, K, 4}, opt);
  auto scores = at::empty({B, K}, opt);
  auto index = at::empty({B, K}, opt.dtype(at::kInt));
  auto count = at::zeros({B}, opt.dtype(at::kInt));
  const int rc = osd_fcos_postprocess(&cfg, pc, pr, pt, hw.data_ptr<int32_t>(), ws.data_ptr(), (size_t)ws.numel(),
                                      boxes.data_ptr<float>(), scores.data_ptr<float>(), index.data_ptr<int32_t>(),
                                      count.data_ptr<int32_t>(), at::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == OSD_OK, osd_last_error());
  return {boxes, scores, index, count};
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("nms", &nms, "non-maximum suppression (libosd_b200, sm_100a)");
  m.def("match_forward", &match_forward, "support -> target matching on the FPN levels (product / concat)",
        pybind11::arg("features"), pybind11::arg("supp_pooled"), pybind11::arg("batch_size"), pybind11::arg("mode") = "product");
  m.def("fcos_postprocess", &fcos_postprocess, "FCOSPostProcessor.forward on tensors: score, top-k, decode, NMS, post-NMS top-n",
        pybind11::arg("box_cls"), pybind11::arg("box_regression"), pybind11::arg("centerness"), pybind11::arg("image_sizes"),
        pybind11::arg("strides"), pybind11::arg("pre_nms_thresh"), pybind11::arg("pre_nms_top_n"), pybind11::arg("nms_thresh"),
        pybind11::arg("fpn_post_nms_top_n"), pybind11::arg("min_size"), pybind11::arg("strict") = false,
        pybind11::arg("early_exit") = true);
  m.def("version", []() { return osd_version(); });
}
