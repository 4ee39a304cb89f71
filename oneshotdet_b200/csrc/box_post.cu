// Second-stage (ROI box head) post-processing: class probability, BoxCoder.decode, clip, score filter, per-image NMS,
// detections-per-image cut -- one call for all images of the batch.
//
// Reference: maskrcnn_benchmark/modeling/roi_heads/box_head/inference.py:46-167 (PostProcessor.forward /
// prepare_boxlist / filter_results), modeling/box_coder.py:52-95 (BoxCoder.decode),
// structures/bounding_box.py:214-224 (clip_to_image).  SURVEY section 8(f) row 2 (post-processing half).
//
// The episode problem has ONE foreground class (num_classes = 2, inference.py:89): only class 1's probability and
// class 1's box (regression columns [4, 8)) survive filter_results (:143-159), so the kernel decodes exactly those.
// Candidates are compacted per image in proposal order (what `inds_all[:, 1].nonzero()` yields, :144) and handed to
// the same NMS pipeline as the FCOS stage (nms.cu).
#include "osd_common.cuh"
#include "osd_device_utils.cuh"

namespace osd {
namespace {

constexpr int kDecThreads = 256;

struct DecodeArgs {
  const float* logits;   // [B*R, nlog]
  const float* reg;      // [B*R, regc]
  const float4* props;   // [B, R] xyxy
  const int32_t* roi_count;  // [B] or null (= R)
  const int32_t* image_hw;   // [B,2] (h, w)
  int R, nlog, regc, reg_off, score_mode;
  float wx, wy, ww, wh, clip, score_thresh;
  float4* cand_boxes;    // [B, R]
  float* cand_scores;    // [B, R]
  int32_t* cand_src;     // [B, R] proposal row of each candidate
  int32_t* cand_count;   // [B]
};

// inference.py:62-70: class-1 probability
__device__ __forceinline__ float class1_prob(const float* lg, int nlog, int mode) {
  if (mode == OSD_SCORE_SIGMOID) {  // focal_loss (:62-65) -- and mse/l1 with a single logit (:68-70)
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-lg[0])));
  }
  // ce / cxe: F.softmax(class_logits, -1)[:, 1]  (:66-67); ATen: exp(x - max), sum, multiply by 1/sum
  float m = lg[0];
  for (int j = 1; j < nlog; ++j) m = fmaxf(m, lg[j]);
  float sum = 0.f, e1 = 0.f;
  for (int j = 0; j < nlog; ++j) {
    const float e = expf(__fsub_rn(lg[j], m));
    sum = __fadd_rn(sum, e);
    if (j == 1) e1 = e;
  }
  return __fmul_rn(e1, __fdiv_rn(1.0f, sum));
}

// box_coder.py:62-95, then bounding_box.py:214-224
__device__ __forceinline__ float4 decode_box(const float4 p, const float* r, const DecodeArgs& A, float xmax, float ymax) {
  const float w = __fadd_rn(__fsub_rn(p.z, p.x), 1.0f);
  const float h = __fadd_rn(__fsub_rn(p.w, p.y), 1.0f);
  const float cx = __fadd_rn(p.x, __fmul_rn(0.5f, w));
  const float cy = __fadd_rn(p.y, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(r[0], A.wx);
  const float dy = __fdiv_rn(r[1], A.wy);
  const float dw = fminf(__fdiv_rn(r[2], A.ww), A.clip);
  const float dh = fminf(__fdiv_rn(r[3], A.wh), A.clip);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx);
  const float pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float hw_ = __fmul_rn(0.5f, __fmul_rn(expf(dw), w));
  const float hh_ = __fmul_rn(0.5f, __fmul_rn(expf(dh), h));
  float4 o;
  o.x = __fsub_rn(pcx, hw_);
  o.y = __fsub_rn(pcy, hh_);
  o.z = __fsub_rn(__fadd_rn(pcx, hw_), 1.0f);
  o.w = __fsub_rn(__fadd_rn(pcy, hh_), 1.0f);
  o.x = fminf(fmaxf(o.x, 0.f), xmax);
  o.y = fminf(fmaxf(o.y, 0.f), ymax);
  o.z = fminf(fmaxf(o.z, 0.f), xmax);
  o.w = fminf(fmaxf(o.w, 0.f), ymax);
  return o;
}

// One CTA per image; rows are visited in chunks of blockDim.x so that the compaction keeps proposal order.
__global__ void __launch_bounds__(kDecThreads) box_decode_kernel(DecodeArgs A) {
  __shared__ int warp_tot[33];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = A.roi_count ? min(max(A.roi_count[b], 0), A.R) : A.R;
  const float xmax = (float)(A.image_hw[2 * b + 1] - 1), ymax = (float)(A.image_hw[2 * b] - 1);
  const size_t row0 = (size_t)b * A.R;
  int running = 0;
  for (int base = 0; base < n; base += kDecThreads) {
    const int i = base + tid;
    bool ok = false;
    float score = 0.f;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      score = class1_prob(A.logits + (row0 + i) * A.nlog, A.nlog, A.score_mode);
      ok = score > A.score_thresh;  // inference.py:142
      if (ok) box = decode_box(A.props[row0 + i], A.reg + (row0 + i) * A.regc + A.reg_off, A, xmax, ymax);
    }
    int tot;
    const int pos = running + block_exclusive_scan(ok ? 1 : 0, warp_tot, tot);
    if (ok) {
      A.cand_boxes[row0 + pos] = box;
      A.cand_scores[row0 + pos] = score;
      A.cand_src[row0 + pos] = i;
    }
    running += tot;
  }
  if (tid == 0) A.cand_count[b] = running;
}

struct BoxBuffers {
  float4* cand_boxes;
  float* cand_scores;
  int32_t* cand_src;
  int32_t* cand_count;
  int32_t* kept_total;
  NmsWorkspace nms;
};

int validate(const osd_box_post_config* cfg) {
  OSD_REQUIRE(cfg != nullptr, "box_post: config is null");
  OSD_REQUIRE(cfg->batch >= 0 && cfg->batch <= 65535, "box_post: batch %d out of range", cfg->batch);
  OSD_REQUIRE(cfg->rois_per_image >= 1 && cfg->rois_per_image < (1 << 24), "box_post: rois_per_image %d out of range",
              cfg->rois_per_image);
  OSD_REQUIRE(cfg->score_mode == OSD_SCORE_SOFTMAX || cfg->score_mode == OSD_SCORE_SIGMOID, "box_post: unknown score_mode %d",
              cfg->score_mode);
  OSD_REQUIRE(cfg->num_logits >= (cfg->score_mode == OSD_SCORE_SOFTMAX ? 2 : 1) && cfg->num_logits <= 1024,
              "box_post: num_logits %d out of range for this score_mode", cfg->num_logits);
  OSD_REQUIRE(cfg->reg_offset >= 0 && cfg->reg_columns >= cfg->reg_offset + 4,
              "box_post: regression columns [%d, %d) do not fit a row of %d", cfg->reg_offset, cfg->reg_offset + 4,
              cfg->reg_columns);
  for (int k = 0; k < 4; ++k)
    OSD_REQUIRE(cfg->weights[k] > 0.f, "box_post: BoxCoder weight %d must be positive", k);
  return OSD_OK;
}

void carve(const osd_box_post_config* cfg, Carver& c, BoxBuffers* buf, osd_box_post_plan* plan) {
  const int B = cfg->batch > 0 ? cfg->batch : 1;
  const int R = cfg->rois_per_image;
  BoxBuffers b{};
  const size_t o_boxes = c.offset_of_next();
  b.cand_boxes = c.take<float4>((size_t)B * R);
  const size_t o_scores = c.offset_of_next();
  b.cand_scores = c.take<float>((size_t)B * R);
  const size_t o_src = c.offset_of_next();
  b.cand_src = c.take<int32_t>((size_t)B * R);
  const size_t o_cnt = c.offset_of_next();
  b.cand_count = c.take<int32_t>(B);
  const size_t o_kt = c.offset_of_next();
  b.kept_total = c.take<int32_t>(B);
  nms_workspace_carve(c, B, R, &b.nms);
  if (buf) *buf = b;
  if (plan) {
    plan->workspace_bytes = c.total();
    plan->cand_capacity = R;
    plan->out_capacity = (cfg->detections_per_img > 0 && cfg->detections_per_img < R) ? cfg->detections_per_img : R;
    plan->off_cand_boxes = o_boxes;
    plan->off_cand_scores = o_scores;
    plan->off_cand_src = o_src;
    plan->off_cand_count = o_cnt;
    plan->off_kept_count = o_kt;
  }
}

}  // namespace
}  // namespace osd

extern "C" int osd_box_postprocess_plan(const osd_box_post_config* cfg, osd_box_post_plan* plan) {
  OSD_REQUIRE(plan != nullptr, "osd_box_postprocess_plan: plan is null");
  int rc = osd::validate(cfg);
  if (rc != OSD_OK) return rc;
  osd::Carver c(nullptr);
  osd::carve(cfg, c, nullptr, plan);
  return OSD_OK;
}

extern "C" int osd_box_postprocess(const osd_box_post_config* cfg, const float* class_logits, const float* box_regression,
                                   const float* proposals, const int32_t* roi_count, const int32_t* image_hw,
                                   void* workspace, size_t workspace_bytes, float* out_boxes, float* out_scores,
                                   int32_t* out_index, int32_t* out_count, void* stream_) {
  using namespace osd;
  int rc = validate(cfg);
  if (rc != OSD_OK) return rc;
  if (cfg->batch == 0) return OSD_OK;
  OSD_REQUIRE(class_logits && box_regression && proposals && image_hw, "osd_box_postprocess: null input");
  OSD_REQUIRE(out_boxes && out_scores && out_index && out_count, "osd_box_postprocess: null output");
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(proposals) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0,
              "osd_box_postprocess: proposals and out_boxes must be 16-byte aligned");
  OSD_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "osd_box_postprocess: workspace must be 256-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver c(workspace);
  BoxBuffers buf{};
  osd_box_post_plan plan{};
  carve(cfg, c, &buf, &plan);
  if (plan.workspace_bytes > workspace_bytes) {
    set_error("osd_box_postprocess: workspace of %zu bytes needed, %zu given", plan.workspace_bytes, workspace_bytes);
    return OSD_ERR_WORKSPACE;
  }
  DecodeArgs A{};
  A.logits = class_logits;
  A.reg = box_regression;
  A.props = reinterpret_cast<const float4*>(proposals);
  A.roi_count = roi_count;
  A.image_hw = image_hw;
  A.R = cfg->rois_per_image;
  A.nlog = cfg->num_logits;
  A.regc = cfg->reg_columns;
  A.reg_off = cfg->reg_offset;
  A.score_mode = cfg->score_mode;
  A.wx = cfg->weights[0];
  A.wy = cfg->weights[1];
  A.ww = cfg->weights[2];
  A.wh = cfg->weights[3];
  A.clip = cfg->bbox_xform_clip;
  A.score_thresh = cfg->score_thresh;
  A.cand_boxes = buf.cand_boxes;
  A.cand_scores = buf.cand_scores;
  A.cand_src = buf.cand_src;
  A.cand_count = buf.cand_count;
  box_decode_kernel<<<cfg->batch, kDecThreads, 0, stream>>>(A);
  OSD_LAUNCH_CHECK("box_decode_kernel");

  CandLayout L{};
  L.boxes = buf.cand_boxes;
  L.scores = buf.cand_scores;
  L.seg = nullptr;
  L.level_count = buf.cand_count;   // one "level" per image: the compacted candidates
  L.nl = 1;
  L.cap = cfg->rois_per_image;
  L.slot[0] = 0;
  NmsParams P{};
  P.thr = cfg->nms_thresh;
  P.strict = cfg->strict ? 1 : 0;
  P.post_top_n = cfg->detections_per_img;
  P.early_exit = cfg->early_exit ? 1 : 0;
  P.max_len = cfg->rois_per_image;
  P.passthrough = !(cfg->nms_thresh > 0.0f);  // boxlist_ops.py:22-23
  NmsOutputs O{};
  O.out_boxes = out_boxes;
  O.out_scores = out_scores;
  O.out_index = out_index;
  O.out_count = out_count;
  O.K = plan.out_capacity;
  O.kept_total = buf.kept_total;
  return nms_run(L, buf.nms, P, O, stream);
}
