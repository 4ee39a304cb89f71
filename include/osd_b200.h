/*
 * osd_b200.h -- C ABI of the B200-native OneshotDet hot path (libosd_b200.so).
 *
 * Plain C: device/host pointers, sizes and a CUDA stream handle; no torch types.  Every entry
 * point enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises the device,
 * never allocates device memory (the caller provides the workspace that the matching *_plan call
 * sized) and returns 0 on success or a negative osd_status; osd_last_error() returns the message of
 * the calling thread's last failure.  Floating-point inputs are fp32 unless a dtype field says
 * otherwise.  All device pointers must belong to the device that is current on the calling thread.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the reference
 * repository root, RyanXLi/OneshotDet).
 */
#ifndef OSD_B200_H_
#define OSD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OSD_MAX_LEVELS 8

typedef enum {
  OSD_OK = 0,
  OSD_ERR_INVALID = -1,    /* bad argument (shape, alignment, null pointer, unsupported mode) */
  OSD_ERR_WORKSPACE = -2,  /* workspace too small for the plan */
  OSD_ERR_CUDA = -3,       /* a CUDA runtime call failed; see osd_last_error() */
  OSD_ERR_UNSUPPORTED = -4 /* device is not sm_100 */
} osd_status;

/* Library version (major*10000 + minor*100 + patch). */
int osd_version(void);
/* Message for the last error raised on this thread ("" if none). */
const char* osd_last_error(void);
/* 0 if the current device can run the sm_100a kernels, OSD_ERR_UNSUPPORTED otherwise. */
int osd_check_device(void);
/* Number of kernels this library has launched on this thread since the last reset (bench accounting). */
int64_t osd_launch_count(void);
void osd_reset_launch_count(void);
/* Profiling aid (replaces the reference's host-side Timer, utils/timer.py, for this path): when enabled (or with
 * OSD_TIMELINE=1 in the environment) every kernel launch of the library is followed by a CUDA event on its stream;
 * osd_timeline_read synchronises the device and writes "name milliseconds-since-first-mark" lines, returns the
 * number of marks and clears them.  Launches inside a stream capture are not marked. */
void osd_timeline_enable(int on);
int osd_timeline_read(char* buf, size_t capacity);

/* ------------------------------------------------------------------------------------------------
 * Batched NMS.
 *
 * Replaces  maskrcnn_benchmark/csrc/nms.h:10-28  (`at::Tensor nms(dets, scores, threshold)`, the
 *           pybind export `_C.nms` of csrc/vision.cpp:8) with its CPU and CUDA kernels
 *           csrc/cpu/nms_cpu.cpp:5-75 and csrc/cuda/nms.cu:23-131, as called through
 *           maskrcnn_benchmark/structures/boxlist_ops.py:10-34 (boxlist_nms).
 *
 * E independent problems ("segments" = episodes) are solved in one call: segment e owns rows
 * [seg_offsets[e], seg_offsets[e+1]) of boxes/scores.  Semantics per segment are exactly
 * nms_cpu_kernel's: visiting order = score descending (ties: lower index first), legacy +1 box
 * widths, fp32 IoU = inter / (area_i + area_j - inter) with IEEE rounding of every operation,
 * suppress when IoU >= threshold (strict = 0, nms_cpu.cpp:60) or IoU > threshold (strict = 1,
 * nms.cu:60).  Kept rows are written as ascending GLOBAL row indices (int64) to
 * keep_out[seg_offsets[e] .. seg_offsets[e] + keep_counts[e]).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  size_t workspace_bytes; /* device scratch the call needs */
  int32_t padded_len;     /* per-segment row capacity (max_seg_len rounded up to 64) */
  int32_t mask_words;     /* 64-bit words per bitmask row */
} osd_nms_plan;

int osd_batched_nms_plan(int64_t num_segments, int64_t max_seg_len, osd_nms_plan* plan);

int osd_batched_nms(const float* boxes,          /* device [N,4] xyxy */
                    const float* scores,         /* device [N] */
                    const int64_t* seg_offsets,  /* device [E+1], ascending, seg_offsets[0] may be > 0 */
                    int64_t num_segments,        /* E */
                    int64_t max_seg_len,         /* host-side upper bound of any segment length */
                    float threshold, int strict,
                    void* workspace, size_t workspace_bytes,
                    int64_t* keep_out,           /* device, same extent as scores */
                    int32_t* keep_counts,        /* device [E] */
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * FCOS post-processing: score, per-level pre-NMS top-k, ltrb decode, clip, size filter, per-episode
 * NMS, post-NMS top-n -- one call for all levels and all episodes.
 *
 * Replaces  maskrcnn_benchmark/modeling/rpn/fcos/inference.py:251-281 (FCOSPostProcessor.forward at
 *           eval), i.e. :46-137 (forward_for_single_feature_map), :289-323 (select_over_all_levels),
 *           the location grid of modeling/rpn/fcos/fcos.py:220-234, BoxList.clip_to_image
 *           (structures/bounding_box.py:214-224) and remove_small_boxes (structures/boxlist_ops.py:202-216).
 *
 * Inputs per level l (NCHW-contiguous fp32, B episodes): cls[l] [B,1,H,W] logits, reg[l] [B,4,H,W]
 * (already exp'd ltrb distances, fcos.py:95), ctr[l] [B,1,H,W] logits.
 * score = sigmoid(cls) * sigmoid(ctr); a location is a candidate iff sigmoid(cls) > pre_nms_thresh;
 * per (episode, level) the pre_nms_top_n best scores survive; boxes are
 * (x - l, y - t, x + r, y + b) with (x, y) = (j*stride + stride/2, i*stride + stride/2), clipped to
 * [0, w-1] x [0, h-1] of the episode's image size, dropped unless both sides (+1) >= min_size.
 * Candidates of an episode are ordered level-major, location-ascending.  Then NMS (nms_thresh <= 0:
 * none) and, if more than post_nms_top_n > 0 boxes survive, the post_nms_top_n best by score in
 * descending order, else all survivors in ascending candidate order.
 * ---------------------------------------------------------------------------------------------- */
/* What the regression input is.  OSD_REG_DISTANCES: ltrb distances as FCOSHead.forward returns them (already
 * exp'd, fcos.py:95-97).  OSD_REG_RAW_EXP_SCALE: the raw output of the bbox_pred conv; the head's tail
 * torch.exp(self.scales[l](x)) = exp(x * scale_l) (fcos.py:95-97, layers/scale.py:10-11) is then evaluated here, for
 * the <= pre_nms_top_n selected locations of a level only -- the two elementwise passes over [B,4,H,W] disappear. */
typedef enum { OSD_REG_DISTANCES = 0, OSD_REG_RAW_EXP_SCALE = 1 } osd_reg_transform;

typedef struct {
  int32_t num_levels;
  int32_t batch;                   /* B episodes */
  int32_t height[OSD_MAX_LEVELS];  /* H_l */
  int32_t width[OSD_MAX_LEVELS];   /* W_l */
  int32_t stride[OSD_MAX_LEVELS];
  float pre_nms_thresh;
  int32_t pre_nms_top_n;
  float nms_thresh;
  int32_t post_nms_top_n;          /* fpn_post_nms_top_n; <= 0: unlimited */
  float min_size;
  int32_t strict;                  /* 0: IoU >= thr suppresses (nms_cpu), 1: IoU > thr (nms.cu) */
  int32_t early_exit;              /* 1: stop an episode's NMS once post_nms_top_n + 1 boxes are kept
                                      (result identical; only meaningful when post_nms_top_n > 0) */
  int32_t reg_transform;           /* osd_reg_transform: what reg[l] holds */
  float reg_scale[OSD_MAX_LEVELS]; /* OSD_REG_RAW_EXP_SCALE: the per-level Scale parameter (layers/scale.py:5-11) */
} osd_fcos_config;

typedef struct {
  size_t workspace_bytes;
  int32_t cand_capacity;   /* CAP: candidate slots per episode = sum_l min(H_l*W_l, pre_nms_top_n) */
  int32_t out_capacity;    /* K: rows per episode in the outputs */
  int32_t level_slot[OSD_MAX_LEVELS]; /* first candidate slot of each level inside an episode */
  /* byte offsets into the workspace of the intermediate results (for inspection / tests) */
  size_t off_cand_boxes;   /* float  [B, CAP, 4]  slotted per level */
  size_t off_cand_scores;  /* float  [B, CAP] */
  size_t off_cand_loc;     /* int32  [B, CAP]     location index inside the level */
  size_t off_level_count;  /* int32  [B, num_levels] candidates per (episode, level) */
  size_t off_kept_count;   /* int32  [B]          boxes kept by NMS before the post-NMS cut (>= post_nms_top_n + 1 means "more") */
} osd_fcos_plan;

int osd_fcos_postprocess_plan(const osd_fcos_config* cfg, osd_fcos_plan* plan);

int osd_fcos_postprocess(const osd_fcos_config* cfg,
                         const float* const* cls,   /* host array of num_levels device pointers */
                         const float* const* reg,
                         const float* const* ctr,
                         const int32_t* image_hw,   /* device int32 [B,2]: (h, w) per episode */
                         void* workspace, size_t workspace_bytes,
                         float* out_boxes,          /* device [B, K, 4] */
                         float* out_scores,         /* device [B, K] */
                         int32_t* out_index,        /* device [B, K] compact candidate index of each output row */
                         int32_t* out_count,        /* device [B] */
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Support -> target feature matching on the FPN levels.
 *
 * Replaces  maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:100-104 (batch_pooling, K-shot
 *           mean) and :306-311 (the inline product loop), and offers the concat form of
 *           modeling/roi_heads/box_head/box_head.py:147 (:144 reversed) on [B,C,H,W] maps.
 * ---------------------------------------------------------------------------------------------- */
typedef enum { OSD_MATCH_PRODUCT = 0, OSD_MATCH_CONCAT = 1, OSD_MATCH_CONCAT_REVERSED = 2 } osd_match_mode;
typedef enum { OSD_LAYOUT_NCHW = 0, OSD_LAYOUT_NHWC = 1 } osd_layout;
typedef enum { OSD_DTYPE_F32 = 0, OSD_DTYPE_BF16 = 1 } osd_dtype;

typedef struct {
  int32_t num_levels;
  int32_t batch;     /* B */
  int32_t shots;     /* S: supp[l] holds B*S embeddings, episode-major, shot-minor */
  int32_t channels;  /* C */
  int32_t mode;      /* osd_match_mode */
  int32_t layout;    /* osd_layout of feat/out (support vectors are [B*S, C] either way) */
  int32_t dtype;     /* osd_dtype of feat/out; supp has the same dtype */
  int32_t hw[OSD_MAX_LEVELS];         /* H_l * W_l */
  const void* feat[OSD_MAX_LEVELS];   /* device [B,C,H,W] (or [B,H,W,C]) */
  const void* supp[OSD_MAX_LEVELS];   /* device [B*S, C] */
  void* out[OSD_MAX_LEVELS];          /* device [B,C,H,W]; concat modes: [B,2C,H,W] */
} osd_match_desc;

int osd_match_forward(const osd_match_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 1x1 fusion conv matching mode on tcgen05 tensor cores (bf16 operands, fp32 accumulation).
 *
 * Replaces  `compress_dim_conv` of maskrcnn_benchmark/modeling/roi_heads/box_head/box_head.py:43-54 applied to
 *           cat((x, support.expand_as(x)), dim=1) (:147-149), here on the FPN maps [B,C,H,W] (NCHW fp32):
 *           Conv1x1(2C->2C) + GroupNorm(32,2C) + LeakyReLU + Conv1x1(2C->C) + GroupNorm(32,C) + LeakyReLU.
 * The support half of conv1 is folded into a per-episode bias (W1 = [W1x | W1s]).  Weight operands are prepared
 * once by the host: w1x_bf16 = bf16(W1[:, :C]) [2C, C]; w1s_t = W1[:, C:]^T fp32 [C, 2C]; w2_bf16 = bf16(W2) [C, 2C].
 * stage OSD_FUSION_CONV1 writes the first convolution only (out[l] is [B,2C,H,W]); OSD_FUSION_FULL writes
 * the module output (out[l] is [B,C,H,W]).
 * ---------------------------------------------------------------------------------------------- */
typedef enum { OSD_FUSION_CONV1 = 0, OSD_FUSION_FULL = 1 } osd_fusion_stage;

typedef struct {
  int32_t num_levels;
  int32_t batch;     /* B */
  int32_t shots;     /* S */
  int32_t channels;  /* C: 64, 128 or 256 */
  int32_t stage;     /* osd_fusion_stage */
  float gn_eps;      /* GroupNorm epsilon (torch default 1e-5) */
  float lrelu_slope; /* 0.2 */
  int32_t hw[OSD_MAX_LEVELS];
  const void* feat[OSD_MAX_LEVELS];  /* device fp32 [B,C,H,W] */
  const void* supp[OSD_MAX_LEVELS];  /* device fp32 [B*S, C] */
  void* out[OSD_MAX_LEVELS];         /* device fp32 [B,2C,H,W] (CONV1) or [B,C,H,W] (FULL) */
  const void* w1x_bf16;  /* device bf16 [2C, C] */
  const float* w1s_t;    /* device fp32 [C, 2C] */
  const float* b1;       /* device fp32 [2C] */
  const float* gn1_w;    /* device fp32 [2C] */
  const float* gn1_b;
  const void* w2_bf16;   /* device bf16 [C, 2C] */
  const float* b2;       /* device fp32 [C] */
  const float* gn2_w;    /* device fp32 [C] */
  const float* gn2_b;
  const float* w1x_gram; /* optional device fp32 [32, C, C]: M_g = sum over the channels c of GroupNorm-1 group g of
                            w_c w_c^T, w_c = row c of w1x_bf16 (as fp32).  With it (C >= 128) pass A takes the GroupNorm-1
                            statistics from a Gram GEMM of the activations with themselves instead of running conv1 */
} osd_fusion_desc;

int osd_fusion_workspace_bytes(const osd_fusion_desc* desc, size_t* bytes);
int osd_fusion_forward(const osd_fusion_desc* desc, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Support-embedding producer (the step right before the matching path).
 *
 * Replaces  SuppAlignLayer (maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:20-52, call :305): ROIAlign with a
 *           (1,1) output, one whole-image ROI per support, per FPN level (kernel csrc/cuda/ROIAlign_cuda.cu:65-122,
 *           CPU twin csrc/cpu/ROIAlign_cpu.cpp:113-214) -- mode OSD_POOL_ROIALIGN; and nn.AdaptiveAvgPool2d((1,1))
 *           (generalized_rcnn.py:94, :303) -- mode OSD_POOL_AVG.
 * feat[l] is [N, C, H_l, W_l] fp32 NCHW (N = B*S supports), out[l] is [N, C]; rois is [N, 4] (x1, y1, x2, y2) in
 * support-image coordinates, scaled by spatial_scale[l] per level.  OSD_POOL_ROIALIGN is bit-identical to the
 * reference CPU operator.
 * ---------------------------------------------------------------------------------------------- */
typedef enum { OSD_POOL_ROIALIGN = 0, OSD_POOL_AVG = 1 } osd_pool_mode;

typedef struct {
  int32_t num_levels;
  int32_t num_supports;   /* N = B*S */
  int32_t channels;
  int32_t mode;           /* osd_pool_mode */
  int32_t sampling_ratio; /* ROIAlign sampling ratio (<= 0: adaptive, ceil(roi extent)) */
  int32_t height[OSD_MAX_LEVELS];
  int32_t width[OSD_MAX_LEVELS];
  float spatial_scale[OSD_MAX_LEVELS];
  const void* feat[OSD_MAX_LEVELS];
  void* out[OSD_MAX_LEVELS];
  const float* rois;      /* device fp32 [N, 4]; unused for OSD_POOL_AVG */
} osd_support_pool_desc;

int osd_support_pool(const osd_support_pool_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Second-stage (ROI box head) post-processing -- SURVEY section 8(f) row 2, the step after the hot path.
 *
 * Replaces  PostProcessor.forward / prepare_boxlist / filter_results
 *           (maskrcnn_benchmark/modeling/roi_heads/box_head/inference.py:46-167), BoxCoder.decode
 *           (modeling/box_coder.py:52-95), BoxList.clip_to_image (structures/bounding_box.py:214-224) and the
 *           boxlist_nms call of :150-152, for the one-foreground-class episode problem (num_classes = 2, :89).
 *
 * class_logits [B*R, num_logits], box_regression [B*R, reg_columns] (class 1 uses columns
 * [reg_offset, reg_offset+4): 4 for the 8-column head output of :60 with or without CLS_AGNOSTIC_BBOX_REG),
 * proposals [B, R, 4] xyxy (R rows per image; roi_count[b] <= R valid rows when roi_count != NULL), image_hw [B,2].
 * score = softmax(logits)[1] (ce/cxe, :66-67) or sigmoid(logits[0]) (focal, :62-65); a proposal survives iff
 * score > score_thresh (:142); its box is BoxCoder.decode with `weights` and the dw/dh clamp, clipped to the image;
 * per image NMS(nms_thresh) in proposal order (:144-152), then, if more than detections_per_img > 0 survive, the best
 * detections_per_img by score in descending order (:162-166), else all survivors in ascending proposal order.
 * Outputs as osd_fcos_postprocess: out_index holds the compact candidate index (its proposal row is
 * cand_src[out_index], see the plan offsets).  No host synchronisation.
 * ---------------------------------------------------------------------------------------------- */
typedef enum { OSD_SCORE_SOFTMAX = 0, OSD_SCORE_SIGMOID = 1 } osd_score_mode;

typedef struct {
  int32_t batch;               /* B images */
  int32_t rois_per_image;      /* R */
  int32_t num_logits;          /* columns of class_logits */
  int32_t reg_columns;         /* columns of box_regression */
  int32_t reg_offset;          /* first regression column of class 1 */
  int32_t score_mode;          /* osd_score_mode */
  float weights[4];            /* BoxCoder weights (wx, wy, ww, wh); MODEL.ROI_HEADS.BBOX_REG_WEIGHTS */
  float bbox_xform_clip;       /* box_coder.py:21: log(1000/16) */
  float score_thresh;          /* MODEL.ROI_HEADS.SCORE_THRESH */
  float nms_thresh;            /* MODEL.ROI_HEADS.NMS; <= 0: no suppression */
  int32_t detections_per_img;  /* MODEL.ROI_HEADS.DETECTIONS_PER_IMG; <= 0: unlimited */
  int32_t strict;              /* as osd_fcos_config */
  int32_t early_exit;
} osd_box_post_config;

typedef struct {
  size_t workspace_bytes;
  int32_t cand_capacity;   /* R */
  int32_t out_capacity;    /* K: rows per image in the outputs */
  size_t off_cand_boxes;   /* float [B, R, 4] decoded + clipped boxes of the surviving proposals, compacted */
  size_t off_cand_scores;  /* float [B, R] */
  size_t off_cand_src;     /* int32 [B, R] proposal row of each candidate */
  size_t off_cand_count;   /* int32 [B] */
  size_t off_kept_count;   /* int32 [B] kept by NMS before the cut */
} osd_box_post_plan;

int osd_box_postprocess_plan(const osd_box_post_config* cfg, osd_box_post_plan* plan);

int osd_box_postprocess(const osd_box_post_config* cfg,
                        const float* class_logits, const float* box_regression, const float* proposals,
                        const int32_t* roi_count,  /* device int32 [B] or NULL */
                        const int32_t* image_hw,   /* device int32 [B,2]: (h, w) */
                        void* workspace, size_t workspace_bytes,
                        float* out_boxes, float* out_scores, int32_t* out_index, int32_t* out_count,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-level ROI pooler of the second stage -- SURVEY section 8(f) row 2, pooling half.
 *
 * Replaces  Pooler.forward (maskrcnn_benchmark/modeling/poolers.py:93-125) with LevelMapper (:10-41) and the
 *           per-level ROIAlign modules (layers/roi_align.py; kernel csrc/cuda/ROIAlign_cuda.cu:65-122, CPU twin
 *           csrc/cpu/ROIAlign_cpu.cpp:14-214).
 * feat[l] [B, C, H_l, W_l] fp32 NCHW; rois [B, R, 4] xyxy in image coordinates (ROI (b, r) reads image b -- the
 * (img_id, box) rows of convert_to_roi_format, poolers.py:77-91); out [B*R, C, P, P] fp32, the reference's layout
 * (viewed [B, R, C, P, P] at poolers.py:123).  ROI -> level: floor(canonical_level + log2(sqrt(area)/canonical_scale
 * + eps)) clamped to [k_min, k_max], minus k_min; a single level skips the mapper (:104-105).  Rows r >= roi_count[b]
 * (when roi_count != NULL) are written as zeros, levels_out = -1.  Bit-identical to the reference CPU operator.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t num_levels;
  int32_t batch;            /* B */
  int32_t rois_per_image;   /* R */
  int32_t channels;         /* C */
  int32_t pooled_size;      /* P: MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION */
  int32_t sampling_ratio;   /* MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO (<= 0: adaptive) */
  int32_t height[OSD_MAX_LEVELS];
  int32_t width[OSD_MAX_LEVELS];
  float spatial_scale[OSD_MAX_LEVELS];   /* MODEL.ROI_BOX_HEAD.POOLER_SCALES */
  int32_t k_min, k_max;     /* -log2(scales[0]), -log2(scales[-1])  (poolers.py:72-74) */
  float canonical_scale;    /* 224 */
  int32_t canonical_level;  /* 4 */
  float eps;                /* 1e-6 */
  const void* feat[OSD_MAX_LEVELS];
  const float* rois;        /* device [B, R, 4], 16-byte aligned */
  const int32_t* roi_count; /* device int32 [B] or NULL */
  float* out;               /* device [B*R, C, P, P] */
  int32_t* levels_out;      /* device int32 [B*R] or NULL: the level each ROI was pooled from */
  void* workspace;          /* device scratch of osd_roi_pool_workspace_bytes() bytes (256-byte aligned) for the
                               channels-last copy of the maps; NULL: pool straight from NCHW (slower, same result) */
  size_t workspace_bytes;
  void* out_nhwc_bf16;      /* optional device bf16 [B*R, P*P, C] (needs `workspace`): the same values rounded to bf16 in the
                               K-major row layout osd_box_head_forward reads (pooled_nhwc_bf16); `out` may then be NULL */
} osd_roi_pool_desc;

int osd_roi_pool_workspace_bytes(const osd_roi_pool_desc* desc, size_t* bytes);
int osd_roi_pool(const osd_roi_pool_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Second-stage dense head -- SURVEY section 8(f) row 2, the part between the pooler and the post-processor.
 *
 * Replaces  ROIBoxHead.forward for comparison_method 'concat', one support, no negative support, LINEAR_FUSION off
 *           (maskrcnn_benchmark/modeling/roi_heads/box_head/box_head.py:118-157): cat((x, support.expand_as(x)), 1)
 *           -> compress_dim_conv (:43-54) -> feature_aggreg (:62-67) -> relu(fc6) -> relu(fc7) (:75-76, :152-154) and
 *           FPNPredictor.forward (modeling/roi_heads/box_head/roi_box_predictors.py:80-84).
 * pooled [B*R, C, 7, 7] fp32 is the Pooler's output (osd_roi_pool), supp [B, C, 7, 7] the ROI-pooled support of each
 * episode.  Six GEMM launches per chunk of ROIs on tcgen05 (bf16 operands, fp32 accumulation in tensor memory), each
 * with its bias / per-ROI GroupNorm / LeakyReLU / ReLU in the epilogue; activations between layers are bf16.
 * Weights are prepared by the host once (bf16, K-major; see oneshotdet_b200/box_head.py for the permutations):
 *   w1 [2C, 2C]  compress_dim_conv.0.weight            w2 [C, 2C]  compress_dim_conv.3.weight
 *   w3 [C/2, 9C] feature_aggreg.0.weight as (out, ky, kx, in)
 *   w6 [mlp, 49*C/2] fc6.weight as (out, pixel, channel)  w7 [mlp, mlp]  fc7.weight
 *   wp [num_classes + num_box_out, mlp]  cls_score.weight rows, then bbox_pred.weight rows; bp likewise
 * Outputs fp32: class_logits [B*R, num_classes], box_regression [B*R, num_box_out].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch;            /* B episodes */
  int32_t rois_per_image;   /* R */
  int32_t channels;         /* C: 64, 128 or 256 */
  int32_t pooled_size;      /* 7 (MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION) */
  int32_t mlp_dim;          /* MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM, a multiple of 8 */
  int32_t num_classes;      /* rows of cls_score (2) */
  int32_t num_box_out;      /* rows of bbox_pred (8) */
  int32_t roi_chunk;        /* ROIs per pass over the six layers; <= 0: sized so a chunk's activations stay in L2 */
  float gn_eps;             /* 1e-5 */
  float lrelu_slope;        /* 0.2 */
  const float* pooled;      /* device [B*R, C, 7, 7], or NULL when pooled_nhwc_bf16 is given */
  const float* supp;        /* device [B, C, 7, 7] */
  const void* w1; const float* b1; const float* gn1_w; const float* gn1_b;
  const void* w2; const float* b2; const float* gn2_w; const float* gn2_b;
  const void* w3; const float* b3; const float* gn3_w; const float* gn3_b;
  const void* w6; const float* b6;
  const void* w7; const float* b7;
  const void* wp; const float* bp;
  float* class_logits;      /* device [B*R, num_classes] */
  float* box_regression;    /* device [B*R, num_box_out] */
  const void* pooled_nhwc_bf16;  /* optional device bf16 [B*R, 49, C] written by osd_roi_pool (out_nhwc_bf16): skips the
                                    fp32 round trip and the repacking pass */
} osd_box_head_desc;

int osd_box_head_workspace_bytes(const osd_box_head_desc* desc, size_t* bytes);
int osd_box_head_forward(const osd_box_head_desc* desc, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Result hand-off: detections -> COCO detection records -> the reference's JSON file.  SURVEY section 8(f) row 4.
 *
 * Replaces  the per-image body of prepare_for_coco_detection
 *           (maskrcnn_benchmark/data/datasets/evaluation/coco/coco_eval.py:137-156): BoxList.resize to the original
 *           image size (structures/bounding_box.py:91-127), convert("xywh") (:55-73), tolist(), one dict per box;
 *           and the json.dump(..., sort_keys=True, indent=4, separators=(',', ':')) of :163-165.
 * osd_coco_records (device): boxes [E,K,4] / scores [E,K] / count [E] as osd_fcos_postprocess / osd_box_postprocess
 * emit them; det_wh [E,2] = the (w, h) the detections live in; orig_wh [E,2] = (width, height) of img_info.  Writes
 * records [sum count, 5] = (x, y, w, h, score), episode-major in detection order, record_episode [sum count], total [1].
 * osd_coco_write_json (host, no GPU): records on the HOST; writes exactly the bytes CPython's json.dump writes for the
 * reference's list of dicts {"bbox": [x, y, w, h], "category_id": c, "image_id": i, "score": s}.
 * ---------------------------------------------------------------------------------------------- */
int osd_coco_records(const float* boxes, const float* scores, const int32_t* count,
                     const int32_t* det_wh, const int32_t* orig_wh,
                     int32_t num_episodes, int32_t rows_per_episode,
                     float* records, int32_t* record_episode, int32_t* total, void* stream);

int osd_coco_write_json(const float* records, const int32_t* record_episode, int64_t num_records,
                        const int64_t* image_ids, const int64_t* category_ids, int32_t num_episodes,
                        const char* path);

/* ------------------------------------------------------------------------------------------------
 * Detection exchange over peer memory -- SURVEY section 8(e).
 *
 * Replaces  all_gather / scatter_gather of maskrcnn_benchmark/utils/comm.py:48-88 (pickle -> ByteTensor -> two
 *           dist.all_gather -> unpickle) as used by engine/inference.py:133-152 (_accumulate_predictions_from_multiple_gpus)
 *           for the fixed-shape result block of this path.
 *
 * Every rank (one process per GPU of one NVLink/NVSwitch box) owns a receive buffer allocated by osd_comm_alloc and
 * exports it as a 64-byte CUDA IPC handle; the handles travel through the host-side process group once, at set-up,
 * and osd_comm_import maps each peer's buffer into this process.  A step's result block is then PUSHED into every
 * rank's buffer -- an all-gather with no rendezvous and no collective kernel: osd_comm_push issues one peer copy per
 * destination on the copy engines (no SM is taken from the step's own kernels); osd_comm_push_kernel is the same
 * exchange as ONE small kernel of ours storing through the mapped peer pointers (st.global over NVLink), for when the
 * copy engines are busy with host traffic.  Both enqueue on `stream` and do not synchronise.  Ordering between ranks
 * (when a receive slot may be overwritten, when it is complete) is the caller's: see PeerBlockGatherer.
 * ---------------------------------------------------------------------------------------------- */
#define OSD_IPC_HANDLE_BYTES 64
int osd_comm_alloc(size_t bytes, void** ptr);   /* zero-filled device buffer outside any caching allocator (IPC-exportable) */
int osd_comm_free(void* ptr);
int osd_comm_export(void* ptr, unsigned char handle[OSD_IPC_HANDLE_BYTES]);
int osd_comm_import(const unsigned char handle[OSD_IPC_HANDLE_BYTES], void** ptr);   /* peer access is enabled lazily */
int osd_comm_close(void* ptr);                  /* unmaps an imported buffer */
int osd_comm_push(void* const* dst, int32_t num_dst, const void* src, size_t bytes, void* stream);
int osd_comm_push_kernel(void* const* dst, int32_t num_dst, const void* src, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSD_B200_H_ */
