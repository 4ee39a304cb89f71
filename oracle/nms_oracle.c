/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference's greedy NMS,
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the
 * checker.  The product path never links or calls this file.
 *
 * Parity pinned: yes -- against the reference's own known-answer vectors
 * (/root/reference/tests/test_nms.py:16-58 and :65-217, committed as
 * tests/golden/nms_kat.json) and against the reference's own nms_cpu compiled from
 * source (oracle/_ref/osd_ref_C.so) on seeded random inputs (tests/test_oracle_nms.py).
 *
 * Follows maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp:5-64 (nms_cpu_kernel<float>):
 *   :22     areas = (x2 - x1 + 1) * (y2 - y1 + 1)          -- three rounded fp32 ops per factor
 *   :24     order = scores.sort(0, descending=true)        -- ATen's sort is not stable; this
 *                                                             restatement takes the order as an
 *                                                             argument, or builds the canonical
 *                                                             (score desc, index asc) order
 *   :37-63  for i in order: skip if suppressed; for j after i in order: skip if suppressed;
 *           w = max(0, min(x2) - max(x1) + 1), h likewise, inter = w*h,
 *           ovr = inter / (area_i + area_j - inter); suppress j if ovr >= thr
 *   :64     return nonzero(suppressed == 0)                -- ascending ORIGINAL indices
 * The reference CUDA kernel (csrc/cuda/nms.cu:60) tests `ovr > thr` instead; `strict` != 0
 * selects that comparison.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (oracle/build_ref.py); with contraction off
 * every statement below is one IEEE-754 binary32 operation, as in the reference's x86-64
 * build (no -mfma there, so no fused multiply-add).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* std::max / std::min semantics of the reference (nms_cpu.cpp:51-57): (a < b) ? b : a */
static inline float max_f(float a, float b) { return (a < b) ? b : a; }
static inline float min_f(float a, float b) { return (b < a) ? b : a; }

/* canonical order: score descending, original index ascending on ties (stable). */
static void merge_sort_desc(const float *scores, int64_t *idx, int64_t *tmp, int64_t n) {
  for (int64_t width = 1; width < n; width *= 2) {
    for (int64_t lo = 0; lo < n; lo += 2 * width) {
      int64_t mid = lo + width < n ? lo + width : n;
      int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
      int64_t a = lo, b = mid, k = lo;
      while (a < mid && b < hi) {
        /* take from the right run only if strictly greater -> stable */
        if (scores[idx[b]] > scores[idx[a]]) tmp[k++] = idx[b++];
        else tmp[k++] = idx[a++];
      }
      while (a < mid) tmp[k++] = idx[a++];
      while (b < hi) tmp[k++] = idx[b++];
    }
    memcpy(idx, tmp, (size_t)n * sizeof(int64_t));
  }
}

/* Writes the canonical order for `scores` into order_out[n]. */
void osd_oracle_stable_order_f32(const float *scores, int64_t n, int64_t *order_out) {
  int64_t *tmp = (int64_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) order_out[i] = i;
  merge_sort_desc(scores, order_out, tmp, n);
  free(tmp);
}

/*
 * boxes  [n,4] xyxy fp32, scores [n] fp32, thr fp32.
 * order  optional int64[n] visiting order (NULL -> canonical stable order).
 * strict 0: suppress if ovr >= thr (nms_cpu.cpp:60); 1: ovr > thr (nms.cu:60).
 * keep_out int64[n]; returns number kept; indices are ascending original indices.
 */
int64_t osd_oracle_nms_f32(const float *boxes, const float *scores, int64_t n, float thr,
                           const int64_t *order, int strict, int64_t *keep_out) {
  if (n <= 0) return 0;
  float *areas = (float *)malloc((size_t)n * sizeof(float));
  uint8_t *suppressed = (uint8_t *)calloc((size_t)n, 1);
  int64_t *ord = (int64_t *)malloc((size_t)n * sizeof(int64_t));
  if (order) memcpy(ord, order, (size_t)n * sizeof(int64_t));
  else osd_oracle_stable_order_f32(scores, n, ord);

  for (int64_t i = 0; i < n; ++i) {
    const float *b = boxes + 4 * i;
    float w = b[2] - b[0];
    w = w + 1.0f;
    float h = b[3] - b[1];
    h = h + 1.0f;
    areas[i] = w * h;
  }

  for (int64_t _i = 0; _i < n; ++_i) {
    int64_t i = ord[_i];
    if (suppressed[i]) continue;
    float ix1 = boxes[4 * i + 0], iy1 = boxes[4 * i + 1];
    float ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < n; ++_j) {
      int64_t j = ord[_j];
      if (suppressed[j]) continue;
      float xx1 = max_f(ix1, boxes[4 * j + 0]);
      float yy1 = max_f(iy1, boxes[4 * j + 1]);
      float xx2 = min_f(ix2, boxes[4 * j + 2]);
      float yy2 = min_f(iy2, boxes[4 * j + 3]);
      float w = xx2 - xx1;
      w = w + 1.0f;
      w = max_f(0.0f, w);
      float h = yy2 - yy1;
      h = h + 1.0f;
      h = max_f(0.0f, h);
      float inter = w * h;
      float uni = iarea + areas[j];
      uni = uni - inter;
      float ovr = inter / uni;
      if (strict ? (ovr > thr) : (ovr >= thr)) suppressed[j] = 1;
    }
  }
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i)
    if (!suppressed[i]) keep_out[k++] = i;
  free(areas);
  free(suppressed);
  free(ord);
  return k;
}

/*
 * Segmented form: E independent problems laid end to end (episode e owns
 * [seg[e], seg[e+1])).  Boxes never suppress across segments
 * (modeling/rpn/fcos/inference.py:289-323 runs one boxlist_nms per image).
 * keep_out receives GLOBAL row indices, segment after segment; counts_out[e] per segment.
 */
int64_t osd_oracle_batched_nms_f32(const float *boxes, const float *scores, const int64_t *seg,
                                   int64_t num_seg, float thr, int strict, int64_t *keep_out,
                                   int64_t *counts_out) {
  int64_t total = 0;
  for (int64_t e = 0; e < num_seg; ++e) {
    int64_t lo = seg[e], n = seg[e + 1] - seg[e];
    int64_t k = osd_oracle_nms_f32(boxes + 4 * lo, scores + lo, n, thr, NULL, strict,
                                   keep_out + total);
    for (int64_t t = 0; t < k; ++t) keep_out[total + t] += lo;
    counts_out[e] = k;
    total += k;
  }
  return total;
}
