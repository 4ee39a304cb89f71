// Multi-level ROI pooler of the second stage (pooling half): LevelMapper + ROIAlign(PxP).
//
// Reference: maskrcnn_benchmark/modeling/poolers.py:10-41 (LevelMapper: floor(lvl0 + log2(sqrt(area)/s0 + eps)), clamped),
// :93-125 (Pooler.forward: per level nonzero / gather / ROIAlign / scatter into a zero-initialised result),
// kernel csrc/cuda/ROIAlign_cuda.cu:65-122, CPU twin csrc/cpu/ROIAlign_cpu.cpp:14-214.
//
// One launch for all levels and all ROIs, no per-level gather/scatter, no host sync: a CTA per ROI derives the level
// from the box, builds the per-axis sample tables once (what pre_calc_for_bilinear_interpolate does per ROI,
// ROIAlign_cpu.cpp:14-110) and then walks the C*P*P outputs in their memory order, so the writes of a warp are one
// contiguous 128-byte run and its taps fall into a few rows of one or two channel planes.  Every arithmetic operation
// is an explicit _rn intrinsic in the order of ROIAlign_cpu.cpp, so fp32 results are bit-identical to the reference
// CPU operator.
#include "osd_common.cuh"

namespace osd {
namespace {

constexpr int kPoolThreads = 256;
constexpr int kTab = 128;   // per-axis sample-table entries held in shared memory (P * grid); larger grids are computed inline

struct RoiPoolArgs {
  int nl, B, R, C, P, sampling;
  const float* feat[OSD_MAX_LEVELS];
  int H[OSD_MAX_LEVELS], W[OSD_MAX_LEVELS];
  float scale[OSD_MAX_LEVELS];
  const float4* rois;
  const int32_t* roi_count;
  float k_min, k_max, s0, lvl0, eps;
  float* out;
  int32_t* levels_out;
};

struct AxisTap {
  int lo, hi;
  float l, h;   // weights of hi / lo
  int valid;
};

// one axis of pre_calc_for_bilinear_interpolate (ROIAlign_cpu.cpp:36-96): sample i of bin p
__device__ __forceinline__ AxisTap axis_tap(float start, float bin, int p, int i, int grid, int extent) {
  // roi_start + p * bin_size + (i + .5f) * bin_size / grid
  float v = __fadd_rn(__fadd_rn(start, __fmul_rn((float)p, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)grid));
  AxisTap t;
  t.valid = !(v < -1.0f || v > (float)extent);
  if (v <= 0.f) v = 0.f;
  int lo = (int)v, hi;
  if (lo >= extent - 1) {
    hi = lo = extent - 1;
    v = (float)lo;
  } else {
    hi = lo + 1;
  }
  t.lo = lo;
  t.hi = hi;
  t.l = __fsub_rn(v, (float)lo);
  t.h = __fsub_rn(1.0f, t.l);
  return t;
}

__global__ void __launch_bounds__(kPoolThreads) roi_pool_kernel(RoiPoolArgs A) {
  __shared__ AxisTap ytab[kTab], xtab[kTab];
  const int roi = blockIdx.x;          // b * R + r
  const int b = roi / A.R, r = roi - b * A.R;
  const int tid = threadIdx.x;
  const int PP = A.P * A.P, per_roi = A.C * PP;
  float* out = A.out + (size_t)roi * per_roi;
  const int n_valid = A.roi_count ? min(max(A.roi_count[b], 0), A.R) : A.R;
  if (r >= n_valid) {  // padded row: the reference's result buffer starts as zeros (poolers.py:113-117)
    for (int e = tid; e < per_roi; e += kPoolThreads) out[e] = 0.f;
    if (tid == 0 && A.levels_out) A.levels_out[roi] = -1;
    return;
  }
  const float4 box = A.rois[roi];
  // LevelMapper (poolers.py:33-41); BoxList.area (structures/bounding_box.py:226-238, TO_REMOVE = 1)
  const float area = __fmul_rn(__fadd_rn(__fsub_rn(box.z, box.x), 1.0f), __fadd_rn(__fsub_rn(box.w, box.y), 1.0f));
  const float s = __fsqrt_rn(area);
  float lv = floorf(__fadd_rn(A.lvl0, log2f(__fadd_rn(__fdiv_rn(s, A.s0), A.eps))));
  lv = fminf(fmaxf(lv, A.k_min), A.k_max);   // a NaN area (malformed box) lands on k_min
  const int l = A.nl == 1 ? 0 : (int)(lv - A.k_min);
  if (tid == 0 && A.levels_out) A.levels_out[roi] = l;
  const int H = A.H[l], W = A.W[l];
  const float sc = A.scale[l];
  // ROIAlign_cpu.cpp:147-171
  const float roi_start_w = __fmul_rn(box.x, sc), roi_start_h = __fmul_rn(box.y, sc);
  const float roi_end_w = __fmul_rn(box.z, sc), roi_end_h = __fmul_rn(box.w, sc);
  const float roi_width = fmaxf(__fsub_rn(roi_end_w, roi_start_w), 1.0f);
  const float roi_height = fmaxf(__fsub_rn(roi_end_h, roi_start_h), 1.0f);
  const float bin_h = __fdiv_rn(roi_height, (float)A.P), bin_w = __fdiv_rn(roi_width, (float)A.P);
  const int gh = A.sampling > 0 ? A.sampling : (int)ceilf(__fdiv_rn(roi_height, (float)A.P));
  const int gw = A.sampling > 0 ? A.sampling : (int)ceilf(__fdiv_rn(roi_width, (float)A.P));
  const float count = (float)(gh * gw);
  const bool tabbed = A.P * gh <= kTab && A.P * gw <= kTab;
  if (tabbed) {
    for (int t = tid; t < A.P * gh; t += kPoolThreads) ytab[t] = axis_tap(roi_start_h, bin_h, t / gh, t % gh, gh, H);
    for (int t = tid; t < A.P * gw; t += kPoolThreads) xtab[t] = axis_tap(roi_start_w, bin_w, t / gw, t % gw, gw, W);
  }
  __syncthreads();
  const float* fmap = A.feat[l] + (size_t)b * A.C * H * W;
  for (int e = tid; e < per_roi; e += kPoolThreads) {
    const int c = e / PP, bin = e - c * PP, ph = bin / A.P, pw = bin - ph * A.P;
    const float* __restrict__ plane = fmap + (size_t)c * H * W;
    float acc = 0.f;
    for (int iy = 0; iy < gh; ++iy) {
      const AxisTap ty = tabbed ? ytab[ph * gh + iy] : axis_tap(roi_start_h, bin_h, ph, iy, gh, H);
      const float* row_lo = plane + ty.lo * W;
      const float* row_hi = plane + ty.hi * W;
      for (int ix = 0; ix < gw; ++ix) {
        const AxisTap tx = tabbed ? xtab[pw * gw + ix] : axis_tap(roi_start_w, bin_w, pw, ix, gw, W);
        if (!(ty.valid && tx.valid)) continue;   // all-zero weights in the reference: adds +0
        const float w1 = __fmul_rn(ty.h, tx.h), w2 = __fmul_rn(ty.h, tx.l), w3 = __fmul_rn(ty.l, tx.h), w4 = __fmul_rn(ty.l, tx.l);
        const float v1 = __ldg(row_lo + tx.lo), v2 = __ldg(row_lo + tx.hi), v3 = __ldg(row_hi + tx.lo), v4 = __ldg(row_hi + tx.hi);
        // ROIAlign_cpu.cpp:199-202: output_val += w1*v1 + w2*v2 + w3*v3 + w4*v4
        acc = __fadd_rn(acc, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)),
                                       __fmul_rn(w4, v4)));
      }
    }
    out[e] = __fdiv_rn(acc, count);
  }
}

}  // namespace
}  // namespace osd

extern "C" int osd_roi_pool(const osd_roi_pool_desc* d, void* stream_) {
  using namespace osd;
  OSD_REQUIRE(d != nullptr, "osd_roi_pool: desc is null");
  OSD_REQUIRE(d->num_levels >= 1 && d->num_levels <= OSD_MAX_LEVELS, "osd_roi_pool: num_levels %d out of range", d->num_levels);
  OSD_REQUIRE(d->batch >= 0 && d->rois_per_image >= 0 && d->channels >= 1, "osd_roi_pool: bad sizes");
  OSD_REQUIRE(d->pooled_size >= 1 && d->pooled_size <= 64, "osd_roi_pool: pooled_size %d out of range", d->pooled_size);
  OSD_REQUIRE((int64_t)d->channels * d->pooled_size * d->pooled_size < (1ll << 30), "osd_roi_pool: ROI output too large");
  OSD_REQUIRE(d->num_levels == 1 || d->k_max - d->k_min + 1 == d->num_levels,
              "osd_roi_pool: LevelMapper range [%d, %d] does not match %d levels", d->k_min, d->k_max, d->num_levels);
  const int64_t n = (int64_t)d->batch * d->rois_per_image;
  if (n == 0) return OSD_OK;
  OSD_REQUIRE(n < (1ll << 31), "osd_roi_pool: too many ROIs");
  OSD_REQUIRE(d->rois != nullptr && d->out != nullptr, "osd_roi_pool: null rois / out");
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(d->rois) & 15) == 0, "osd_roi_pool: rois must be 16-byte aligned");
  RoiPoolArgs A{};
  A.nl = d->num_levels; A.B = d->batch; A.R = d->rois_per_image; A.C = d->channels; A.P = d->pooled_size;
  A.sampling = d->sampling_ratio;
  for (int l = 0; l < d->num_levels; ++l) {
    OSD_REQUIRE(d->feat[l] != nullptr, "osd_roi_pool: null feature map at level %d", l);
    OSD_REQUIRE(d->height[l] >= 1 && d->width[l] >= 1, "osd_roi_pool: empty level %d", l);
    A.feat[l] = static_cast<const float*>(d->feat[l]);
    A.H[l] = d->height[l]; A.W[l] = d->width[l]; A.scale[l] = d->spatial_scale[l];
  }
  A.rois = reinterpret_cast<const float4*>(d->rois);
  A.roi_count = d->roi_count;
  A.k_min = (float)d->k_min; A.k_max = (float)d->k_max;
  A.s0 = d->canonical_scale; A.lvl0 = (float)d->canonical_level; A.eps = d->eps;
  A.out = d->out;
  A.levels_out = d->levels_out;
  roi_pool_kernel<<<(unsigned)n, kPoolThreads, 0, static_cast<cudaStream_t>(stream_)>>>(A);
  OSD_LAUNCH_CHECK("roi_pool_kernel");
  return OSD_OK;
}
