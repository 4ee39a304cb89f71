// Support-embedding producer: the step right before the matching path.
//   * ROIAlign with a 1x1 output over the whole support image, one ROI per support, per FPN level --
//     `SuppAlignLayer` (maskrcnn_benchmark/modeling/detector/generalized_rcnn.py:20-52, call :305) on top of
//     RoIAlignForward (csrc/cuda/ROIAlign_cuda.cu:65-122, CPU twin csrc/cpu/ROIAlign_cpu.cpp:113-214);
//   * or `nn.AdaptiveAvgPool2d((1, 1))` (generalized_rcnn.py:94, :303).
// Output [B*S, C] per level is exactly what osd_match_forward / osd_fusion_forward take as `supp`.
// One launch for all levels; a warp per (support, channel) plane for the average pool, a thread per
// (support, channel) for ROIAlign (sampling_ratio^2 bilinear taps).  Every operation is an explicit _rn intrinsic in
// the order of ROIAlign_cpu.cpp, so fp32 results are bit-identical to the reference CPU operator.
#include "osd_common.cuh"

namespace osd {
namespace {

struct PoolArgs {
  int nl, N, C, mode, sampling;
  const float* in[OSD_MAX_LEVELS];   // [N, C, H, W]
  float* out[OSD_MAX_LEVELS];        // [N, C]
  int H[OSD_MAX_LEVELS], W[OSD_MAX_LEVELS];
  float scale[OSD_MAX_LEVELS];
  const float* rois;                 // [N, 4] x1, y1, x2, y2 in image coordinates (mode ROIALIGN)
};

// bilinear tap of ROIAlign_cpu.cpp:30-108 (pre_calc_for_bilinear_interpolate) for one (y, x)
__device__ __forceinline__ float bilinear_tap(const float* __restrict__ plane, int height, int width, float y, float x) {
  if (y < -1.0f || y > (float)height || x < -1.0f || x > (float)width) return 0.f;   // contributes w = 0
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= height - 1) {
    y_high = y_low = height - 1;
    y = (float)y_low;
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= width - 1) {
    x_high = x_low = width - 1;
    x = (float)x_low;
  } else {
    x_high = x_low + 1;
  }
  const float ly = __fsub_rn(y, (float)y_low), lx = __fsub_rn(x, (float)x_low);
  const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
  const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx), w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
  const float v1 = plane[y_low * width + x_low], v2 = plane[y_low * width + x_high];
  const float v3 = plane[y_high * width + x_low], v4 = plane[y_high * width + x_high];
  // ROIAlign_cpu.cpp:199-202: w1*v1 + w2*v2 + w3*v3 + w4*v4, left to right
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1), __fmul_rn(w2, v2)), __fmul_rn(w3, v3)), __fmul_rn(w4, v4));
}

__global__ void __launch_bounds__(256) support_roialign_kernel(PoolArgs A) {
  const int l = blockIdx.z;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // n * C + c
  if (idx >= A.N * A.C) return;
  const int n = idx / A.C;
  const int H = A.H[l], W = A.W[l];
  const float s = A.scale[l];
  const float* roi = A.rois + 4 * n;
  // ROIAlign_cpu.cpp:147-165 with pooled_height = pooled_width = 1
  const float roi_start_w = __fmul_rn(roi[0], s), roi_start_h = __fmul_rn(roi[1], s);
  const float roi_end_w = __fmul_rn(roi[2], s), roi_end_h = __fmul_rn(roi[3], s);
  const float roi_width = fmaxf(__fsub_rn(roi_end_w, roi_start_w), 1.0f);
  const float roi_height = fmaxf(__fsub_rn(roi_end_h, roi_start_h), 1.0f);
  const float bin_h = roi_height, bin_w = roi_width;   // / 1
  const int gh = A.sampling > 0 ? A.sampling : (int)ceilf(roi_height);
  const int gw = A.sampling > 0 ? A.sampling : (int)ceilf(roi_width);
  const float count = (float)(gh * gw);
  const float* plane = A.in[l] + (size_t)idx * H * W;
  float acc = 0.f;
  for (int iy = 0; iy < gh; ++iy) {
    // roi_start_h + ph*bin_size_h + (iy + .5f) * bin_size_h / grid_h   with ph = 0
    const float yy = __fadd_rn(__fadd_rn(roi_start_h, __fmul_rn(0.f, bin_h)),
                               __fdiv_rn(__fmul_rn((float)iy + 0.5f, bin_h), (float)gh));
    for (int ix = 0; ix < gw; ++ix) {
      const float xx = __fadd_rn(__fadd_rn(roi_start_w, __fmul_rn(0.f, bin_w)),
                                 __fdiv_rn(__fmul_rn((float)ix + 0.5f, bin_w), (float)gw));
      acc = __fadd_rn(acc, bilinear_tap(plane, H, W, yy, xx));
    }
  }
  A.out[l][idx] = __fdiv_rn(acc, count);
}

// AdaptiveAvgPool2d((1,1)): one warp per plane; fp32 tree sum, tolerance-checked (ATen's own summation order is an
// implementation detail of its vectorised reduction)
__global__ void __launch_bounds__(256) support_avgpool_kernel(PoolArgs A) {
  const int l = blockIdx.z;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= A.N * A.C) return;
  const int hw = A.H[l] * A.W[l];
  const float* plane = A.in[l] + (size_t)warp * hw;
  float acc = 0.f;
  for (int i = lane; i < hw; i += 32) acc += plane[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) A.out[l][warp] = acc / (float)hw;
}

}  // namespace
}  // namespace osd

extern "C" int osd_support_pool(const osd_support_pool_desc* d, void* stream_) {
  using namespace osd;
  OSD_REQUIRE(d != nullptr, "osd_support_pool: desc is null");
  OSD_REQUIRE(d->num_levels >= 1 && d->num_levels <= OSD_MAX_LEVELS, "osd_support_pool: num_levels %d out of range", d->num_levels);
  OSD_REQUIRE(d->num_supports >= 0 && d->channels >= 1, "osd_support_pool: bad sizes");
  OSD_REQUIRE(d->mode == OSD_POOL_ROIALIGN || d->mode == OSD_POOL_AVG, "osd_support_pool: unknown mode %d", d->mode);
  if (d->num_supports == 0) return OSD_OK;
  OSD_REQUIRE(d->mode != OSD_POOL_ROIALIGN || d->rois != nullptr, "osd_support_pool: rois is null");
  PoolArgs A{};
  A.nl = d->num_levels; A.N = d->num_supports; A.C = d->channels; A.mode = d->mode; A.sampling = d->sampling_ratio;
  A.rois = d->rois;
  for (int l = 0; l < d->num_levels; ++l) {
    OSD_REQUIRE(d->feat[l] && d->out[l], "osd_support_pool: null pointer at level %d", l);
    OSD_REQUIRE(d->height[l] >= 1 && d->width[l] >= 1, "osd_support_pool: empty level %d", l);
    A.in[l] = static_cast<const float*>(d->feat[l]);
    A.out[l] = static_cast<float*>(d->out[l]);
    A.H[l] = d->height[l]; A.W[l] = d->width[l]; A.scale[l] = d->spatial_scale[l];
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t planes = (int64_t)d->num_supports * d->channels;
  OSD_REQUIRE(planes < (1ll << 26), "osd_support_pool: too many planes");
  if (d->mode == OSD_POOL_ROIALIGN) {
    dim3 grid((unsigned)ceil_div(planes, 256), 1, (unsigned)d->num_levels);
    support_roialign_kernel<<<grid, 256, 0, stream>>>(A);
    OSD_LAUNCH_CHECK("support_roialign_kernel");
  } else {
    dim3 grid((unsigned)ceil_div(planes * 32, 256), 1, (unsigned)d->num_levels);
    support_avgpool_kernel<<<grid, 256, 0, stream>>>(A);
    OSD_LAUNCH_CHECK("support_avgpool_kernel");
  }
  return OSD_OK;
}
