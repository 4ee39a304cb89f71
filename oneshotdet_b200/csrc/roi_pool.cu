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
#include <cuda_bf16.h>

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
  __nv_bfloat16* out_bf16;   // optional [B*R, P*P, C]: the K-major rows the dense head's GEMMs read (channels-last path)
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

// ---- channels-last path -----------------------------------------------------------------------------------------
// NCHW taps of one ROI touch 256 channel planes with ~100-byte row segments each: every warp load splinters into 10+
// sectors.  With the maps transposed once per batch to [B, H*W, C] (a 2 x 367 MB streaming pass at config size) a tap is
// one contiguous C-vector: 64 threads x LDG.128 = 1 KB, fully coalesced, and the hi/lo taps of neighbouring samples hit
// in L1.  Results are staged in shared memory in the reference's [C][P][P] order and written as one linear run.
constexpr int kTr = 32;

struct TransposeArgs {
  int nl, B, C;
  const float* in[OSD_MAX_LEVELS];
  float* out[OSD_MAX_LEVELS];
  int HW[OSD_MAX_LEVELS];
  int tile0[OSD_MAX_LEVELS + 1];   // first blockIdx.x of each level (tiles of 32 pixels)
};

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(TransposeArgs A) {
  __shared__ float tile[kTr][kTr + 1];
  int l = 0;
  while (l + 1 < A.nl && (int)blockIdx.x >= A.tile0[l + 1]) ++l;
  const int p0 = ((int)blockIdx.x - A.tile0[l]) * kTr, c0 = blockIdx.y * kTr, b = blockIdx.z;
  const int HW = A.HW[l];
  const float* in = A.in[l] + (size_t)b * A.C * HW;
  float* out = A.out[l] + (size_t)b * HW * A.C;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int k = 0; k < kTr; k += 8) {
    const int c = c0 + ty + k, p = p0 + tx;
    if (c < A.C && p < HW) tile[ty + k][tx] = in[(size_t)c * HW + p];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTr; k += 8) {
    const int p = p0 + ty + k, c = c0 + tx;
    if (c < A.C && p < HW) out[(size_t)p * A.C + c] = tile[tx][ty + k];
  }
}

__global__ void __launch_bounds__(kPoolThreads) roi_pool_nhwc_kernel(RoiPoolArgs A) {
  extern __shared__ float stage[];   // [C * P * P] in output order
  __shared__ AxisTap ytab[kTab], xtab[kTab];
  const int roi = blockIdx.x;
  const int b = roi / A.R, r = roi - b * A.R;
  const int tid = threadIdx.x;
  const int PP = A.P * A.P, per_roi = A.C * PP;
  float* out = A.out ? A.out + (size_t)roi * per_roi : nullptr;
  __nv_bfloat16* outb = A.out_bf16 ? A.out_bf16 + (size_t)roi * per_roi : nullptr;
  const int n_valid = A.roi_count ? min(max(A.roi_count[b], 0), A.R) : A.R;
  if (r >= n_valid) {
    for (int e = tid; e < per_roi; e += kPoolThreads) {
      if (out) out[e] = 0.f;
      if (outb) outb[e] = __float2bfloat16_rn(0.f);
    }
    if (tid == 0 && A.levels_out) A.levels_out[roi] = -1;
    return;
  }
  const float4 box = A.rois[roi];
  const float area = __fmul_rn(__fadd_rn(__fsub_rn(box.z, box.x), 1.0f), __fadd_rn(__fsub_rn(box.w, box.y), 1.0f));
  const float s = __fsqrt_rn(area);
  float lv = floorf(__fadd_rn(A.lvl0, log2f(__fadd_rn(__fdiv_rn(s, A.s0), A.eps))));
  lv = fminf(fmaxf(lv, A.k_min), A.k_max);
  const int l = A.nl == 1 ? 0 : (int)(lv - A.k_min);
  if (tid == 0 && A.levels_out) A.levels_out[roi] = l;
  const int H = A.H[l], W = A.W[l];
  const float sc = A.scale[l];
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
  // feat[l] here is the transposed copy [B, H*W, C]
  const float4* fmap = reinterpret_cast<const float4*>(A.feat[l] + (size_t)b * H * W * A.C);
  const int C4 = A.C >> 2;
  // work items: (bin, channel quad); consecutive threads take consecutive quads of one bin
  for (int item = tid; item < PP * C4; item += kPoolThreads) {
    const int bin = item / C4, q = item - bin * C4;
    const int ph = bin / A.P, pw = bin - ph * A.P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int iy = 0; iy < gh; ++iy) {
      const AxisTap ty = tabbed ? ytab[ph * gh + iy] : axis_tap(roi_start_h, bin_h, ph, iy, gh, H);
      for (int ix = 0; ix < gw; ++ix) {
        const AxisTap tx = tabbed ? xtab[pw * gw + ix] : axis_tap(roi_start_w, bin_w, pw, ix, gw, W);
        if (!(ty.valid && tx.valid)) continue;
        const float w1 = __fmul_rn(ty.h, tx.h), w2 = __fmul_rn(ty.h, tx.l), w3 = __fmul_rn(ty.l, tx.h), w4 = __fmul_rn(ty.l, tx.l);
        const float4 v1 = __ldg(fmap + (size_t)(ty.lo * W + tx.lo) * C4 + q);
        const float4 v2 = __ldg(fmap + (size_t)(ty.lo * W + tx.hi) * C4 + q);
        const float4 v3 = __ldg(fmap + (size_t)(ty.hi * W + tx.lo) * C4 + q);
        const float4 v4 = __ldg(fmap + (size_t)(ty.hi * W + tx.hi) * C4 + q);
#define OSD_TAP(f) acc.f = __fadd_rn(acc.f, __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w1, v1.f), __fmul_rn(w2, v2.f)), \
                                                                 __fmul_rn(w3, v3.f)), __fmul_rn(w4, v4.f)))
        OSD_TAP(x); OSD_TAP(y); OSD_TAP(z); OSD_TAP(w);
#undef OSD_TAP
      }
    }
    const int c = q << 2;
    const float r0 = __fdiv_rn(acc.x, count), r1 = __fdiv_rn(acc.y, count), r2 = __fdiv_rn(acc.z, count), r3 = __fdiv_rn(acc.w, count);
    if (outb) {   // [bin][channel] rows: consecutive threads write consecutive 8-byte quads of one row
      const __nv_bfloat162 lo = __floats2bfloat162_rn(r0, r1), hi = __floats2bfloat162_rn(r2, r3);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(outb + (size_t)bin * A.C + c) = pk;
    }
    if (out) {
      stage[(c + 0) * PP + bin] = r0;
      stage[(c + 1) * PP + bin] = r1;
      stage[(c + 2) * PP + bin] = r2;
      stage[(c + 3) * PP + bin] = r3;
    }
  }
  if (!out) return;
  __syncthreads();
  if ((per_roi & 3) == 0) {   // out + roi * per_roi is then 16-byte aligned (cudaMalloc / torch allocations are)
    float4* o4 = reinterpret_cast<float4*>(out);
    const float4* s4 = reinterpret_cast<const float4*>(stage);
    for (int e = tid; e < (per_roi >> 2); e += kPoolThreads) __stcs(o4 + e, s4[e]);
  } else {
    for (int e = tid; e < per_roi; e += kPoolThreads) __stcs(out + e, stage[e]);
  }
}

size_t transposed_bytes(const osd_roi_pool_desc* d, size_t* off) {
  size_t total = 0;
  for (int l = 0; l < d->num_levels; ++l) {
    if (off) off[l] = total;
    total += align_up((size_t)d->batch * d->channels * d->height[l] * d->width[l] * sizeof(float), 256);
  }
  return total;
}

}  // namespace
}  // namespace osd

extern "C" int osd_roi_pool_workspace_bytes(const osd_roi_pool_desc* d, size_t* bytes) {
  using namespace osd;
  OSD_REQUIRE(d != nullptr && bytes != nullptr, "osd_roi_pool_workspace_bytes: null argument");
  OSD_REQUIRE(d->num_levels >= 1 && d->num_levels <= OSD_MAX_LEVELS, "osd_roi_pool: num_levels %d out of range", d->num_levels);
  *bytes = transposed_bytes(d, nullptr);
  return OSD_OK;
}

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
  OSD_REQUIRE(d->rois != nullptr && (d->out != nullptr || d->out_nhwc_bf16 != nullptr), "osd_roi_pool: null rois / out");
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
  A.out_bf16 = static_cast<__nv_bfloat16*>(d->out_nhwc_bf16);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t stage_bytes = (size_t)d->channels * d->pooled_size * d->pooled_size * sizeof(float);
  const bool channels_last = d->workspace != nullptr && (d->channels & 3) == 0 && stage_bytes <= 200 * 1024 &&
                             (reinterpret_cast<uintptr_t>(d->out) & 15) == 0;
  if (!channels_last) {  // direct NCHW taps (any C, no workspace)
    OSD_REQUIRE(d->out_nhwc_bf16 == nullptr && d->out != nullptr,
                "osd_roi_pool: the bf16 [roi, pixel, channel] output needs the channels-last path (workspace, C %% 4 == 0)");
    roi_pool_kernel<<<(unsigned)n, kPoolThreads, 0, stream>>>(A);
    OSD_LAUNCH_CHECK("roi_pool_kernel");
    return OSD_OK;
  }
  size_t off[OSD_MAX_LEVELS];
  const size_t need = transposed_bytes(d, off);
  if (need > d->workspace_bytes) {
    set_error("osd_roi_pool: workspace of %zu bytes needed, %zu given", need, d->workspace_bytes);
    return OSD_ERR_WORKSPACE;
  }
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(d->workspace) & 255) == 0, "osd_roi_pool: workspace must be 256-byte aligned");
  TransposeArgs T{};
  T.nl = d->num_levels; T.B = d->batch; T.C = d->channels;
  int tiles = 0;
  for (int l = 0; l < d->num_levels; ++l) {
    T.in[l] = A.feat[l];
    T.out[l] = reinterpret_cast<float*>(static_cast<char*>(d->workspace) + off[l]);
    T.HW[l] = d->height[l] * d->width[l];
    T.tile0[l] = tiles;
    tiles += (int)ceil_div(T.HW[l], kTr);
    A.feat[l] = T.out[l];
  }
  T.tile0[d->num_levels] = tiles;
  OSD_REQUIRE(d->batch <= 65535 && ceil_div(d->channels, kTr) <= 65535, "osd_roi_pool: batch / channels out of range");
  dim3 tgrid((unsigned)tiles, (unsigned)ceil_div(d->channels, kTr), (unsigned)d->batch);
  nchw_to_nhwc_kernel<<<tgrid, 256, 0, stream>>>(T);
  OSD_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  {
    int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(roi_pool_nhwc_kernel), 200 * 1024);
    if (rc2 != OSD_OK) return rc2;
  }
  // the fp32 [C][P][P] result is staged in shared memory; with only the bf16 rows requested nothing is staged and more
  // CTAs fit an SM (the kernel is bound by the latency of its bilinear taps): 2.26 -> 1.9 ms for 16 x 2000 ROIs.  (Unrolling
  // the 2 x 2 sample grid so that all 16 tap loads of an item are in flight changed nothing: 1.905 vs 1.877 ms.)
  roi_pool_nhwc_kernel<<<(unsigned)n, kPoolThreads, d->out ? stage_bytes : 0, stream>>>(A);
  OSD_LAUNCH_CHECK("roi_pool_nhwc_kernel");
  return OSD_OK;
}
