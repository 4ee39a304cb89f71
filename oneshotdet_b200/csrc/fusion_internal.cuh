// Host-side pieces shared by the two translation units of the 1x1 fusion conv (fusion_conv.cu, fusion_fused.cu).
#pragma once
#include <cuda.h>

#include "osd_common.cuh"

namespace osd {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int get_encode_fn(EncodeTiledFn* out);

// bf16 [rows, cols] row-major (pitch_elems between rows) -> 2-D tensor map, box = box_cols x box_rows, 128-byte swizzle,
// out-of-range elements read as zero
int make_bf16_map(const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_cols, int box_rows,
                  CUtensorMap* map);

// (scale, shift) of GroupNorm(32, C) per (level, episode, channel) from fp64 (sum, sum of squares) per group; with
// fold_bias != nullptr the shift additionally absorbs a per-(level, episode, channel) bias that is added before the
// normalisation by the consumer:  (d + bias) * scale + shift  ==  d * scale + (bias * scale + shift)
int fusion_launch_gn_coef(int nl, int B, int C, float eps, const double* stats, const float* gn_w, const float* gn_b,
                          const float* fold_bias, float2* coef, const int32_t* hw, cudaStream_t stream);

// y = LeakyReLU(y * scale[plane] + shift[plane]) in place over [B, C, hw[l]] per level
int fusion_launch_gn_lrelu(int nl, int B, int C, float slope, float* const* y, const float2* coef, const int32_t* hw,
                           cudaStream_t stream);

struct FusionWorkspace {
  double* stats1;     // [nl, B, 32, 2]
  double* stats2;     // [nl, B, 32, 2]
  float* bias_eff;    // [nl, B, 2C]
  float2* coef1;      // [nl, B, 2C]
  float2* coef2;      // [nl, B, C]
  void* xb;           // bf16 copy of the features, per level [B*C, pitch_l]
  // Gram statistics of pass A (C >= 128): one C x C fp32 matrix and one C-vector per (CTA, plane) segment
  float* gseg;
  float* rowsum;
  int* seg_plane;
  int max_segments;
};

// pitch (elements) of level rows in the bf16 copy: hw rounded up to 8 (16-byte rows for TMA)
inline int64_t fusion_xb_pitch(int64_t hw) { return (hw + 7) / 8 * 8; }

// the whole module: pass A (conv1 statistics + bf16 copy), pass B (conv1 -> GN1 -> LeakyReLU -> conv2, fused), GN2 pass
int fusion_full_forward(const osd_fusion_desc* d, const FusionWorkspace& ws, cudaStream_t stream);

}  // namespace osd
