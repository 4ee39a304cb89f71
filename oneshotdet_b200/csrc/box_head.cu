// Second-stage dense head on tcgen05 (sm_100a), hand-written: SURVEY.md section 8(f) row 2c.
//
// Reference: ROIBoxHead.forward between the pooler and the post-processor for comparison_method 'concat', one support,
// no negative support, LINEAR_FUSION off (maskrcnn_benchmark/modeling/roi_heads/box_head/box_head.py:118-157):
//     x  = cat((pooled, support.expand_as(pooled)), 1)                               [B*R, 2C, 7, 7]      (:147)
//     x  = compress_dim_conv(x)   Conv1x1(2C,2C) GN(32) LReLU Conv1x1(2C,C) GN(32) LReLU                  (:43-54, :149)
//     x  = feature_aggreg(x)      Conv3x3(C, C/2, pad 1) GN(32) LReLU                                     (:62-67, :151)
//     x  = relu(fc6(x.view(N, -1)));  x = relu(fc7(x))                                                    (:152-154)
//     class_logits, box_regression = cls_score(x), bbox_pred(x)       (roi_box_predictors.py:80-84)
//
// Unlike the FPN maps of the first stage, GroupNorm here normalises over the 49 pixels of ONE ROI: a 128-row MMA tile
// holds two ROIs (rows 0-48 and 64-112; the other rows are padding), so every normalisation is local to the tile's
// epilogue and each layer is ONE GEMM launch with the whole GN + LeakyReLU in its epilogue -- no statistics pass.
//
// One persistent, warp-specialised GEMM kernel serves all six layers:
//   D[rows, N] = A[rows, K] . W[N, K]^T     bf16 x bf16 -> fp32 in tensor memory
//   * A and W tiles come through TMA (128-byte swizzle, K-major) into a 4-9-stage ring; the A tile of the ROI layers is
//     two boxes of 49 rows.  conv1's concat is never materialised: the K range [0, C) is read from the ROI's rows and
//     [C, 2C) from the support rows of the ROI's episode (a second tensor map).  The 3x3 convolution is an implicit
//     GEMM: its A tile for tap (dy, dx) is a box of the 4-D map [roi, y, x, c] at offset (dy-1, dx-1) -- the TMA unit
//     zero-fills the halo, so there is no im2col buffer and no padding pass.
//   * one elected thread issues tcgen05.mma (M = 128, N = the layer's tile, K = 16) into one of two TMEM accumulators;
//     8 epilogue warps (TMEM lane quadrant x column half) drain the other one: bias, GroupNorm statistics over the
//     ROI's 49 rows (registers -> warp transpose-reduce -> one shared-memory exchange between the two warps that share
//     an ROI), normalise + LeakyReLU on a second read of tensor memory, bf16 rows straight to global memory.
//   * the ROI layers are bound by the L2 -> SM operand stream (every tile streams the layer's whole weight matrix): the
//     layers with N <= 128 therefore run 256 x 128 CTA tiles (two M tiles against one weight stage, two accumulators per
//     TMEM buffer); fc6 (K = 6272) runs at the tensor pipe's peak.
//   * activations between layers are bf16 [roi, pixel, channel]; the host walks the ROIs in chunks (1184 = 4 full waves of
//     two-ROI tiles) so that a chunk's intermediates stay in the 126 MB L2 between consecutive layers; the fully
//     connected layers run once over all ROIs.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "fusion_internal.cuh"
#include "osd_common.cuh"
#include "osd_tc.cuh"

namespace osd {
namespace {
using namespace tc;

constexpr int kP = 7, kPix = kP * kP;
constexpr int kBM = 128;                       // rows per tile (UMMA M)
constexpr int kBK = 64;                        // K elements per stage (one 128-byte swizzle row of bf16)
constexpr uint32_t kABytes = kBM * kBK * 2;    // 16 KB
constexpr uint32_t kRingBytes = 4 * (kABytes + 256 * kBK * 2);   // 192 KB of operand stages
constexpr int kMaxRing = 10;
// ring depth for an N tile of BN columns: a stage is the A tile (16 KB) + BN rows of W (128 B each).  Narrow layers get a
// deeper ring -- their stages are consumed faster (4 MMAs of BN/2 clocks each) while the TMA latency stays the same
__host__ __device__ constexpr uint32_t stage_bytes(int bn, int mt) { return (uint32_t)mt * kABytes + (uint32_t)bn * 128u; }
__host__ __device__ constexpr int ring_depth(int bn, int mt) {
  return (int)(kRingBytes / stage_bytes(bn, mt)) < kMaxRing ? (int)(kRingBytes / stage_bytes(bn, mt)) : kMaxRing;
}
constexpr int kMaxEpiWarps = 16;
constexpr int kMaxParamCols = 2048;            // bias / gamma / beta of up to this many output columns are staged in smem
constexpr uint32_t kXchBytes = 2 * kMaxEpiWarps * 32 * 4;       // statistics exchange between the two warps of an ROI
constexpr uint32_t kCoefBytes = 2 * 2 * 256 * 8;                // (scale, shift) [tile parity][ROI][N tile column]
constexpr uint32_t kParamBytes = kMaxParamCols * 4 + 2 * 512 * 4;   // bias [2048], gamma [512], beta [512]
constexpr uint32_t kGemmSmem = kRingBytes + 256 /* barriers */ + kXchBytes + kCoefBytes + kParamBytes + 1024 /* align */;

enum { A_PLAIN = 0, A_ROI = 1, A_ROI_3X3 = 2 };

struct GemmArgs {
  int M;               // A_PLAIN: rows; ROI modes: number of ROIs (49 rows each)
  int N, K;
  int a_mode;
  int k_split;         // A_ROI: K elements served by map A; the rest come from map A2 (support rows of the ROI's episode)
  int roi0;            // global index of this chunk's first ROI
  int rois_per_image;
  int tap_c;           // A_ROI_3X3: channels per tap (K = 9 * tap_c)
  const float* bias;   // [N]
  const float* gamma;  // [N] GroupNorm weight (EPI 1)
  const float* beta;   // [N] GroupNorm bias (EPI 1)
  float eps, slope;
  int relu;            // EPI 0
  int out_f32;         // EPI 0: fp32 instead of bf16 rows
  void* out;           // rows x ldo
  int ldo;
  void* out2;          // EPI 0, fp32: columns >= n_split go to out2[row * ldo2 + (col - n_split)]
  int ldo2, n_split;
};

// conflict-free: value L of every lane summed over the warp ends up in lane L (31 shuffles for 32 values)
__device__ __forceinline__ void transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
}

__device__ __forceinline__ bool elect_one_lane() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int LD>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&r)[32]) {
  if (LD == 32) {
    tmem_ld32(taddr, r);
  } else {
    uint32_t t[16];
    tmem_ld16(taddr, t);
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = t[i];
  }
}

// store n (8 or 16 or 32) consecutive bf16 values given as packed pairs
template <int NPAIR>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* dst, const uint32_t (&p)[16]) {
#pragma unroll
  for (int v = 0; v < NPAIR / 4; ++v)
    *reinterpret_cast<uint4*>(dst + 8 * v) = make_uint4(p[4 * v], p[4 * v + 1], p[4 * v + 2], p[4 * v + 3]);
}

// EPI 0: bias (+ ReLU) -> bf16 / fp32 rows.  EPI 1: bias + GroupNorm(32 groups of GS channels, over the 49 rows of the
// ROI) + LeakyReLU -> bf16 rows.  The N tile is split into NH column parts of CPW columns; 4 * NH epilogue warps (TMEM
// lane quadrant x column part) drain an accumulator -- 16 warps on the wide layers: with 8 the epilogue issued at 0.3
// instructions per cycle and scheduler (two warps each, a chain of TMEM loads, shuffles and shared-memory reads) and was
// what bound conv1 / conv2.
// CL = 2: launched as clusters of two CTAs that work on two M tiles of the SAME N tile in lock step; every W stage is
// fetched from L2 once per cluster (each CTA loads half of its rows and the TMA unit multicasts them into both CTAs'
// rings).  The ROI layers stream the whole weight matrix per M tile and are bound by that L2 -> SM stream.
// MT = 2 (N tile <= 128): a CTA tile is TWO 128-row M tiles against one weight stage (two accumulators of 128 columns per
// TMEM buffer) -- a 256 x 128 tile moves 28 % fewer operand bytes per FLOP than 128 x 128.
template <int EPI, int CPW, int GS, int NH, int CL, int MT>
__global__ void __launch_bounds__(32 * (4 * NH + 2), 1)
roi_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
                const __grid_constant__ CUtensorMap mapW, const GemmArgs G) {
  constexpr int BN = NH * CPW;
  constexpr int kEpiWarps = 4 * NH;
  static_assert(CL == 1 || CL == 2, "cluster size");
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
  const int cta = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // tile-loop index of this CTA (pair)
  const int ncta = CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int LD = CPW >= 32 ? 32 : 16;
  static_assert(BN <= 256 && BN % 16 == 0, "N tile");
  static_assert(MT == 1 || (MT == 2 && BN <= 128), "two M tiles per CTA tile need two accumulators per TMEM buffer");
  constexpr int kRing = ring_depth(BN, MT);
  constexpr uint32_t kStage = stage_bytes(BN, MT);
  constexpr uint32_t kAStage = (uint32_t)MT * kABytes;
  static_assert(kStage % 1024 == 0 && kRing >= 2 && 8 * (2 * kRing + 4) + 4 <= 256, "ring layout");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar0 = base + kRingBytes;
  const uint32_t b_full = bar0, b_empty = bar0 + 8u * kRing, b_accf = bar0 + 16u * kRing, b_acce = b_accf + 16u;
  const uint32_t tmem_slot = b_acce + 16u;
  float* xch = reinterpret_cast<float*>(gen + kRingBytes + 256);   // [2][kEpiWarps][32]
  float2* coef = reinterpret_cast<float2*>(gen + kRingBytes + 256 + kXchBytes);
  float* sBias = reinterpret_cast<float*>(gen + kRingBytes + 256 + kXchBytes + kCoefBytes);
  float* sGamma = sBias + kMaxParamCols;
  float* sBeta = sGamma + 512;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int m_tiles = G.a_mode == A_PLAIN ? (G.M + kBM * MT - 1) / (kBM * MT) : (G.M + 2 * MT - 1) / (2 * MT);
  const int n_tiles = (G.N + BN - 1) / BN;
  // pairs: tile t of the loop = (pair of M tiles t / n_tiles, N tile t % n_tiles); rank r takes M tile 2 * pair + r (past
  // the last M tile: a dummy whose loads are out of range = zero-filled and whose rows are all invalid)
  const int total = ((m_tiles + CL - 1) / CL) * n_tiles;
  const int num_k = (G.K + kBK - 1) / kBK;

  if (CL == 2) cluster_sync_all();
  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapW);
    if (G.a_mode == A_ROI && G.k_split < G.K) tma_prefetch_desc(&mapA2);
    for (int s = 0; s < kRing; ++s) {
      mbar_init(b_full + 8u * s, 1);
      mbar_init(b_empty + 8u * s, CL);      // pairs: the MMA commits of both CTAs (the stage is refilled in both)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(b_accf + 8u * a, 1);
      mbar_init(b_acce + 8u * a, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  if (CL == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (warp == kEpiWarps) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t c = 0;
      const uint32_t a_bytes = (uint32_t)MT * (G.a_mode == A_PLAIN ? kABytes : 2u * kPix * 128u);
      for (int tile = cta; tile < total; tile += ncta) {
        const int mt = (tile / n_tiles) * CL + (int)crank, nt = tile % n_tiles;
        for (int k = 0; k < num_k; ++k, ++c) {
          const uint32_t s = c % kRing, ph = (c / kRing) & 1u;
          mbar_wait(b_empty + 8u * s, ph ^ 1u);
          const uint32_t sA0 = base + s * kStage, sB = sA0 + kAStage, fb = b_full + 8u * s;
          mbar_expect_tx(fb, a_bytes + (uint32_t)BN * 128u);
          const int k0 = k * kBK;
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
            const uint32_t sA = sA0 + (uint32_t)mi * kABytes;
            const int ms = mt * MT + mi;               // 128-row M tile
            if (G.a_mode == A_PLAIN) {
              tma_load_2d(sA, &mapA, k0, ms * kBM, fb);
            } else {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int roi = ms * 2 + j;            // past the last ROI: out of range -> zero-filled, bytes still counted
                const uint32_t dst = sA + (uint32_t)j * (64u * 128u);
                if (G.a_mode == A_ROI) {
                  if (k0 < G.k_split) tma_load_2d(dst, &mapA, k0, roi * kPix, fb);
                  else tma_load_2d(dst, &mapA2, k0 - G.k_split, ((G.roi0 + roi) / G.rois_per_image) * kPix, fb);
                } else {
                  const int tap = k0 / G.tap_c, kc = k0 - tap * G.tap_c;
                  tma_load_4d(dst, &mapA, kc, tap % 3 - 1, tap / 3 - 1, roi, fb);
                }
              }
            }
          }
          if (CL == 2)
            tma_load_2d_mc(sB + crank * (uint32_t)(BN / 2) * 128u, &mapW, k0, nt * BN + (int)crank * (BN / 2), fb, (uint16_t)3);
          else
            tma_load_2d(sB, &mapW, k0, nt * BN, fb);
        }
      }
      if (CL == 2) {   // the peer's last commits must have landed on this CTA's barriers before it may exit
        for (uint32_t s = 0; s < (uint32_t)kRing; ++s) {
          if (c <= s) continue;
          const uint32_t uses = (c - s + kRing - 1) / kRing;
          mbar_wait(b_empty + 8u * s, (uses - 1) & 1u);
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_ex(kBM, BN, /*a_mn=*/0, /*b_mn=*/0);
    uint32_t c = 0, it = 0;
    for (int tile = cta; tile < total; tile += ncta, ++it) {
      const uint32_t acc = it & 1u;
      mbar_wait(b_acce + 8u * acc, ((it >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 256u;
      for (int k = 0; k < num_k; ++k, ++c) {
        const uint32_t s = c % kRing, ph = (c / kRing) & 1u;
        mbar_wait(b_full + 8u * s, ph);
        tc_fence_after();
        if (elect_one_lane()) {
          const uint64_t bd = make_smem_desc(base + s * kStage + kAStage, 16u, 1024u);
#pragma unroll
          for (int mi = 0; mi < MT; ++mi) {
            const uint64_t ad = make_smem_desc(base + s * kStage + (uint32_t)mi * kABytes, 16u, 1024u);
#pragma unroll
            for (int k16 = 0; k16 < kBK / 16; ++k16)
              umma_bf16(d + (uint32_t)mi * 128u, ad + (uint64_t)(k16 * 2), bd + (uint64_t)(k16 * 2), idesc, (k | k16) != 0 ? 1u : 0u);
          }
          if (CL == 2) umma_commit_mc(b_empty + 8u * s, (uint16_t)3);
          else umma_commit(b_empty + 8u * s);
        }
        __syncwarp();
      }
      if (elect_one_lane()) umma_commit(b_accf + 8u * acc);
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 0-7; thread = tile row) =====================
    // Per-column parameters live in shared memory (a warp reads the same column in all lanes: one broadcast wavefront,
    // four columns per LDS.128).  The global loads they replace (four per element) were what bound the first version:
    // 2500 instructions per thread and tile, LSU queue full.
    const int q = warp & 3, hh = warp >> 2;
    const int etid = threadIdx.x;                       // 0 .. 255
    const bool params_in_smem = G.N <= kMaxParamCols;
    if (params_in_smem) {
      for (int c = etid; c < G.N; c += 32 * kEpiWarps) {
        sBias[c] = __ldg(G.bias + c);
        if (EPI == 1) {
          sGamma[c] = __ldg(G.gamma + c);
          sBeta[c] = __ldg(G.beta + c);
        }
      }
    }
    named_bar_sync(9, 32 * kEpiWarps);
    uint32_t it = 0;
    for (int tile = cta; tile < total; tile += ncta, ++it) {
      const int mt = (tile / n_tiles) * CL + (int)crank, nt = tile % n_tiles;
      const uint32_t acc = it & 1u;
      const int col0 = nt * BN + hh * CPW;
      mbar_wait(b_accf + 8u * acc, (it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int mi = 0; mi < MT; ++mi) {
      const int ms = mt * MT + mi;                       // this accumulator's 128-row M tile
      const uint32_t par = (it * MT + (uint32_t)mi) & 1u; // parity of the exchange / coefficient buffers
      const bool last_sub = mi == MT - 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u + (uint32_t)mi * 128u + (uint32_t)(hh * CPW);
      if constexpr (EPI == 1) {
        constexpr int NG = CPW / GS;
        static_assert(NG <= 16, "a warp keeps (sum, sum of squares) of at most 16 groups");
        const int row = q * 32 + lane, rl = row & 63;
        const int rloc = row >> 6;                       // which of the tile's two ROIs
        const int roi = ms * 2 + rloc;
        const bool valid = rl < kPix && roi < G.M;
        // ---- pass 1: (sum, sum of squares) of y = d + bias per group over this thread's row
        float acc_s[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc_s[i] = 0.f;
#pragma unroll
        for (int cb = 0; cb < CPW; cb += LD) {
          uint32_t r[32];
          tmem_ld_cols<LD>(taddr + (uint32_t)cb, r);
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(sBias + col0 + cb);
#pragma unroll
          for (int i4 = 0; i4 < LD / 4; ++i4) {
            const float4 bv = bp[i4];
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * i4 + e;
              const int gi = (cb + i) / GS;
              const float y = __uint_as_float(r[i]) + bb[e];
              acc_s[2 * gi] += y;
              acc_s[2 * gi + 1] = fmaf(y, y, acc_s[2 * gi + 1]);
            }
          }
        }
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc_s[i] = 0.f;
        }
        transpose_reduce32(acc_s, lane);
        // rows of one ROI sit in two warps (q even: rows 0-31, q odd: rows 32-48): exchange through shared memory
        float* mine = xch + (par * kEpiWarps + warp) * 32;
        const float* theirs = xch + (par * kEpiWarps + (warp ^ 1)) * 32;
        mine[lane] = acc_s[0];
        named_bar_sync(1 + (warp >> 1), 64);
        const float tot = acc_s[0] + theirs[lane];
        const float other = __shfl_xor_sync(0xffffffffu, tot, 1);
        const float sum = (lane & 1) ? other : tot, sq = (lane & 1) ? tot : other;
        constexpr float inv_n = 1.0f / (float)(kPix * GS);
        const float mean = sum * inv_n;
        const float var = fmaxf(fmaf(-mean, mean, sq * inv_n), 0.f);
        const float rstd = 1.0f / sqrtf(var + G.eps);          // lanes 2g, 2g+1: statistics of group g
        // ---- coefficient table of this (ROI, column half): y_out = d * scale + shift, computed once per column by the
        //      64 threads of the warp pair instead of once per element by every row
        float2* ctab = coef + ((par * 2 + rloc) * NH + hh) * CPW;          // [parity][roi][part][CPW columns]
#pragma unroll
        for (int u = 0; u < (CPW + 63) / 64; ++u) {
          const int cl = (q & 1) * 32 + lane + 64 * u;          // column inside this warp pair's half
          const int cc = cl < CPW ? cl : 0;
          const int gi = cc / GS;
          const float rs = __shfl_sync(0xffffffffu, rstd, 2 * gi);
          const float mn = __shfl_sync(0xffffffffu, mean, 2 * gi);
          if (cl < CPW) {
            const float sc = rs * sGamma[col0 + cl];
            ctab[cl] = make_float2(sc, fmaf(sBias[col0 + cl] - mn, sc, sBeta[col0 + cl]));
          }
        }
        named_bar_sync(1 + (warp >> 1), 64);
        // ---- pass 2: normalise + LeakyReLU on a second read of the accumulator, bf16 rows to global memory
        __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(G.out) + ((size_t)roi * kPix + rl) * (size_t)G.ldo + col0;
        const float slope = G.slope;
#pragma unroll
        for (int cb = 0; cb < CPW; cb += LD) {
          uint32_t r[32];
          tmem_ld_cols<LD>(taddr + (uint32_t)cb, r);
          tmem_ld_wait();
          if (cb + LD >= CPW && last_sub) {   // last read of this TMEM buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_acce + 8u * acc);
          }
          const float4* cp4 = reinterpret_cast<const float4*>(ctab + cb);   // (scale, shift) of two columns
          uint32_t p[16];
#pragma unroll
          for (int i = 0; i < LD; i += 2) {
            const float4 k = cp4[i >> 1];
            float a = fmaf(__uint_as_float(r[i]), k.x, k.y);
            float b = fmaf(__uint_as_float(r[i + 1]), k.z, k.w);
            a = a > 0.f ? a : a * slope;
            b = b > 0.f ? b : b * slope;
            p[i >> 1] = pack_bf16x2(a, b);
          }
          if (valid) store_bf16_row<LD / 2>(orow + cb, p);
        }
      } else {
        const int row = ms * kBM + q * 32 + lane;
        const bool valid = row < G.M;
#pragma unroll
        for (int cb = 0; cb < CPW; cb += LD) {
          uint32_t r[32];
          tmem_ld_cols<LD>(taddr + (uint32_t)cb, r);
          tmem_ld_wait();
          if (cb + LD >= CPW && last_sub) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(b_acce + 8u * acc);
          }
          const int c0 = col0 + cb;
          if (c0 >= G.N) continue;
          float y[LD];
#pragma unroll
          for (int i = 0; i < LD; ++i) {
            const int c = c0 + i;
            const float bv = c < G.N ? (params_in_smem ? sBias[c] : __ldg(G.bias + c)) : 0.f;
            const float v = __uint_as_float(r[i]) + bv;
            y[i] = G.relu ? fmaxf(v, 0.f) : v;
          }
          if (!valid) continue;
          if (G.out_f32) {
#pragma unroll
            for (int i = 0; i < LD; ++i) {
              const int c = c0 + i;
              if (c < G.N) {
                if (G.out2 != nullptr && c >= G.n_split) static_cast<float*>(G.out2)[(size_t)row * G.ldo2 + (c - G.n_split)] = y[i];
                else static_cast<float*>(G.out)[(size_t)row * G.ldo + c] = y[i];
              }
            }
          } else {
            __nv_bfloat16* orow = static_cast<__nv_bfloat16*>(G.out) + (size_t)row * G.ldo + c0;
            if (c0 + LD <= G.N && (G.ldo & 7) == 0) {
              uint32_t p[16];
#pragma unroll
              for (int i = 0; i < LD; i += 2) p[i >> 1] = pack_bf16x2(y[i], y[i + 1]);
              store_bf16_row<LD / 2>(orow, p);
            } else {
#pragma unroll
              for (int i = 0; i < LD; ++i)
                if (c0 + i < G.N) orow[i] = __float2bfloat16_rn(y[i]);
            }
          }
        }
      }
      }   // sub-tiles
    }
  }

  tc_fence_before();
  if (CL == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// fp32 [n, C, 49] (the Pooler's NCHW rows) -> bf16 [n, 49, C] (K-major rows for the GEMMs); one ROI per CTA
__global__ void __launch_bounds__(256) pack_roi_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C) {
  extern __shared__ __nv_bfloat16 pk_tile[];   // [49][C + 2]
  const int pitch = C + 2;
  const float* src = in + (size_t)blockIdx.x * C * kPix;
  // C * 49 floats per ROI with C a multiple of 4: the ROI's block is 16-byte aligned whenever `in` is
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (int i4 = threadIdx.x; i4 < C * kPix / 4; i4 += blockDim.x) {
    const float4 v = __ldg(src4 + i4);
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * i4 + k;
      const int c = i / kPix, p = i - c * kPix;
      pk_tile[p * pitch + c] = __float2bfloat16_rn(e[k]);
    }
  }
  __syncthreads();
  __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(out + (size_t)blockIdx.x * C * kPix);
  const int half = C / 2;
  for (int i = threadIdx.x; i < half * kPix; i += blockDim.x) {
    const int p = i / half, c2 = i - p * half;
    dst[i] = *reinterpret_cast<const __nv_bfloat162*>(pk_tile + p * pitch + 2 * c2);
  }
}

// bf16 [N, 7, 7, C] -> 4-D tensor map (c, x, y, roi), box = 64 channels x 7 x 7 x 1 ROI, 128-byte swizzle
int make_roi_map_4d(const void* base, int64_t rois, int C, CUtensorMap* map) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc != OSD_OK) return rc;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)kP, (cuuint64_t)kP, (cuuint64_t)rois};
  cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2 * kP, (cuuint64_t)C * 2 * kPix};
  cuuint32_t box[4] = {64, (cuuint32_t)kP, (cuuint32_t)kP, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D ROI map) failed with CUresult %d", (int)r);
    return OSD_ERR_CUDA;
  }
  return OSD_OK;
}

// OSD_BOX_HEAD_CLUSTER=2 launches clusters of two CTAs that share every weight stage through TMA multicast.  Measured
// (C = 256, 32 000 ROIs): 6.07 ms against 5.88 ms for single CTAs -- the weight stream is not what binds these layers.
bool head_pairs() {
  static const bool on = [] { const char* e = getenv("OSD_BOX_HEAD_CLUSTER"); return e && e[0] == '2'; }();
  return on;
}

bool head_two_m_tiles() {
  static const bool on = [] { const char* e = getenv("OSD_BOX_HEAD_MT"); return !(e && e[0] == '1'); }();
  return on;
}

template <int EPI, int CPW, int GS, int NH, int MT = 1>
int launch_gemm(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mW, const GemmArgs& G, cudaStream_t stream,
                const char* name, bool pairs) {
  constexpr int BN = NH * CPW;
  constexpr int kGemmThreads = 32 * (4 * NH + 2);
  const int m_tiles = G.a_mode == A_PLAIN ? (G.M + kBM * MT - 1) / (kBM * MT) : (G.M + 2 * MT - 1) / (2 * MT);
  const int n_tiles = (G.N + BN - 1) / BN;
  if (m_tiles * n_tiles <= 0) return OSD_OK;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3((unsigned)kGemmThreads);
  cfg.dynamicSmemBytes = kGemmSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pairs) {
    auto k = roi_gemm_kernel<EPI, CPW, GS, NH, 2, MT>;
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(k), kGemmSmem);
    if (rc != OSD_OK) return rc;
    const int total = ((m_tiles + 1) / 2) * n_tiles;
    attr[0].val.clusterDim.x = 2;
    cfg.gridDim = dim3((unsigned)(2 * std::min(total, kNumSMs / 2)));
    OSD_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mA2, mW, G));
  } else {
    auto k = roi_gemm_kernel<EPI, CPW, GS, NH, 1, MT>;
    int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(k), kGemmSmem);
    if (rc != OSD_OK) return rc;
    attr[0].val.clusterDim.x = 1;
    cfg.gridDim = dim3((unsigned)std::min(m_tiles * n_tiles, kNumSMs));
    OSD_CUDA(cudaLaunchKernelEx(&cfg, k, mA, mA2, mW, G));
  }
  OSD_LAUNCH_CHECK(name);
  timeline_mark(name, stream);
  return OSD_OK;
}

// GroupNorm(32, N) layers: N channels -> GS = N / 32 per group, CPW = min(128, N / 2) columns per epilogue warp
int launch_gn_gemm(const CUtensorMap& mA, const CUtensorMap& mA2, const CUtensorMap& mW, const GemmArgs& G, cudaStream_t stream,
                   const char* name, bool pairs) {
  switch (G.N) {
    // (16 epilogue warps -- <1, 64, 16, 4> etc. -- measured no faster than 8: the layers are bound by the L2 -> SM
    // operand stream, not by the epilogue's issue rate; section 8.1 of DESIGN.md)
    case 512: return launch_gemm<1, 128, 16, 2>(mA, mA2, mW, G, stream, name, pairs);
    case 256: return launch_gemm<1, 128, 8, 2>(mA, mA2, mW, G, stream, name, pairs);
    case 128:
      // 256 x 128 CTA tiles (two accumulators per TMEM buffer); OSD_BOX_HEAD_MT=1: 128 x 128
      if (head_two_m_tiles()) return launch_gemm<1, 64, 4, 2, 2>(mA, mA2, mW, G, stream, name, pairs);
      return launch_gemm<1, 64, 4, 2>(mA, mA2, mW, G, stream, name, pairs);
    case 64: return launch_gemm<1, 32, 2, 2>(mA, mA2, mW, G, stream, name, pairs);
    case 32: return launch_gemm<1, 16, 1, 2>(mA, mA2, mW, G, stream, name, pairs);
  }
  set_error("osd_box_head: no GroupNorm epilogue for %d channels", G.N);
  return OSD_ERR_INVALID;
}

struct HeadWorkspace {
  __nv_bfloat16* sb;   // [B*49, C]
  __nv_bfloat16* xb;   // [chunk*49, C]
  __nv_bfloat16* a1;   // [chunk*49, 2C]
  __nv_bfloat16* a2;   // [chunk*49, C]
  __nv_bfloat16* a3;   // [B*R*49, C/2] == [B*R, 49*C/2]: all ROIs, the fully connected layers run once over everything
  __nv_bfloat16* h6;   // [B*R, mlp]
  __nv_bfloat16* h7;   // [B*R, mlp]
};

int head_chunk(const osd_box_head_desc* d) {
  const int64_t n = (int64_t)d->batch * d->rois_per_image;
  // default: 592 tiles of two ROIs = 4 full waves of the 148 SMs per layer (8 for conv1's two N tiles at C = 256); the
  // widest intermediate (a1, 59 MB at C = 256) then stays in L2 from conv1 to conv2
  int64_t c = d->roi_chunk > 0 ? d->roi_chunk : 8 * kNumSMs;
  c = (c + 1) / 2 * 2;
  return (int)std::min<int64_t>(c, std::max<int64_t>(n, 2));
}

size_t head_carve(const osd_box_head_desc* d, Carver& c, HeadWorkspace* ws) {
  const size_t C = d->channels, ch = head_chunk(d), mlp = d->mlp_dim, n_all = (size_t)d->batch * d->rois_per_image;
  ws->sb = c.take<__nv_bfloat16>((size_t)d->batch * kPix * C);
  ws->xb = c.take<__nv_bfloat16>(ch * kPix * C);
  ws->a1 = c.take<__nv_bfloat16>(ch * kPix * 2 * C);
  ws->a2 = c.take<__nv_bfloat16>(ch * kPix * C);
  ws->a3 = c.take<__nv_bfloat16>(n_all * kPix * (C / 2));
  ws->h6 = c.take<__nv_bfloat16>(n_all * mlp);
  ws->h7 = c.take<__nv_bfloat16>(n_all * mlp);
  return c.total();
}

int head_validate(const osd_box_head_desc* d) {
  OSD_REQUIRE(d != nullptr, "osd_box_head: desc is null");
  OSD_REQUIRE(d->batch >= 0 && d->rois_per_image >= 1, "osd_box_head: bad batch / rois_per_image");
  OSD_REQUIRE(d->channels == 64 || d->channels == 128 || d->channels == 256, "osd_box_head: channels must be 64, 128 or 256 (got %d)",
              d->channels);
  OSD_REQUIRE(d->pooled_size == kP, "osd_box_head: pooled_size must be 7 (got %d)", d->pooled_size);
  OSD_REQUIRE(d->mlp_dim >= 8 && d->mlp_dim % 8 == 0, "osd_box_head: mlp_dim %d must be a positive multiple of 8", d->mlp_dim);
  OSD_REQUIRE(d->num_classes >= 1 && d->num_box_out >= 1 && d->num_classes + d->num_box_out <= 32,
              "osd_box_head: at most 32 predictor outputs (got %d + %d)", d->num_classes, d->num_box_out);
  OSD_REQUIRE((int64_t)d->batch * d->rois_per_image * kPix < (1ll << 31) / 8, "osd_box_head: too many ROIs; split the batch");
  return OSD_OK;
}

}  // namespace
}  // namespace osd

extern "C" int osd_box_head_workspace_bytes(const osd_box_head_desc* d, size_t* bytes) {
  int rc = osd::head_validate(d);
  if (rc != OSD_OK) return rc;
  OSD_REQUIRE(bytes != nullptr, "osd_box_head_workspace_bytes: bytes is null");
  osd::Carver c(nullptr);
  osd::HeadWorkspace ws;
  *bytes = osd::head_carve(d, c, &ws);
  return OSD_OK;
}

extern "C" int osd_box_head_forward(const osd_box_head_desc* d, void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace osd;
  int rc = head_validate(d);
  if (rc != OSD_OK) return rc;
  if (d->batch == 0) return OSD_OK;
  OSD_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "osd_box_head_forward: workspace must be 256-byte aligned");
  OSD_REQUIRE((d->pooled || d->pooled_nhwc_bf16) && d->supp && d->class_logits && d->box_regression,
              "osd_box_head_forward: null input / output");
  OSD_REQUIRE(d->pooled == nullptr || (reinterpret_cast<uintptr_t>(d->pooled) & 15) == 0, "osd_box_head_forward: pooled must be 16-byte aligned");
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(d->supp) & 15) == 0, "osd_box_head_forward: supp must be 16-byte aligned");
  OSD_REQUIRE(d->pooled_nhwc_bf16 == nullptr || (reinterpret_cast<uintptr_t>(d->pooled_nhwc_bf16) & 15) == 0,
              "osd_box_head_forward: pooled_nhwc_bf16 must be 16-byte aligned");
  OSD_REQUIRE(d->w1 && d->b1 && d->gn1_w && d->gn1_b && d->w2 && d->b2 && d->gn2_w && d->gn2_b && d->w3 && d->b3 && d->gn3_w &&
                  d->gn3_b && d->w6 && d->b6 && d->w7 && d->b7 && d->wp && d->bp,
              "osd_box_head_forward: null parameter");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver cv(workspace);
  HeadWorkspace ws;
  const size_t need = head_carve(d, cv, &ws);
  if (need > workspace_bytes) {
    set_error("osd_box_head_forward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    return OSD_ERR_WORKSPACE;
  }
  const int C = d->channels, C2 = 2 * C, Ch = C / 2, mlp = d->mlp_dim, R = d->rois_per_image;
  const int64_t n_all = (int64_t)d->batch * R;
  const int chunk = head_chunk(d);
  const int npred = d->num_classes + d->num_box_out;
  const size_t pk_smem = (size_t)kPix * (C + 2) * sizeof(__nv_bfloat16);

  timeline_mark("box_head_begin", stream);
  pack_roi_kernel<<<(unsigned)d->batch, 256, pk_smem, stream>>>(d->supp, ws.sb, C);
  OSD_LAUNCH_CHECK("pack_roi_kernel");

  // tensor maps that do not depend on the chunk
  CUtensorMap mSupp, mW1, mW2, mW3, mW6, mW7, mWp, mNone;
  memset(&mNone, 0, sizeof(mNone));
  if ((rc = make_bf16_map(ws.sb, (int64_t)d->batch * kPix, C, C, kBK, kPix, &mSupp)) != OSD_OK) return rc;
  // W boxes: the N tile's rows; pairs load half of them per CTA and multicast
  const bool pairs = head_pairs();
  const int wdiv = pairs ? 2 : 1;
  if ((rc = make_bf16_map(d->w1, C2, C2, C2, kBK, std::min(C2, 256) / wdiv, &mW1)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(d->w2, C, C2, C2, kBK, std::min(C, 256) / wdiv, &mW2)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(d->w3, Ch, 9 * C, 9 * C, kBK, Ch / wdiv, &mW3)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(d->w6, mlp, (int64_t)kPix * Ch, (int64_t)kPix * Ch, kBK, 256 / wdiv, &mW6)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(d->w7, mlp, mlp, mlp, kBK, 256 / wdiv, &mW7)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(d->wp, npred, mlp, mlp, kBK, 32 / wdiv, &mWp)) != OSD_OK) return rc;

  for (int64_t r0 = 0; r0 < n_all; r0 += chunk) {
    const int n = (int)std::min<int64_t>(chunk, n_all - r0);
    const void* xrows = ws.xb;
    if (d->pooled_nhwc_bf16) {
      xrows = static_cast<const __nv_bfloat16*>(d->pooled_nhwc_bf16) + (size_t)r0 * kPix * C;
    } else {
      pack_roi_kernel<<<(unsigned)n, 256, pk_smem, stream>>>(d->pooled + (size_t)r0 * C * kPix, ws.xb, C);
      OSD_LAUNCH_CHECK("pack_roi_kernel");
      timeline_mark("pack_roi_kernel", stream);
    }

    CUtensorMap mX, mA1, mA2;
    if ((rc = make_bf16_map(xrows, (int64_t)n * kPix, C, C, kBK, kPix, &mX)) != OSD_OK) return rc;
    if ((rc = make_bf16_map(ws.a1, (int64_t)n * kPix, C2, C2, kBK, kPix, &mA1)) != OSD_OK) return rc;
    if ((rc = make_roi_map_4d(ws.a2, n, C, &mA2)) != OSD_OK) return rc;

    GemmArgs G{};
    G.eps = d->gn_eps; G.slope = d->lrelu_slope; G.roi0 = (int)r0; G.rois_per_image = R;
    // conv1: K = [x | support] (concat never materialised) -> GN1 -> LeakyReLU
    G.M = n; G.N = C2; G.K = C2; G.a_mode = A_ROI; G.k_split = C;
    G.bias = d->b1; G.gamma = d->gn1_w; G.beta = d->gn1_b; G.out = ws.a1; G.ldo = C2;
    if ((rc = launch_gn_gemm(mX, mSupp, mW1, G, stream, "box_head_conv1", pairs)) != OSD_OK) return rc;
    // conv2 -> GN2 -> LeakyReLU
    G.N = C; G.K = C2; G.k_split = C2;
    G.bias = d->b2; G.gamma = d->gn2_w; G.beta = d->gn2_b; G.out = ws.a2; G.ldo = C;
    if ((rc = launch_gn_gemm(mA1, mNone, mW2, G, stream, "box_head_conv2", pairs)) != OSD_OK) return rc;
    // feature_aggreg: 3x3 conv as an implicit GEMM over 9 shifted boxes -> GN3 -> LeakyReLU
    G.N = Ch; G.K = 9 * C; G.a_mode = A_ROI_3X3; G.tap_c = C;
    G.bias = d->b3; G.gamma = d->gn3_w; G.beta = d->gn3_b; G.out = ws.a3 + (size_t)r0 * kPix * Ch; G.ldo = Ch;
    if ((rc = launch_gn_gemm(mA2, mNone, mW3, G, stream, "box_head_aggreg", pairs)) != OSD_OK) return rc;
  }

  // ---- the fully connected layers, once over all ROIs (full waves of 128-row tiles)
  CUtensorMap mA3, mH6, mH7;
  const int n = (int)n_all;
  if ((rc = make_bf16_map(ws.a3, n, (int64_t)kPix * Ch, (int64_t)kPix * Ch, kBK, kBM, &mA3)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(ws.h6, n, mlp, mlp, kBK, kBM, &mH6)) != OSD_OK) return rc;
  if ((rc = make_bf16_map(ws.h7, n, mlp, mlp, kBK, kBM, &mH7)) != OSD_OK) return rc;
  GemmArgs G{};
  G.eps = d->gn_eps; G.slope = d->lrelu_slope; G.rois_per_image = R;
  // fc6, fc7 (+ ReLU)
  G.a_mode = A_PLAIN; G.M = n; G.N = mlp; G.K = kPix * Ch; G.relu = 1; G.out_f32 = 0;
  G.bias = d->b6; G.out = ws.h6; G.ldo = mlp;
  if ((rc = launch_gemm<0, 128, 1, 2>(mA3, mNone, mW6, G, stream, "box_head_fc6", pairs)) != OSD_OK) return rc;
  G.K = mlp; G.bias = d->b7; G.out = ws.h7;
  if ((rc = launch_gemm<0, 128, 1, 2>(mH6, mNone, mW7, G, stream, "box_head_fc7", pairs)) != OSD_OK) return rc;
  // predictor: cls_score rows then bbox_pred rows of one small GEMM, fp32 outputs in the reference's two tensors
  G.N = npred; G.relu = 0; G.out_f32 = 1; G.bias = d->bp;
  G.out = d->class_logits; G.ldo = d->num_classes;
  G.out2 = d->box_regression; G.ldo2 = d->num_box_out; G.n_split = d->num_classes;
  return launch_gemm<0, 16, 1, 2>(mH7, mNone, mWp, G, stream, "box_head_predictor", pairs);
}
