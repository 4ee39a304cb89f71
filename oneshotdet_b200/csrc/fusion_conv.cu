// 1x1 fusion conv of the matching module on tcgen05 tensor cores (sm_100a), hand-written: no CUTLASS.
//
// Reference: `compress_dim_conv` (maskrcnn_benchmark/modeling/roi_heads/box_head/box_head.py:43-54) applied to
// cat((x, support.expand_as(x)), dim=1) (:147-149):
//     Conv1x1(2C -> 2C) + GroupNorm(32, 2C) + LeakyReLU(0.2) + Conv1x1(2C -> C) + GroupNorm(32, C) + LeakyReLU(0.2)
// here on the FPN maps [B, C, H, W] (north_star), bf16 operands / fp32 accumulation.
//
//  * The concat is never materialised:  W1 . [x ; s] = W1x . x + (W1s . s + b1).  The support half folds into a
//    per-(episode, level) bias vector (fusion_bias_kernel, fp32), so conv1 is a [2C x C] GEMM per pixel tile.
//  * One GEMM kernel serves both convs (conv1x1_tc_kernel).  Output channels are the UMMA M dimension (A = bf16
//    weights, K-major, TMA with 128-byte swizzle), pixels are N (B = activations).  NCHW activations are already
//    "N-major": producer warps read fp32 rows, apply the input transform (identity, or GroupNorm+LeakyReLU of the
//    previous conv), convert to bf16 and store straight into the MN-major 128B-swizzled canonical layout the
//    tensor core reads -- no transpose, no extra pass over HBM.
//  * Accumulators live in TMEM: (Cout/128) tiles of 128 x 64 fp32, double buffered (512 columns for Cout = 512),
//    so the epilogue of tile i overlaps the MMAs of tile i+1.  Epilogue warps read TMEM (tcgen05.ld 32x32b), add
//    the bias, accumulate GroupNorm statistics (fp64 atomics per (episode, group)) and store NCHW rows.
//  * Warp roles (576 threads, 1 CTA/SM, persistent over pixel tiles): warps 0-7 epilogue, 8-15 activation
//    producers, warp 16 TMA producer (weights), warp 17 MMA issuer + TMEM owner.  All hand-offs are mbarriers.
//    The epilogue transposes each 32x32 accumulator block through a swizzled shared-memory stage so that global
//    stores are 128-byte row segments (4 rows per warp instruction) instead of 32 scattered 16-byte pieces.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstring>

#include "osd_common.cuh"
#include "osd_device_utils.cuh"
#include "osd_tc.cuh"
#include "fusion_internal.cuh"

namespace osd {
namespace {

constexpr int kBlockN = 64;   // pixels per tile (UMMA N)
constexpr int kBlockK = 64;   // K elements per weight stage (128 bytes of bf16 = one swizzle row)
constexpr int kUmmaM = 128;
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8, kProdWarps = 8;          // warps 0-7 epilogue, 8-15 activation producers
constexpr int kTmaWarp = 16, kMmaWarp = 17;
constexpr int kThreads = 32 * 18;
constexpr int kStageBytesPerWarp = 4096;              // epilogue transpose buffer: 32 rows x 32 fp32
constexpr int kMaxStages = 4;

using namespace tc;

// ------------------------------------------------------------------------------------------------
// kernel arguments
// ------------------------------------------------------------------------------------------------
struct ConvLevel {
  const float* in;     // [B, Cin, HW] fp32
  float* out;          // [B, Cout, HW] fp32
  int hw;
  int tiles_per_img;   // ceil(hw / 64)
  int tile_begin;      // first global tile index of this level
};

struct ConvArgs {
  int nl, B, Cin, Cout;
  int num_mt;           // ceil(Cout / 128)
  int num_kc;           // Cin / 64
  int stages;           // weight stages in shared memory
  int total_tiles;
  int tma_out_mask;     // bit l: level l's output has a tensor map (epilogue stores through TMA)
  int xform;            // 0: identity; 1: GroupNorm(32, Cin) + LeakyReLU on the input (stats_in)
  float eps, slope;
  const float* bias;    // [nl, B, Cout] (bias_level_stride / bias_img_stride may be 0)
  int bias_level_stride, bias_img_stride;
  const float2* coef_in;    // [nl, B, Cin] (scale, shift) of the input's GroupNorm (xform = 1)
  double* stats_out;        // [nl, B, 32, 2]
  ConvLevel lv[OSD_MAX_LEVELS];
};

struct TileInfo {
  int level, img, px0, nvalid;
};

// output tensor maps, one per level ([B*Cout rows, HW columns] fp32, box 32 x 32, 128-byte swizzle); only used for
// levels with HW % 4 == 0 (TMA needs a 16-byte row pitch)
struct OutMaps {
  CUtensorMap m[OSD_MAX_LEVELS];
};

__device__ __forceinline__ TileInfo decode_tile(const ConvArgs& A, int tile) {
  int li = 0;
#pragma unroll
  for (int k = 1; k < OSD_MAX_LEVELS; ++k)
    if (k < A.nl && tile >= A.lv[k].tile_begin) li = k;
  const int local = tile - A.lv[li].tile_begin;
  const int tpi = A.lv[li].tiles_per_img;
  TileInfo t;
  t.level = li;
  t.img = local / tpi;
  t.px0 = (local - t.img * tpi) * kBlockN;
  t.nvalid = min(kBlockN, A.lv[li].hw - t.px0);
  return t;
}


// ------------------------------------------------------------------------------------------------
// the GEMM kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ OutMaps out_maps, const ConvArgs A) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [B buffers: 2 x (Cin x 128 B)] [A stages: stages x (num_mt x 16 KB)] [barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_bytes = (uint32_t)A.Cin * 128u;
  const uint32_t a_stage_bytes = (uint32_t)A.num_mt * 16384u;
  const uint32_t sB = smem_base;
  const uint32_t sA = sB + 2u * b_bytes;
  const uint32_t sStage = sA + (uint32_t)A.stages * a_stage_bytes;   // epilogue transpose buffers
  const uint32_t sBar = sStage + kEpiWarps * kStageBytesPerWarp;
  // barrier slots (8 bytes each)
  const uint32_t a_full = sBar, a_empty = sBar + 8u * kMaxStages;
  const uint32_t b_full = sBar + 16u * kMaxStages, b_empty = b_full + 16u;
  const uint32_t t_full = b_full + 32u, t_empty = t_full + 16u;
  const uint32_t tmem_slot = t_full + 32u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));  // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (A.num_mt * 2 * kBlockN <= 128) ? 128u : (A.num_mt * 2 * kBlockN <= 256 ? 256u : 512u);

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < A.stages; ++s) {
      mbar_init(a_full + 8u * s, 1);
      mbar_init(a_empty + 8u * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_full + 8u * s, kProdWarps);
      mbar_init(b_empty + 8u * s, 1);
      mbar_init(t_full + 8u * s, 1);
      mbar_init(t_empty + 8u * s, kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp == kTmaWarp) {
    // ===================== TMA producer: weight chunks [Cout x 64] =====================
    if (lane == 0) {
      uint32_t c = 0;
      for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < A.num_kc; ++kc, ++c) {
          const uint32_t s = c % A.stages, ph = (c / A.stages) & 1u;
          mbar_wait(a_empty + 8u * s, ph ^ 1u);
          mbar_expect_tx(a_full + 8u * s, a_stage_bytes);
          for (int mt = 0; mt < A.num_mt; ++mt)
            tma_load_2d(sA + s * a_stage_bytes + mt * 16384u, &tmap_w, kc * kBlockK, mt * kUmmaM, a_full + 8u * s);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kUmmaM, kBlockN);
      uint32_t c = 0, it = 0;
      for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait(t_empty + 8u * buf, ph ^ 1u);   // epilogue drained this accumulator stage
        mbar_wait(b_full + 8u * buf, ph);         // activations converted
        tc_fence_after();
        const uint32_t acc_col = buf * (uint32_t)(A.num_mt * kBlockN);
        for (int kc = 0; kc < A.num_kc; ++kc, ++c) {
          const uint32_t s = c % A.stages, aph = (c / A.stages) & 1u;
          mbar_wait(a_full + 8u * s, aph);
          tc_fence_after();
#pragma unroll
          for (int k16 = 0; k16 < kBlockK / kUmmaK; ++k16) {
            // B: MN-major SW128; 8-channel groups are 1024 B apart (SBO); one 64-pixel block (LBO unused)
            const uint32_t kgrp = (uint32_t)(kc * kBlockK + k16 * kUmmaK) >> 3;
            const uint64_t bdesc = make_smem_desc(sB + buf * b_bytes + kgrp * 1024u, b_bytes, 1024u);
            for (int mt = 0; mt < A.num_mt; ++mt) {
              // A: K-major SW128; 8-row groups are 1024 B apart (SBO), K advances by 32 B inside the swizzle row
              const uint64_t adesc = make_smem_desc(sA + s * a_stage_bytes + mt * 16384u + k16 * 32u, 16u, 1024u);
              umma_bf16(tmem_base + acc_col + mt * kBlockN, adesc, bdesc, idesc, (kc | k16) != 0 ? 1u : 0u);
            }
          }
          umma_commit(a_empty + 8u * s);  // weights of this stage consumed
        }
        umma_commit(t_full + 8u * buf);   // accumulators ready for the epilogue
        umma_commit(b_empty + 8u * buf);  // activation buffer may be refilled
      }
    }
  } else if (warp >= kEpiWarps) {
    // ===================== activation producers (warps 8-15) =====================
    const int pw = warp - kEpiWarps;
    const int rsub = lane >> 3;   // row inside a group of 4
    const int j = lane & 7;       // 16-byte chunk (8 pixels) inside the 128-byte row
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      const TileInfo t = decode_tile(A, tile);
      const ConvLevel& L = A.lv[t.level];
      const float* src = L.in + (size_t)t.img * A.Cin * L.hw + t.px0 + j * 8;
      const bool vec_ok = ((L.hw & 3) == 0) && (t.px0 + j * 8 + 8 <= L.hw);
      const int nleft = L.hw - (t.px0 + j * 8);  // valid pixels from this chunk's start (may be <= 0)
      const float2* coef = A.xform ? A.coef_in + ((size_t)t.level * A.B + t.img) * A.Cin : nullptr;
      // 256 input channels at a time: every thread first issues ALL its loads of the block (8 rows x 2 x 128 bit
      // = 64 KB in flight per SM, enough to cover HBM latency at the SM's share of the bandwidth) and only then
      // waits for the activation buffer to be free -- the loads of tile i+1 fly while tile i is multiplied.
      for (int kh = 0; kh < A.Cin; kh += 256) {
        float v[8][8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = kh + pw * 4 + rsub + u * 32;
          const float* p = src + (size_t)k * L.hw;
          if (k < A.Cin) {
            if (vec_ok) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(p));
              const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
              v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
              v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q) v[u][q] = (q < nleft) ? __ldg(p + q) : 0.f;
            }
          }
        }
        if (kh == 0) mbar_wait(b_empty + 8u * buf, ph ^ 1u);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = kh + pw * 4 + rsub + u * 32;
          if (k < A.Cin) {
            if (A.xform) {
              // GroupNorm + LeakyReLU of the producer conv's output: y = x * scale + shift
              const float2 cf = __ldg(coef + k);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float y = fmaf(v[u][q], cf.x, cf.y);
                y = y > 0.f ? y : y * A.slope;
                v[u][q] = (q < nleft) ? y : 0.f;   // padded pixels stay exactly zero
              }
            }
            const uint32_t dst = sB + buf * b_bytes + (uint32_t)(k >> 3) * 1024u + (uint32_t)(k & 7) * 128u +
                                 (uint32_t)((j ^ (k & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16x2(v[u][0], v[u][1])),
                         "r"(pack_bf16x2(v[u][2], v[u][3])), "r"(pack_bf16x2(v[u][4], v[u][5])),
                         "r"(pack_bf16x2(v[u][6], v[u][7]))
                         : "memory");
          }
        }
      }
      fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(b_full + 8u * buf);
    }
  } else {
    // ===================== epilogue (warps 0-7) =====================
    // warp w reads TMEM lanes 32q..32q+31 (q = w & 3: the hardware's lane quadrant of the warp) and the 32-pixel
    // column half h = w >> 2 of every 128-channel accumulator tile
    const int q = warp & 3, h = warp >> 2;
    const int gs_out = A.Cout / 32;  // channels per GroupNorm group of the output (power of two <= 16)
    const uint32_t stage = sStage + (uint32_t)warp * kStageBytesPerWarp;
    float* stage_gen = reinterpret_cast<float*>(gen_base + (stage - smem_base));
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      const TileInfo t = decode_tile(A, tile);
      const ConvLevel& L = A.lv[t.level];
      mbar_wait(t_full + 8u * buf, ph);
      tc_fence_after();
      const bool vec_ok = (L.hw & 3) == 0;
      const int pxh = h * 32;                       // first pixel of this warp's half inside the tile
      const int nval = min(32, t.nvalid - pxh);     // valid pixels in the half (may be <= 0)
      for (int mt = 0; mt < A.num_mt; ++mt) {
        const int oc0 = mt * kUmmaM + q * 32;       // first channel of this warp's 32 rows
        const int oc = oc0 + lane;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(A.num_mt * kBlockN) +
                               mt * kBlockN + pxh;
        const bool oc_ok = oc < A.Cout;
        const float bias = oc_ok ? A.bias[(size_t)t.level * A.bias_level_stride + (size_t)t.img * A.bias_img_stride + oc] : 0.f;
        uint32_t r[32];
        tmem_ld32(taddr, r);
        // the previous bulk store of this warp must have read the stage before it is rewritten; waiting here rather
        // than right after issuing it lets the store overlap this tile's TMEM load
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
        tmem_ld_wait();
        float s1 = 0.f, s2 = 0.f;
        // bias, statistics, and the row (channel) of this lane into the swizzled stage: 16-byte chunk c of row
        // `lane` lives at chunk (c ^ (lane & 7))
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float y[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            y[e] = __uint_as_float(r[4 * c + e]) + bias;
            if (4 * c + e < nval) {
              s1 += y[e];
              s2 += y[e] * y[e];
            }
          }
          const uint32_t a = stage + (uint32_t)lane * 128u + (uint32_t)((c ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(y[0]), "f"(y[1]), "f"(y[2]), "f"(y[3]) : "memory");
        }
        if ((A.tma_out_mask >> t.level) & 1) {
          // the stage is exactly a SWIZZLE_128B box of 32 rows x 32 fp32: hand it to the TMA unit, which clips pixels
          // past the end of the level; global stores leave the LSU / L1 path entirely
          fence_proxy_async();
          __syncwarp();
          if (lane == 0 && oc0 < A.Cout) {
            tma_store_2d(&out_maps.m[t.level], stage, t.px0 + pxh, t.img * A.Cout + oc0);
            tma_store_commit();      // (waited for at the top of the next accumulator tile)
          }
          __syncwarp();
        } else {
          __syncwarp();
          // levels whose rows are not 16-byte aligned (H*W % 4 != 0): read back transposed and store with the LSU
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + (lane >> 3), c = lane & 7;
            const int orow = oc0 + row;
            const float4 v = *reinterpret_cast<const float4*>(stage_gen + row * 32 + ((c ^ (row & 7)) << 2));
            if (orow < A.Cout) {
              float* dst = L.out + ((size_t)t.img * A.Cout + orow) * L.hw + t.px0 + pxh + 4 * c;
              const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (4 * c + e < nval) dst[e] = vs[e];
            }
          }
          __syncwarp();  // the stage is rewritten by the next accumulator tile
        }
        if (A.stats_out) {
          // GroupNorm statistics of this conv's output: reduce over the gs_out consecutive channels (lanes) of a group
          for (int o = 1; o < gs_out; o <<= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (oc_ok && (lane & (gs_out - 1)) == 0 && nval > 0) {
            double* so = A.stats_out + ((size_t)t.level * A.B + t.img) * 64 + 2 * (oc / gs_out);
            atomicAdd(so, (double)s1);
            atomicAdd(so + 1, (double)s2);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + 8u * buf);
    }
    if (lane == 0) tma_store_wait_all();   // every bulk store of this warp has been written
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// small kernels around the GEMMs
// ------------------------------------------------------------------------------------------------
// bias_eff[l, b, oc] = b1[oc] + sum_c W1s_t[c, oc] * mean_s supp[l][b*S + s, c]      (fp32)
struct BiasArgs {
  int nl, B, S, C, Cout;
  const float* supp[OSD_MAX_LEVELS];  // [B*S, C]
  const float* w1s_t;                 // [C, Cout]
  const float* b1;                    // [Cout]
  float* bias_eff;                    // [nl, B, Cout]
};

// grid (nl*B, Cout/64): 256 threads = 64 output channels x 4 slices of the C inputs, reduced through shared memory
__global__ void __launch_bounds__(256) fusion_bias_kernel(BiasArgs A) {
  extern __shared__ float bias_smem[];  // pooled[C] | partial[4][64]
  float* pooled = bias_smem;
  float* part = bias_smem + A.C;
  const int l = blockIdx.x / A.B, b = blockIdx.x % A.B;
  const float* s = A.supp[l] + (size_t)b * A.S * A.C;
  for (int c = threadIdx.x; c < A.C; c += blockDim.x) {
    float acc = s[c];
    for (int k = 1; k < A.S; ++k) acc = __fadd_rn(acc, s[(size_t)k * A.C + c]);
    pooled[c] = A.S == 1 ? acc : __fdiv_rn(acc, (float)A.S);
  }
  __syncthreads();
  const int ocl = threadIdx.x & 63, slice = threadIdx.x >> 6;
  const int oc = blockIdx.y * 64 + ocl;
  const int per = A.C / 4;
  float acc = 0.f;
  if (oc < A.Cout) {
    const float* w = A.w1s_t + (size_t)(slice * per) * A.Cout + oc;
#pragma unroll 8
    for (int c = 0; c < per; ++c) acc = fmaf(w[(size_t)c * A.Cout], pooled[slice * per + c], acc);
  }
  part[slice * 64 + ocl] = acc;
  __syncthreads();
  if (slice == 0 && oc < A.Cout)
    A.bias_eff[((size_t)l * A.B + b) * A.Cout + oc] = A.b1[oc] + ((part[ocl] + part[64 + ocl]) + (part[128 + ocl] + part[192 + ocl]));
}

// (scale, shift) of GroupNorm(32, C) per (level, episode, channel) from the fp64 sums the conv epilogue accumulated:
//   y = x * scale + shift,  scale = gamma * rstd,  shift = beta - mean * scale   (biased variance, as torch)
struct CoefArgs {
  int nl, B, C;
  float eps;
  const double* stats;  // [nl, B, 32, 2]
  const float* gn_w;
  const float* gn_b;
  const float* fold_bias;  // [nl, B, C] or nullptr: bias the consumer adds before normalising, absorbed into the shift
  float2* coef;         // [nl, B, C]
  int hw[OSD_MAX_LEVELS];
};

__global__ void __launch_bounds__(512) fusion_gn_coef_kernel(CoefArgs A) {
  const int l = blockIdx.x / A.B, b = blockIdx.x % A.B;
  const int gs = A.C / 32;
  const double inv_cnt = 1.0 / ((double)gs * (double)A.hw[l]);
  for (int c = threadIdx.x; c < A.C; c += blockDim.x) {
    const double* st = A.stats + ((size_t)l * A.B + b) * 64 + 2 * (c / gs);
    const double mean = st[0] * inv_cnt;
    const double var = fmax(st[1] * inv_cnt - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + (double)A.eps));
    const float sc = A.gn_w[c] * rstd;
    float sh = A.gn_b[c] - (float)mean * sc;
    if (A.fold_bias) sh = fmaf(A.fold_bias[((size_t)l * A.B + b) * A.C + c], sc, sh);
    A.coef[((size_t)l * A.B + b) * A.C + c] = make_float2(sc, sh);
  }
}

// out = LeakyReLU(y * scale[plane] + shift[plane]) in place: persistent grid, 128-bit loads/stores, 4 in flight
struct GnLevel {
  float* y;
  const float2* coef;   // [B*C]
  uint32_t hw;
  uint32_t elems;       // B * C * hw
  uint32_t chunk_begin;
  FastDiv div_hw;
};
struct GnArgs {
  int nl;
  float slope;
  uint32_t total_chunks;
  GnLevel lv[OSD_MAX_LEVELS];
};
constexpr int kGnThreads = 256, kGnVec = 4, kGnChunk = kGnThreads * kGnVec;

__global__ void __launch_bounds__(kGnThreads, 5) fusion_gn_lrelu_kernel(GnArgs A) {
  for (uint32_t chunk = blockIdx.x; chunk < A.total_chunks; chunk += gridDim.x) {
    int li = 0;
#pragma unroll
    for (int k = 1; k < OSD_MAX_LEVELS; ++k)
      if (k < A.nl && chunk >= A.lv[k].chunk_begin) li = k;
    const GnLevel& L = A.lv[li];
    const uint32_t nvec = (L.elems + 3) / 4;
    const uint32_t v0 = (chunk - L.chunk_begin) * kGnChunk;
    float4 in[kGnVec];
    uint32_t pl[kGnVec];
    bool whole[kGnVec];
#pragma unroll
    for (int j = 0; j < kGnVec; ++j) {
      const uint32_t v = v0 + j * kGnThreads + threadIdx.x;
      whole[j] = false;
      if (v < nvec) {
        const uint32_t g = v * 4;
        pl[j] = fdiv(g, L.div_hw);
        const uint32_t r = g - pl[j] * L.hw;
        whole[j] = (r + 4 <= L.hw) && (g + 4 <= L.elems) && ((L.hw & 3) == 0 || ((reinterpret_cast<uintptr_t>(L.y + g) & 15) == 0));
        if (whole[j]) in[j] = *reinterpret_cast<const float4*>(L.y + g);
      }
    }
#pragma unroll
    for (int j = 0; j < kGnVec; ++j) {
      const uint32_t v = v0 + j * kGnThreads + threadIdx.x;
      if (v >= nvec) continue;
      const uint32_t g = v * 4;
      if (whole[j]) {
        const float2 cf = __ldg(L.coef + pl[j]);
        float4 o;
        o.x = fmaf(in[j].x, cf.x, cf.y); o.x = o.x > 0.f ? o.x : o.x * A.slope;
        o.y = fmaf(in[j].y, cf.x, cf.y); o.y = o.y > 0.f ? o.y : o.y * A.slope;
        o.z = fmaf(in[j].z, cf.x, cf.y); o.z = o.z > 0.f ? o.z : o.z * A.slope;
        o.w = fmaf(in[j].w, cf.x, cf.y); o.w = o.w > 0.f ? o.w : o.w * A.slope;
        *reinterpret_cast<float4*>(L.y + g) = o;
      } else {
        for (int k = 0; k < 4 && g + k < L.elems; ++k) {
          const uint32_t p = fdiv(g + k, L.div_hw);
          const float2 cf = __ldg(L.coef + p);
          float o = fmaf(L.y[g + k], cf.x, cf.y);
          L.y[g + k] = o > 0.f ? o : o * A.slope;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
}  // namespace

int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    OSD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled is not available from the driver");
      return OSD_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return OSD_OK;
}

int make_bf16_map(const void* base, int64_t rows, int64_t cols, int64_t pitch_elems, int box_cols, int box_rows,
                  CUtensorMap* map) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc != OSD_OK) return rc;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return OSD_ERR_CUDA;
  }
  return OSD_OK;
}

namespace {

// bf16 weights [rows, cols] row-major -> 2-D tensor map, box = 64 (K) x 128 (rows), 128-byte swizzle;
// rows past the end (Cout < 128) are zero-filled by the TMA unit
int make_weight_map(const void* w, int rows, int cols, CUtensorMap* map) {
  return make_bf16_map(w, rows, cols, cols, kBlockK, kUmmaM, map);
}

struct ConvPlan {
  int stages;
  size_t smem;
};

int plan_conv(int Cin, int Cout, ConvPlan* p) {
  OSD_REQUIRE(Cin % 64 == 0 && Cin >= 64 && Cin <= 512, "fusion: input channels %d must be a multiple of 64 in [64, 512]", Cin);
  OSD_REQUIRE(Cout == 64 || Cout == 128 || Cout == 256 || Cout == 512, "fusion: output channels %d must be 64, 128, 256 or 512", Cout);
  const int num_mt = (Cout + kUmmaM - 1) / kUmmaM;
  const size_t b_bytes = 2 * (size_t)Cin * 128;
  const size_t a_stage = (size_t)num_mt * 16384;
  const size_t fixed = 1024 + 256 + (size_t)kEpiWarps * kStageBytesPerWarp;   // alignment slack, barriers, epilogue stages
  const size_t budget = 227 * 1024;
  int stages = (int)((budget - b_bytes - fixed) / a_stage);
  if (stages > kMaxStages) stages = kMaxStages;
  OSD_REQUIRE(stages >= 2, "fusion: %d -> %d channels does not fit in shared memory", Cin, Cout);
  p->stages = stages;
  p->smem = fixed + b_bytes + stages * a_stage;
  return OSD_OK;
}

// fp32 output [rows, cols] row-major -> 2-D tensor map, box = 32 (pixels) x 32 (rows), 128-byte swizzle
int make_output_map(float* base, int64_t rows, int64_t cols, CUtensorMap* map) {
  EncodeTiledFn fn;
  int rc = get_encode_fn(&fn);
  if (rc != OSD_OK) return rc;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)r);
    return OSD_ERR_CUDA;
  }
  return OSD_OK;
}

int launch_conv(const CUtensorMap& map, ConvArgs& A, cudaStream_t stream) {
  ConvPlan p;
  int rc = plan_conv(A.Cin, A.Cout, &p);
  if (rc != OSD_OK) return rc;
  A.stages = p.stages;
  A.num_mt = (A.Cout + kUmmaM - 1) / kUmmaM;
  A.num_kc = A.Cin / kBlockK;
  rc = ensure_dynamic_smem(reinterpret_cast<const void*>(conv1x1_tc_kernel), 227 * 1024);
  if (rc != OSD_OK) return rc;
  if (A.total_tiles <= 0) return OSD_OK;
  const int grid = A.total_tiles < kNumSMs ? A.total_tiles : kNumSMs;
  OutMaps om;
  memset(&om, 0, sizeof(om));
  A.tma_out_mask = 0;
  for (int l = 0; l < A.nl; ++l) {
    if ((A.lv[l].hw & 3) == 0 && (reinterpret_cast<uintptr_t>(A.lv[l].out) & 15) == 0) {
      rc = make_output_map(A.lv[l].out, (int64_t)A.B * A.Cout, A.lv[l].hw, &om.m[l]);
      if (rc != OSD_OK) return rc;
      A.tma_out_mask |= 1 << l;
    }
  }
  conv1x1_tc_kernel<<<grid, kThreads, p.smem, stream>>>(map, om, A);
  OSD_LAUNCH_CHECK("conv1x1_tc_kernel");
  return OSD_OK;
}

}  // namespace

int fusion_launch_gn_coef(int nl, int B, int C, float eps, const double* stats, const float* gn_w, const float* gn_b,
                          const float* fold_bias, float2* coef, const int32_t* hw, cudaStream_t stream) {
  CoefArgs K{};
  K.nl = nl; K.B = B; K.C = C; K.eps = eps; K.stats = stats; K.gn_w = gn_w; K.gn_b = gn_b; K.fold_bias = fold_bias;
  K.coef = coef;
  for (int l = 0; l < nl; ++l) K.hw[l] = hw[l];
  fusion_gn_coef_kernel<<<nl * B, 512, 0, stream>>>(K);
  OSD_LAUNCH_CHECK("fusion_gn_coef_kernel");
  return OSD_OK;
}

int fusion_launch_gn_lrelu(int nl, int B, int C, float slope, float* const* y, const float2* coef, const int32_t* hw,
                           cudaStream_t stream) {
  GnArgs G{};
  G.nl = nl; G.slope = slope;
  uint64_t chunks = 0;
  for (int l = 0; l < nl; ++l) {
    GnLevel& L = G.lv[l];
    L.y = y[l];
    L.coef = coef + (size_t)l * B * C;
    L.hw = (uint32_t)hw[l];
    L.elems = (uint32_t)((size_t)B * C * hw[l]);
    L.chunk_begin = (uint32_t)chunks;
    L.div_hw = make_fastdiv(L.hw);
    chunks += ((uint64_t)L.elems + 4ull * kGnChunk - 1) / (4ull * kGnChunk);
  }
  G.total_chunks = (uint32_t)chunks;
  const int ctas = (int)std::min<uint64_t>(chunks, (uint64_t)kNumSMs * 16);
  if (ctas > 0) {
    fusion_gn_lrelu_kernel<<<ctas, kGnThreads, 0, stream>>>(G);
    OSD_LAUNCH_CHECK("fusion_gn_lrelu_kernel");
  }
  return OSD_OK;
}

}  // namespace osd

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static size_t fusion_carve(const osd_fusion_desc* d, osd::Carver& c, osd::FusionWorkspace* ws) {
  const size_t per = (size_t)d->num_levels * d->batch;
  ws->stats1 = c.take<double>(per * 64);
  ws->stats2 = c.take<double>(per * 64);
  ws->bias_eff = c.take<float>(per * 2 * d->channels);
  ws->coef1 = c.take<float2>(per * 2 * d->channels);
  ws->coef2 = c.take<float2>(per * d->channels);
  ws->xb = nullptr;
  if (d->stage == OSD_FUSION_FULL) {
    size_t elems = 0;
    for (int l = 0; l < d->num_levels; ++l)
      elems += (size_t)d->batch * d->channels * (size_t)osd::fusion_xb_pitch(d->hw[l]);
    ws->xb = c.take<uint16_t>(elems);
  }
  ws->gseg = nullptr; ws->rowsum = nullptr; ws->seg_plane = nullptr; ws->max_segments = 0;
  if (d->stage == OSD_FUSION_FULL && d->channels >= 128) {
    ws->max_segments = osd::kNumSMs + (int)per;
    ws->gseg = c.take<float>((size_t)ws->max_segments * d->channels * d->channels);
    ws->rowsum = c.take<float>((size_t)ws->max_segments * d->channels);
    ws->seg_plane = c.take<int>((size_t)ws->max_segments);
  }
  return c.total();
}

static int fusion_validate(const osd_fusion_desc* d) {
  OSD_REQUIRE(d != nullptr, "osd_fusion: desc is null");
  OSD_REQUIRE(d->num_levels >= 1 && d->num_levels <= OSD_MAX_LEVELS, "osd_fusion: num_levels %d out of range", d->num_levels);
  OSD_REQUIRE(d->batch >= 0 && d->shots >= 1, "osd_fusion: bad batch / shots");
  OSD_REQUIRE(d->channels == 64 || d->channels == 128 || d->channels == 256,
              "osd_fusion: channels must be 64, 128 or 256 (got %d)", d->channels);
  OSD_REQUIRE(d->stage == OSD_FUSION_CONV1 || d->stage == OSD_FUSION_FULL, "osd_fusion: unknown stage %d", d->stage);
  for (int l = 0; l < d->num_levels; ++l) {
    OSD_REQUIRE(d->hw[l] >= 1, "osd_fusion: level %d is empty", l);
    OSD_REQUIRE((int64_t)d->batch * 2 * d->channels * d->hw[l] < (1ll << 31), "osd_fusion: level %d too large; split the batch", l);
  }
  return OSD_OK;
}

extern "C" int osd_fusion_workspace_bytes(const osd_fusion_desc* d, size_t* bytes) {
  int rc = fusion_validate(d);
  if (rc != OSD_OK) return rc;
  OSD_REQUIRE(bytes != nullptr, "osd_fusion_workspace_bytes: bytes is null");
  osd::Carver c(nullptr);
  osd::FusionWorkspace ws;
  *bytes = fusion_carve(d, c, &ws);
  return OSD_OK;
}

extern "C" int osd_fusion_forward(const osd_fusion_desc* d, void* workspace, size_t workspace_bytes, void* stream_) {
  using namespace osd;
  int rc = fusion_validate(d);
  if (rc != OSD_OK) return rc;
  if (d->batch == 0) return OSD_OK;
  OSD_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "osd_fusion_forward: workspace must be 256-byte aligned");
  OSD_REQUIRE(d->w1x_bf16 && d->w1s_t && d->b1, "osd_fusion_forward: conv1 weights are null");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver c(workspace);
  FusionWorkspace ws;
  const size_t need = fusion_carve(d, c, &ws);
  if (need > workspace_bytes) {
    set_error("osd_fusion_forward: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    return OSD_ERR_WORKSPACE;
  }
  const int C = d->channels, C2 = 2 * C, B = d->batch, nl = d->num_levels;
  const bool full = d->stage == OSD_FUSION_FULL;
  if (full)
    OSD_REQUIRE(d->w2_bf16 && d->b2 && d->gn1_w && d->gn1_b && d->gn2_w && d->gn2_b, "osd_fusion_forward: conv2 / GroupNorm parameters are null");

  // ---- folded bias: W1s . mean_s(support) + b1 per (level, episode)
  BiasArgs BA{};
  BA.nl = nl; BA.B = B; BA.S = d->shots; BA.C = C; BA.Cout = C2;
  for (int l = 0; l < nl; ++l) {
    OSD_REQUIRE(d->feat[l] && d->supp[l] && d->out[l], "osd_fusion_forward: null pointer at level %d", l);
    BA.supp[l] = static_cast<const float*>(d->supp[l]);
  }
  BA.w1s_t = d->w1s_t; BA.b1 = d->b1; BA.bias_eff = ws.bias_eff;
  fusion_bias_kernel<<<dim3((unsigned)(nl * B), (unsigned)((C2 + 63) / 64)), 256, (C + 256) * sizeof(float), stream>>>(BA);
  OSD_LAUNCH_CHECK("fusion_bias_kernel");

  if (full) return fusion_full_forward(d, ws, stream);

  // ---- stage CONV1: x [B,C,HW] -> out [B,2C,HW] (+ folded bias)
  CUtensorMap map1;
  rc = make_weight_map(d->w1x_bf16, C2, C, &map1);
  if (rc != OSD_OK) return rc;
  ConvArgs A1{};
  A1.nl = nl; A1.B = B; A1.Cin = C; A1.Cout = C2; A1.xform = 0; A1.eps = d->gn_eps; A1.slope = d->lrelu_slope;
  A1.bias = ws.bias_eff; A1.bias_level_stride = B * C2; A1.bias_img_stride = C2;
  A1.stats_out = nullptr;
  int tiles = 0;
  for (int l = 0; l < nl; ++l) {
    ConvLevel& L = A1.lv[l];
    L.in = static_cast<const float*>(d->feat[l]);
    L.out = static_cast<float*>(d->out[l]);
    L.hw = d->hw[l];
    L.tiles_per_img = (d->hw[l] + kBlockN - 1) / kBlockN;
    L.tile_begin = tiles;
    tiles += B * L.tiles_per_img;
  }
  A1.total_tiles = tiles;
  return launch_conv(map1, A1, stream);
}
