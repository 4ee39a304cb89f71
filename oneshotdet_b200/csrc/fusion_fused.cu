// The full 1x1 fusion module as a back-to-back GEMM on tcgen05 (sm_100a), hand-written: no CUTLASS.
//
// Reference: `compress_dim_conv` (maskrcnn_benchmark/modeling/roi_heads/box_head/box_head.py:43-54) applied to
// cat((x, support.expand_as(x)), dim=1) (:147-149):
//     Conv1x1(2C -> 2C) + GroupNorm(32, 2C) + LeakyReLU(0.2) + Conv1x1(2C -> C) + GroupNorm(32, C) + LeakyReLU(0.2)
//
// Round 1 ran the two convolutions as two GEMM launches with an fp32 [B, 2C, HW] intermediate in HBM (written once, read
// once) and streamed every weight byte from L2 once per 64 pixels; it was bound by that L2 stream and by the 2.9 GB of
// DRAM traffic.  Here:
//
//  * PIXELS are the UMMA M dimension (tile = 128 pixels = the 128 TMEM lanes), channels are N.  The activations
//    [C, 128 px] (NCHW rows, pixel-contiguous) are the MN-major A operand; the weights [Cout, Cin] row-major are the K-major
//    B operand (TMA, 128-byte swizzle, 16 KB stages of 128 rows x 64 K).  One pass over the weights serves 128 pixels.
//  * conv1 accumulates 128 output channels at a time into a TMEM chunk D1 [128 px x 128 ch] (two chunk buffers).
//    The chunk's epilogue warps read it (tcgen05.ld, thread = pixel), apply GroupNorm-1 (scale/shift per channel, the
//    folded support bias absorbed into the shift) and LeakyReLU, convert to bf16 and write the result back INTO THE
//    SAME TMEM COLUMNS (tcgen05.st) -- as the K-major A operand of conv2:  D2 [128 px x C] += y1_chunk . W2[:, chunk]^T
//    (tcgen05.mma with A in tensor memory).  The 2C-channel intermediate never leaves the SM.
//  * D2's epilogue adds b2, accumulates GroupNorm-2 statistics and stores NCHW rows straight from registers: for a
//    fixed channel the 32 lanes of a warp are 32 consecutive pixels = one 128-byte line per store instruction.
//  * GroupNorm is a reduction over all pixels of an (episode, level), so the module is three passes:
//      A  x (fp32) -> bf16 smem tile (+ a bf16 copy of x in the workspace) -> conv1 -> statistics of GroupNorm-1 only
//      B  bf16 x (TMA) -> conv1 -> GN1 + LeakyReLU -> conv2 (+ b2) -> raw y2 to `out` + statistics of GroupNorm-2
//      C  out = LeakyReLU(GN2(out)) in place (streaming kernel in fusion_conv.cu)
//    DRAM: A reads 4C and writes 2C bytes per pixel, B reads 2C and writes 4C, C reads and writes 4C: 20C (1.84 GB at
//    16 x 22 400 pixels, C = 256) instead of 32C; weights from L2: 1.5 x 512 KB per 128 pixels instead of 512 KB per 64.
//
// Warp roles (the SM's issue arbiter favours the higher warp id, so the latency-critical roles sit on top).
// Pass A (18 warps): 0-7 activation producers (fp32 rows -> bf16, st.shared into the swizzled operand layout and
// st.global into the bf16 copy), 8-15 statistics epilogue (two warps per TMEM lane quadrant, half of each chunk's
// columns each), 16 TMA (weights), 17 MMA issuer + TMEM owner.  Pass B (19 warps): 0-7 output epilogue, 8-15 chunk
// epilogue (GN1 -> TMEM), 16 TMA (weights), 17 TMA (activations), 18 MMA issuer.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "fusion_internal.cuh"
#include "osd_common.cuh"
#include "osd_tc.cuh"

namespace osd {
namespace {
using namespace tc;

constexpr int kTileM = 128;             // pixels per tile (UMMA M, TMEM lanes)
constexpr int kChunk = 128;             // conv1 output channels per chunk (UMMA N of conv1, K of one conv2 block)
constexpr int kStageK = 64;             // K elements per weight stage (one 128-byte swizzle row of bf16)
// weight ring: 80 KB per CTA.  One CTA per tile: 5 stages of 16 KB (128 rows x 64 K).  CTA pairs (cta_group::2): every
// CTA holds HALF the rows of a stage (8 KB) -- 10 stages, and half the L2 -> SM weight stream per CTA.
template <bool TWO> struct Ring {
  static constexpr int kStages = TWO ? 10 : 5;
  static constexpr int kStageBytes = TWO ? 64 * 128 : 128 * 128;   // bytes of a stage in THIS CTA's shared memory
};
constexpr int kRingBytes = 5 * 128 * 128;
constexpr int kUmmaK = 16;
constexpr uint32_t kD2Col = 256;        // TMEM columns: D1/y1 chunk buffers at 0 and 128, D2 at 256..511

constexpr int kDefaultFusionMode = 0;    // see fusion_full_forward
// multicast-weights mode: CTAs per cluster.  Two CTAs save nothing (the L2 already merges concurrent requests for the
// same lines of up to ~4 SMs); eight halve the L2 -> SM weight traffic, which at ~75 GB/s per SM sits at the L2's
// throughput cap (~6300 B/clk for the whole chip)
constexpr int kMc = 8;
constexpr uint16_t kMcMask = (uint16_t)((1u << kMc) - 1u);
constexpr int kThreadsA = 32 * 18;
constexpr int kThreadsB = 32 * 19;

struct FLevel {
  const float* in;        // [B, C, hw] fp32 (pass A)
  __nv_bfloat16* xb;      // [B*C, pitch] bf16 copy (written by pass A, read through TMA by pass B)
  float* out;             // [B, C, hw] fp32 (pass B)
  int hw, pitch;
  int tiles_per_img;      // ceil(hw / 128)
  int tile_begin;
};

struct FArgs {
  int nl, B;
  int total_tiles;
  int store;              // pass B: write y2 to out
  int final_xform;        // pass B: apply GroupNorm-2 + LeakyReLU in the output epilogue (coef2), else raw y2 (+ b2)
  float slope;
  const float* bias1;     // [nl, B, 2C] folded conv1 bias (pass A statistics)
  const float2* coef1;    // [nl, B, 2C] GN1 (scale, shift incl. folded bias) (pass B)
  const float* b2;        // [C]
  const float2* coef2;    // [nl, B, C] GN2 (scale, shift) (pass B with final_xform)
  double* stats1;         // [nl, B, 32, 2] (pass A)
  double* stats2;         // [nl, B, 32, 2] (pass B, may be null)
  long long* prof;        // diagnosis (OSD_FUSION_PROF=1): per CTA 16 clock64 sums of the roles' wait / work phases
  FLevel lv[OSD_MAX_LEVELS];
};

struct XMaps {
  CUtensorMap m[OSD_MAX_LEVELS];
};

struct TileInfo {
  int level, img, px0, nvalid;
};

__device__ __forceinline__ TileInfo decode_tile(const FArgs& A, int tile) {
  if (tile >= A.total_tiles) {   // the odd tile out of the last CTA pair: no valid pixel, coordinates past level 0's rows
    TileInfo g;
    g.level = 0; g.img = 0; g.px0 = A.lv[0].tiles_per_img * kTileM; g.nvalid = 0;
    return g;
  }
  int li = 0;
#pragma unroll
  for (int k = 1; k < OSD_MAX_LEVELS; ++k)
    if (k < A.nl && tile >= A.lv[k].tile_begin) li = k;
  const int local = tile - A.lv[li].tile_begin;
  const int tpi = A.lv[li].tiles_per_img;
  TileInfo t;
  t.level = li;
  t.img = local / tpi;
  t.px0 = (local - t.img * tpi) * kTileM;
  t.nvalid = min(kTileM, A.lv[li].hw - t.px0);
  return t;
}

// shared-memory plan (offsets from the 1024-byte aligned base)
template <int C>
struct Smem {
  static constexpr uint32_t x_bytes = C * 256;                 // one activation tile: [C rows][128 px] bf16
  static constexpr uint32_t x_off = 0;
  static constexpr uint32_t w_off = 2 * x_bytes;
  static constexpr uint32_t aux_off = w_off + kRingBytes;
  // aux: pass A: bias1 [2][2C] floats; pass B: coef1 [2][2C] float2, b2 [C] floats, coef2 [2][C] float2
  static constexpr uint32_t aux_bytes = 2 * 2 * C * 8 + C * 4 + 2 * C * 8;
  static constexpr uint32_t bar_off = aux_off + aux_bytes;
  static constexpr uint32_t total = bar_off + 256 + 1024;       // barriers + alignment slack
};

constexpr int kMaxRingStages = 10;
struct Bars {
  uint32_t w_full, w_empty;      // [ring stages] each
  uint32_t x_full, x_empty;      // [2]
  uint32_t d1_full, d1_done;     // [2]: conv1 chunk accumulated / chunk epilogue finished (y1 written or D1 drained)
  uint32_t d2_full, d2_empty;    // [1]
  uint32_t tmem_slot;
};
__device__ __forceinline__ Bars make_bars(uint32_t base) {
  Bars b;
  b.w_full = base;
  b.w_empty = base + 8u * kMaxRingStages;
  b.x_full = base + 16u * kMaxRingStages;
  b.x_empty = b.x_full + 16u;
  b.d1_full = b.x_full + 32u;
  b.d1_done = b.x_full + 48u;
  b.d2_full = b.x_full + 64u;
  b.d2_empty = b.x_full + 72u;
  b.tmem_slot = b.x_full + 80u;
  return b;
}

// diagnosis: wait and add the cycles it took to *acc (acc == nullptr: plain wait)
__device__ __forceinline__ void mbar_wait_timed(uint32_t bar, uint32_t parity, long long* acc) {
  if (acc) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    *acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// The MMA warp runs its loops with all 32 lanes (warp-uniform values: the descriptor arithmetic stays in the uniform
// datapath, which is where UTCHMMA takes its operands from) and elects one lane only around the tcgen05 instructions.
// A single-lane loop (if (lane == 0) {...}) spent ~130 clk per MMA on address arithmetic and register moves -- more
// than the 64 clk the tensor core needs for it.

// a weight stage has been consumed by every MMA issued so far.  Pairs: one multicast commit of the leader reaches both
// CTAs' barriers.  Multicast weights (MC): each CTA's own MMAs read its copy of the stage, and the stage is refilled in
// BOTH CTAs by both TMA threads -- so each commit arrives on the "empty" barrier of both CTAs (count 2).
template <bool TWO, bool MC>
__device__ __forceinline__ void release_stage(uint32_t bar) {
  if (TWO) umma_commit_2sm(bar);
  else if (MC) umma_commit_mc(bar, kMcMask);
  else umma_commit(bar);
}

// conv1 block for one chunk: D1[buf] = x_tile . W1x[chunk rows]^T, K = C in stages of 64
template <int C, bool TWO, bool MC>
__device__ __forceinline__ void issue_conv1(const Bars& bar, uint32_t sX, uint32_t sW, uint32_t tmem_d, uint32_t& wc,
                                            long long* tw = nullptr) {
  constexpr int kStages = Ring<TWO>::kStages, kStageBytes = Ring<TWO>::kStageBytes;
  constexpr uint32_t idesc = make_idesc_ex(TWO ? 2 * kTileM : kTileM, kChunk, /*a_mn=*/1, /*b_mn=*/0);
  // A: MN-major SW128: 64-pixel atoms are x_bytes/2 apart (LBO), 8-channel groups 1024 B apart (SBO)
  const uint64_t adesc0 = make_smem_desc(sX, (uint32_t)C * 128u, 1024u);
  // B: K-major SW128: 8-row groups 1024 B apart (SBO), K advances by 32 B inside the swizzle row
  const uint64_t bdesc0 = make_smem_desc(sW, 16u, 1024u);
#pragma unroll
  for (int kc = 0; kc < C / kStageK; ++kc, ++wc) {
    const uint32_t s = wc % kStages, ph = (wc / kStages) & 1u;
    mbar_wait_timed(bar.w_full + 8u * s, ph, tw);
    tc_fence_after();
    if (elect_one()) {
      const uint64_t bs = bdesc0 + (uint64_t)(s * (kStageBytes >> 4));
#pragma unroll
      for (int k16 = 0; k16 < kStageK / kUmmaK; ++k16) {
        const uint32_t kgrp = (uint32_t)(kc * kStageK + k16 * kUmmaK) >> 3;
        if (TWO) umma_bf16_2sm(tmem_d, adesc0 + (uint64_t)(kgrp * 64u), bs + (uint64_t)(k16 * 2), idesc, (kc | k16) != 0 ? 1u : 0u);
        else umma_bf16(tmem_d, adesc0 + (uint64_t)(kgrp * 64u), bs + (uint64_t)(k16 * 2), idesc, (kc | k16) != 0 ? 1u : 0u);
      }
      release_stage<TWO, MC>(bar.w_empty + 8u * s);
    }
    __syncwarp();
  }
}

// conv2 block for one chunk: D2 (+)= y1[buf] (TMEM, 128 px x 128 ch bf16) . W2[:, chunk]^T
template <int C, bool TWO, bool MC>
__device__ __forceinline__ void issue_conv2(const Bars& bar, uint32_t sW, uint32_t tmem_base, uint32_t buf, bool first_chunk,
                                            uint32_t& wc, long long* tw = nullptr) {
  constexpr int kStages = Ring<TWO>::kStages, kStageBytes = Ring<TWO>::kStageBytes;
  constexpr int N2 = C < 128 ? C : 128;
  constexpr uint32_t idesc = make_idesc_ex(TWO ? 2 * kTileM : kTileM, N2, /*a_mn=*/0, /*b_mn=*/0);
  const uint64_t bdesc0 = make_smem_desc(sW, 16u, 1024u);
  const uint32_t acc0 = first_chunk ? 0u : 1u;
#pragma unroll
  for (int h = 0; h < (C + 127) / 128; ++h) {
#pragma unroll
    for (int kc2 = 0; kc2 < kChunk / kStageK; ++kc2, ++wc) {
      const uint32_t s = wc % kStages, ph = (wc / kStages) & 1u;
      mbar_wait_timed(bar.w_full + 8u * s, ph, tw);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bs = bdesc0 + (uint64_t)(s * (kStageBytes >> 4));
#pragma unroll
        for (int k16 = 0; k16 < kStageK / kUmmaK; ++k16) {
          // y1 channels [64 kc2, 64 kc2 + 64) of the chunk sit at columns [64 kc2, 64 kc2 + 32): each half was written in
          // place by the epilogue warp that read it; 16 bf16 = 8 columns per K step
          const uint32_t a_taddr = tmem_base + buf * (uint32_t)kChunk + (uint32_t)kc2 * 64u + (uint32_t)k16 * 8u;
          const uint32_t d_taddr = tmem_base + kD2Col + (uint32_t)h * 128u;
          const uint32_t acc = (kc2 | k16) != 0 ? 1u : acc0;
          if (TWO) umma_bf16_ts_2sm(d_taddr, a_taddr, bs + (uint64_t)(k16 * 2), idesc, acc);
          else umma_bf16_ts(d_taddr, a_taddr, bs + (uint64_t)(k16 * 2), idesc, acc);
        }
        release_stage<TWO, MC>(bar.w_empty + 8u * s);
      }
      __syncwarp();
    }
  }
}

// wait of a role with slack (producers, TMA): backs off so that the spin does not take issue slots from the epilogue
// warps that share its scheduler
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(40);
    if (++spins > kSpinLimit) __trap();
  }
}

template <bool TWO>
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  if (elect_one()) {
    if (TWO) umma_commit_2sm(bar);
    else umma_commit(bar);
  }
  __syncwarp();
}

// arrival of a consumer role on a barrier that gates the MMA issue: in a CTA pair that barrier is the leader's
template <bool TWO>
__device__ __forceinline__ void arrive_gate(uint32_t bar) {
  if (TWO) mbar_arrive_leader(bar);
  else mbar_arrive(bar);
}

// One weight stage.  Pairs: this CTA loads ITS half of the rows (rows_half each) into its own ring; the bytes of both
// halves are counted on the leader's full barrier (leader: arrive + expect_tx of both halves, follower: plain arrive).
template <bool TWO, bool MC>
__device__ __forceinline__ void load_weight_stage(const Bars& bar, uint32_t sW, const CUtensorMap* map, int col, int row,
                                                  uint32_t stage_tx_bytes, int rows_half, uint32_t rank, uint32_t& wc) {
  constexpr int kStages = Ring<TWO>::kStages, kStageBytes = Ring<TWO>::kStageBytes;
  const uint32_t s = wc % kStages, ph = (wc / kStages) & 1u;
  mbar_wait(bar.w_empty + 8u * s, ph ^ 1u);
  if (MC) {
    // every CTA of the cluster holds the WHOLE stage; this CTA fetches its 1/kMc of the rows once from L2 and the TMA
    // unit writes them into all the CTAs' rings (same offset), signalling each CTA's own "full" barrier
    const int rows_part = (int)(stage_tx_bytes / 128u) / kMc;
    mbar_expect_tx(bar.w_full + 8u * s, stage_tx_bytes);
    tma_load_2d_mc(sW + s * kStageBytes + rank * (uint32_t)(rows_part * 128), map, col, row + (int)rank * rows_part,
                   bar.w_full + 8u * s, kMcMask);
  } else if (TWO) {
    if (rank == 0) mbar_expect_tx(bar.w_full + 8u * s, stage_tx_bytes);
    else mbar_arrive_leader(bar.w_full + 8u * s);
    tma_load_2d_2sm(sW + s * kStageBytes, map, col, row + (int)rank * rows_half, bar.w_full + 8u * s);
  } else {
    mbar_expect_tx(bar.w_full + 8u * s, stage_tx_bytes);
    tma_load_2d(sW + s * kStageBytes, map, col, row, bar.w_full + 8u * s);
  }
  ++wc;
}

// pairs: before a CTA exits, every multicast completion addressed to its barriers must have landed -- wait for the last
// phase of each weight-stage "empty" barrier (wc = stages loaded) / activation "empty" barrier (it = tiles processed)
template <bool TWO>
__device__ __forceinline__ void drain_ring(const Bars& bar, uint32_t wc) {
  constexpr int kStages = Ring<TWO>::kStages;
  for (uint32_t s = 0; s < (uint32_t)kStages; ++s) {
    if (wc <= s) continue;
    const uint32_t uses = (wc - s + kStages - 1) / kStages;
    mbar_wait(bar.w_empty + 8u * s, (uses - 1) & 1u);
  }
}
__device__ __forceinline__ void drain_x(const Bars& bar, uint32_t it) {
  for (uint32_t b = 0; b < 2; ++b) {
    if (it <= b) continue;
    const uint32_t uses = (it - b + 1) / 2;
    mbar_wait(bar.x_empty + 8u * b, (uses - 1) & 1u);
  }
}

// Sum 32 per-lane values over the warp's lanes: afterwards v[0] of lane L holds the total of value L (31 shuffles
// instead of 160: at every step a lane hands the half it does not keep to its partner).
__device__ __forceinline__ void warp_transpose_reduce32(float (&v)[32], int lane) {
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

// ------------------------------------------------------------------------------------------------
// pass A: conv1 statistics (+ bf16 copy of x)
// ------------------------------------------------------------------------------------------------
template <int C, bool TWO, bool MC>
__global__ void __launch_bounds__(kThreadsA, 1)
fusion_stats1_kernel(const __grid_constant__ CUtensorMap tmap_w1, const FArgs A) {
  using S = Smem<C>;
  static_assert(!(TWO && MC), "pairs (cta_group::2) and multicast weights are alternatives");
  constexpr bool CL = TWO || MC;                               // launched as clusters of 2 CTAs
  constexpr int kStages = Ring<TWO>::kStages;
  constexpr uint32_t kCl = MC ? (uint32_t)kMc : (TWO ? 2u : 1u);   // CTAs per cluster; cluster c works on tiles c*kCl + rank
  const uint32_t rank = blockIdx.x % kCl;
  const int tile0 = (int)(blockIdx.x - rank);
  constexpr int C2 = 2 * C;
  constexpr int NCH = C2 / kChunk;          // conv1 chunks per tile
  constexpr int GS = C2 / 32;               // channels per GroupNorm-1 group
  constexpr int GPH = 64 / GS;              // groups per chunk half (the 64 columns one statistics warp reads)
  static_assert(NCH * GPH == 16, "a statistics warp owns 16 of the 32 groups");
  constexpr int kTmaWarp = 16, kMmaWarp = 17;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sX = smem_base + S::x_off, sW = smem_base + S::w_off;
  float* sBias = reinterpret_cast<float*>(gen_base + S::aux_off);   // [2][C2]
  const Bars bar = make_bars(smem_base + S::bar_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  constexpr uint32_t kPair = TWO ? 2u : 1u;   // arrivals on the leader's gate barriers come from both CTAs of a pair
  if (CL) cluster_sync_all();                  // both CTAs are resident before TMEM is allocated for the pair
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_w1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar.w_full + 8u * s, kPair);
      mbar_init(bar.w_empty + 8u * s, MC ? kMc : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar.x_full + 8u * s, 8 * kPair);    // producer warps
      mbar_init(bar.x_empty + 8u * s, 1);
      mbar_init(bar.d1_full + 8u * s, 1);
      mbar_init(bar.d1_done + 8u * s, 8 * kPair);   // statistics warps
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (TWO) tmem_alloc_2sm(bar.tmem_slot, 512);
    else tmem_alloc(bar.tmem_slot, 512);
  }
  tc_fence_before();
  if (CL) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (bar.tmem_slot - smem_base));

  if (warp == kTmaWarp) {
    // ===================== TMA: weight stages, in the order the MMA warp consumes them =====================
    if (lane == 0) {
      uint32_t wc = 0;
      for (int base = tile0; base < A.total_tiles; base += gridDim.x)
        for (int j = 0; j < NCH; ++j)
          for (int kc = 0; kc < C / kStageK; ++kc)
            load_weight_stage<TWO, MC>(bar, sW, &tmap_w1, kc * kStageK, j * kChunk, 128 * 128, 64, rank, wc);
      if (CL) drain_ring<TWO>(bar, wc);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (pairs: the leader CTA issues for both) =====================
    if (!TWO || rank == 0) {
      uint32_t wc = 0, g = 0, it = 0;
      long long pr[4] = {0, 0, 0, 0};   // x_full, w_full, d1_done, total
      long long* P = A.prof ? pr : nullptr;
      const long long tstart = clock64();
      for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it) {
        const uint32_t xb = it & 1u;
        mbar_wait_timed(bar.x_full + 8u * xb, (it >> 1) & 1u, P);
        tc_fence_after();
        for (int j = 0; j < NCH; ++j, ++g) {
          const uint32_t b = g & 1u, u = g >> 1;
          mbar_wait_timed(bar.d1_done + 8u * b, (u & 1u) ^ 1u, P ? P + 2 : nullptr);   // statistics warps drained the chunk two back
          tc_fence_after();
          issue_conv1<C, TWO, MC>(bar, sX + xb * S::x_bytes, sW, tmem_base + b * (uint32_t)kChunk, wc, P ? P + 1 : nullptr);
          commit_elect<TWO>(bar.d1_full + 8u * b);
        }
        commit_elect<TWO>(bar.x_empty + 8u * xb);
      }
      if (P && lane == 0) {
        pr[3] = clock64() - tstart;
        for (int i = 0; i < 4; ++i) A.prof[blockIdx.x * 16 + 10 + i] = pr[i];
      }
    }
  } else if (warp < 8) {
    // ===================== activation producers (warps 0-7) =====================
    const int pw = warp;
    const int chunk = lane & 15;       // 16-byte bf16 chunk = 8 pixels; 16 chunks = 128 pixels
    const int rsub = lane >> 4;        // row inside the pair this warp handles per round
    uint32_t it = 0;
    for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      const TileInfo t = decode_tile(A, base + (int)rank);
      const FLevel& L = A.lv[t.level];
      const int px = t.px0 + chunk * 8;
      const float* src = L.in + (size_t)t.img * C * L.hw + px;
      __nv_bfloat16* dstg = L.xb + (size_t)t.img * C * L.pitch + px;
      const bool vec_ok = ((L.hw & 3) == 0) && (px + 8 <= L.hw);
      const int nleft = L.hw - px;     // valid pixels from this chunk's start (may be <= 0)
      constexpr int kRounds = (C / 16) < 8 ? (C / 16) : 8;   // 16 rows per round, up to 128 rows per batch
#pragma unroll 1
      for (int kh = 0; kh < C; kh += 128) {
        float v[kRounds][8];
#pragma unroll
        for (int u = 0; u < kRounds; ++u) {
          const int k = kh + u * 16 + pw * 2 + rsub;
          const float* p = src + (size_t)k * L.hw;
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
        if (kh == 0) mbar_wait_relaxed(bar.x_empty + 8u * buf, ph ^ 1u);
#pragma unroll
        for (int u = 0; u < kRounds; ++u) {
          const int k = kh + u * 16 + pw * 2 + rsub;
          uint4 pk;
          pk.x = pack_bf16x2(v[u][0], v[u][1]);
          pk.y = pack_bf16x2(v[u][2], v[u][3]);
          pk.z = pack_bf16x2(v[u][4], v[u][5]);
          pk.w = pack_bf16x2(v[u][6], v[u][7]);
          const uint32_t dst = sX + buf * S::x_bytes + (uint32_t)(chunk >> 3) * (uint32_t)(C * 128) + (uint32_t)(k >> 3) * 1024u +
                               (uint32_t)(k & 7) * 128u + (uint32_t)(((chunk & 7) ^ (k & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          if (nleft > 0) *reinterpret_cast<uint4*>(dstg + (size_t)k * L.pitch) = pk;
        }
      }
      fence_proxy_async();   // generic-proxy stores -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) arrive_gate<TWO>(bar.x_full + 8u * buf);
    }
    if (TWO && warp == 0 && lane == 0) drain_x(bar, it);
  } else {
    // ===================== GroupNorm-1 statistics (warps 8-15; thread = pixel) =====================
    // warp = (lane quadrant q, column half hh): reads columns [64 hh, 64 hh + 64) of every chunk.  The per-pixel sums
    // stay in registers for the whole tile (16 groups x {sum, sum of squares}); one cross-lane reduction per tile.
    const int q = warp & 3, hh = (warp - 8) >> 2;
    const int tid = threadIdx.x - 256;
    uint32_t it = 0, g = 0;
    for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it) {
      const TileInfo t = decode_tile(A, base + (int)rank);
      const bool valid = (q * 32 + lane) < t.nvalid;
      const size_t plane = (size_t)t.level * A.B + t.img;
      float* sb = sBias + (it & 1u) * C2;
      for (int c = tid; c < C2; c += 256) sb[c] = __ldg(A.bias1 + plane * C2 + c);
      named_bar_sync(1, 256);
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll
      for (int j = 0; j < NCH; ++j, ++g) {
        const uint32_t b = g & 1u;
        mbar_wait(bar.d1_full + 8u * b, (g >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < 64; cb += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + b * (uint32_t)kChunk + (uint32_t)(hh * 64 + cb), r);
          tmem_ld_wait();
          if (cb == 32) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_gate<TWO>(bar.d1_done + 8u * b);   // the chunk buffer may be overwritten
          }
          const float4* bp = reinterpret_cast<const float4*>(sb + j * kChunk + hh * 64 + cb);
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 bv = bp[i4];
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = 4 * i4 + e;
              const int gi = j * GPH + (cb + i) / GS;   // thread-local group index, compile-time after unrolling
              const float y = __uint_as_float(r[i]) + bb[e];
              acc[2 * gi] += y;
              acc[2 * gi + 1] = fmaf(y, y, acc[2 * gi + 1]);
            }
          }
        }
      }
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      }
      warp_transpose_reduce32(acc, lane);
      {
        // lane L holds value L = (group gi = L >> 1, sum / sum of squares = L & 1)
        const int gi = lane >> 1;
        const int jj = gi / GPH, within = gi % GPH;
        const int group = (jj * kChunk + hh * 64) / GS + within;
        atomicAdd(A.stats1 + plane * 64 + 2 * group + (lane & 1), (double)acc[0]);
      }
    }
  }

  tc_fence_before();
  if (CL) cluster_sync_all();
  else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// pass A': GroupNorm-1 statistics WITHOUT running conv1 (+ the bf16 copy of x)
// ------------------------------------------------------------------------------------------------
// y_c(p) = w_c . x_p + b_c is linear in x, so over the pixels p of one (episode, level) plane
//     sum_p y_c   = w_c . sx + HW b_c                          sx = sum_p x_p            (C-vector)
//     sum_p y_c^2 = w_c^T G w_c + 2 b_c (w_c . sx) + HW b_c^2   G  = sum_p x_p x_p^T      (C x C Gram matrix)
// and summed over the channels of a GroupNorm group g:  sum_c w_c^T G w_c = <G, M_g>,  M_g = sum_{c in g} w_c w_c^T, which
// depends on the weights only (prepared once by the host: w1x_gram).  The Gram matrix is a GEMM of the activation tile
// with ITSELF: the [C, 128 px] tile the producers write for conv1's MN-major A operand is, read with the other
// descriptor, a K-major [C rows x 128 K] operand -- A and B of  G += X X^T  are the same shared memory.  Compared with
// pass A above: half the tensor work (C x C instead of 2C x C per pixel), NO weight stream from L2, and no per-tile
// epilogue -- the accumulator (C x C fp32 = all of tensor memory at C = 256) stays resident while a CTA walks the
// consecutive tiles of a plane and is written out once per (CTA, plane) segment.  The pass is then bound by its HBM
// traffic alone (read x fp32, write the bf16 copy).
constexpr int kThreadsG = 32 * 13;   // 8 producers, 4 drain warps, MMA issuer

struct GramArgs {
  int tile_first[kNumSMs + 1];   // CTA i owns the consecutive tiles [tile_first[i], tile_first[i + 1])
  int seg_first[kNumSMs];        // id of CTA i's first (CTA, plane) segment
  float* gseg;                   // [segments, C, C]
  float* rowsum;                 // [segments, C]  sum over the segment's pixels of the bf16-rounded x
  int* seg_plane;                // [segments]
};

__device__ __forceinline__ int tile_plane(const FArgs& A, int tile) {
  const TileInfo t = decode_tile(A, tile);
  return t.level * A.B + t.img;
}

template <int C>
__global__ void __launch_bounds__(kThreadsG, 1) fusion_gram_kernel(const FArgs A, const GramArgs Q) {
  using S = Smem<C>;
  static_assert(C == 128 || C == 256, "Gram statistics: C = 128 or 256");
  constexpr int MH = C / 128;               // M halves of the accumulator
  constexpr int kMmaWarp = 12;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sX = smem_base + S::x_off;
  const Bars bar = make_bars(smem_base + S::bar_off);
  const uint32_t acc_full = bar.d1_full, acc_empty = bar.d1_done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = Q.tile_first[blockIdx.x], t1 = Q.tile_first[blockIdx.x + 1];

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar.x_full + 8u * s, 8);      // producer warps
      mbar_init(bar.x_empty + 8u * s, 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);                  // drain warps
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(bar.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (bar.tmem_slot - smem_base));

  if (warp == kMmaWarp) {
    // ===================== MMA issuer: G (+)= X_tile X_tile^T =====================
    constexpr uint32_t idesc = make_idesc_ex(128, C, /*a_mn=*/0, /*b_mn=*/0);
    uint32_t it = 0, run = 0;
    bool first = true;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const uint32_t xb = it & 1u;
      mbar_wait(bar.x_full + 8u * xb, (it >> 1) & 1u);
      if (first && run > 0) mbar_wait(acc_empty, (run - 1) & 1u);   // the previous segment has been drained
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sXt = sX + xb * S::x_bytes;
#pragma unroll
        for (int atom = 0; atom < 2; ++atom) {          // 64-pixel halves of the tile = K blocks of one swizzle row
          const uint64_t b0 = make_smem_desc(sXt + (uint32_t)atom * (uint32_t)(C * 128), 16u, 1024u);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
            for (int mh = 0; mh < MH; ++mh) {
              const uint64_t a0 = make_smem_desc(sXt + (uint32_t)atom * (uint32_t)(C * 128) + (uint32_t)mh * (128u * 128u), 16u, 1024u);
              umma_bf16(tmem_base + (uint32_t)mh * 256u, a0 + (uint64_t)(k16 * 2), b0 + (uint64_t)(k16 * 2), idesc,
                        (first && atom == 0 && k16 == 0) ? 0u : 1u);
            }
          }
        }
        umma_commit(bar.x_empty + 8u * xb);
      }
      __syncwarp();
      const bool last = (tile + 1 == t1) || (tile_plane(A, tile + 1) != tile_plane(A, tile));
      if (last) {
        commit_elect<false>(acc_full);
        ++run;
      }
      first = last;
    }
  } else if (warp < 8) {
    // ===================== activation producers (as pass A) + per-row sums of the rounded values =====================
    const int pw = warp;
    const int chunk = lane & 15, rsub = lane >> 4;
    constexpr int kRounds = 8;
    float rs[(C / 128) * kRounds];
#pragma unroll
    for (int i = 0; i < (C / 128) * kRounds; ++i) rs[i] = 0.f;
    uint32_t it = 0;
    int seg = Q.seg_first[blockIdx.x];
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const uint32_t buf = it & 1u, ph = (it >> 1) & 1u;
      const TileInfo t = decode_tile(A, tile);
      const FLevel& L = A.lv[t.level];
      const int px = t.px0 + chunk * 8;
      const float* src = L.in + (size_t)t.img * C * L.hw + px;
      __nv_bfloat16* dstg = L.xb + (size_t)t.img * C * L.pitch + px;
      const bool vec_ok = ((L.hw & 3) == 0) && (px + 8 <= L.hw);
      const int nleft = L.hw - px;
#pragma unroll 1
      for (int kh = 0; kh < C; kh += 128) {
        float v[kRounds][8];
#pragma unroll
        for (int u = 0; u < kRounds; ++u) {
          const int k = kh + u * 16 + pw * 2 + rsub;
          const float* p = src + (size_t)k * L.hw;
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
        if (kh == 0) mbar_wait_relaxed(bar.x_empty + 8u * buf, ph ^ 1u);
#pragma unroll
        for (int u = 0; u < kRounds; ++u) {
          const int k = kh + u * 16 + pw * 2 + rsub;
          uint4 pk;
          pk.x = pack_bf16x2(v[u][0], v[u][1]);
          pk.y = pack_bf16x2(v[u][2], v[u][3]);
          pk.z = pack_bf16x2(v[u][4], v[u][5]);
          pk.w = pack_bf16x2(v[u][6], v[u][7]);
          const uint32_t dst = sX + buf * S::x_bytes + (uint32_t)(chunk >> 3) * (uint32_t)(C * 128) + (uint32_t)(k >> 3) * 1024u +
                               (uint32_t)(k & 7) * 128u + (uint32_t)(((chunk & 7) ^ (k & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
          if (nleft > 0) *reinterpret_cast<uint4*>(dstg + (size_t)k * L.pitch) = pk;
          // the values the tensor core sees: bf16 -> fp32 is a 16-bit shift
          const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
          float s8 = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) s8 += __uint_as_float(w[q] << 16) + __uint_as_float(w[q] & 0xffff0000u);
          rs[(kh >> 7) * kRounds + u] += s8;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar.x_full + 8u * buf);
      const bool last = (tile + 1 == t1) || (tile_plane(A, tile + 1) != t.level * A.B + t.img);
      if (last) {
#pragma unroll
        for (int i = 0; i < (C / 128) * kRounds; ++i) {
          float x = rs[i];
          x += __shfl_xor_sync(0xffffffffu, x, 8);
          x += __shfl_xor_sync(0xffffffffu, x, 4);
          x += __shfl_xor_sync(0xffffffffu, x, 2);
          x += __shfl_xor_sync(0xffffffffu, x, 1);
          if (chunk == 0) Q.rowsum[(size_t)seg * C + (i / kRounds) * 128 + (i % kRounds) * 16 + pw * 2 + rsub] = x;
          rs[i] = 0.f;
        }
        ++seg;
      }
    }
  } else {
    // ===================== drain warps (8-11; thread = accumulator row): one write per (CTA, plane) segment =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t run = 0;
    int seg = Q.seg_first[blockIdx.x];
    int tile = t0;
    while (tile < t1) {
      const int p = tile_plane(A, tile);
      int e = tile + 1;
      while (e < t1 && tile_plane(A, e) == p) ++e;
      mbar_wait(acc_full, run & 1u);
      tc_fence_after();
      float* gout = Q.gseg + (size_t)seg * C * C;
#pragma unroll
      for (int mh = 0; mh < MH; ++mh) {
        float4* orow = reinterpret_cast<float4*>(gout + (size_t)(mh * 128 + row) * C);
#pragma unroll 2
        for (int cb = 0; cb < C; cb += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)mh * 256u + (uint32_t)cb, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            orow[(cb >> 2) + i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                                              __uint_as_float(r[4 * i + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      if (warp == 8 && lane == 0) Q.seg_plane[seg] = p;
      tile = e;
      ++seg;
      ++run;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// <G, M_g> for every plane and group: thread = one (i, j) entry of the C x C index space with its 32 M_g values in
// registers; segments arrive in plane order (8 loads in flight), partial sums are flushed whenever the plane changes: warp
// transpose-reduce, then shared-memory atomics into a per-plane table (no block barrier inside the loop); one fp64
// atomic per (plane, group) and CTA at the end
constexpr int kMaxPlanes = 160;
template <int C>
__global__ void __launch_bounds__(256) fusion_gram_frob_kernel(const float* __restrict__ gseg, const int* __restrict__ seg_plane,
                                                                int nseg, int nplanes, const float* __restrict__ mg, double* stats1) {
  __shared__ float sacc[kMaxPlanes][32];
  __shared__ int splane[kNumSMs + kMaxPlanes];
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < nplanes * 32; i += 256) (&sacc[0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < nseg; i += 256) splane[i] = seg_plane[i];
  float m[32], acc[32];
#pragma unroll
  for (int g = 0; g < 32; ++g) {
    m[g] = __ldg(mg + (size_t)g * C * C + idx);
    acc[g] = 0.f;
  }
  __syncthreads();
  int cur = splane[0];
  auto flush = [&](int plane) {
    float t[32];
#pragma unroll
    for (int g = 0; g < 32; ++g) { t[g] = acc[g]; acc[g] = 0.f; }
    warp_transpose_reduce32(t, lane);
    atomicAdd(&sacc[plane][lane], t[0]);
  };
  constexpr int kPre = 8;
  for (int s0 = 0; s0 < nseg; s0 += kPre) {
    float v[kPre];
#pragma unroll
    for (int u = 0; u < kPre; ++u) v[u] = (s0 + u < nseg) ? __ldg(gseg + (size_t)(s0 + u) * C * C + idx) : 0.f;
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      if (s0 + u < nseg) {
        const int p = splane[s0 + u];
        if (p != cur) {
          flush(cur);
          cur = p;
        }
#pragma unroll
        for (int g = 0; g < 32; ++g) acc[g] = fmaf(v[u], m[g], acc[g]);
      }
    }
  }
  flush(cur);
  __syncthreads();
  for (int i = threadIdx.x; i < nplanes * 32; i += 256) {
    const float x = (&sacc[0][0])[i];
    if (x != 0.f) atomicAdd(stats1 + (size_t)(i >> 5) * 64 + 2 * (i & 31) + 1, (double)x);
  }
}

// the linear terms: sx of the plane, u = W1x sx, then per group  sum_c (u_c + HW b_c)  and  sum_c (2 b_c u_c + HW b_c^2)
template <int C>
__global__ void __launch_bounds__(256) fusion_gram_linear_kernel(const float* __restrict__ rowsum, const int* __restrict__ seg_plane,
                                                                  int nseg, const __nv_bfloat16* __restrict__ w1x,
                                                                  const float* __restrict__ bias_eff, FArgs A, double* stats1) {
  __shared__ float sx[C];
  constexpr int C2 = 2 * C, GS = C2 / 32;
  const int plane = blockIdx.x;
  const float hw = (float)A.lv[plane / A.B].hw;
  for (int k = threadIdx.x; k < C; k += 256) {
    float s = 0.f;
    for (int sgi = 0; sgi < nseg; ++sgi)
      if (seg_plane[sgi] == plane) s += rowsum[(size_t)sgi * C + k];
    sx[k] = s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C2; c += 256) {
    const __nv_bfloat162* wr = reinterpret_cast<const __nv_bfloat162*>(w1x + (size_t)c * C);
    float u = 0.f;
#pragma unroll 8
    for (int k2 = 0; k2 < C / 2; ++k2) {
      const float2 w = __bfloat1622float2(wr[k2]);
      u = fmaf(w.x, sx[2 * k2], u);
      u = fmaf(w.y, sx[2 * k2 + 1], u);
    }
    const float b = bias_eff[(size_t)plane * C2 + c];
    float s1 = fmaf(hw, b, u);
    float s2 = fmaf(2.f * b, u, hw * b * b);
#pragma unroll
    for (int o = GS / 2; o >= 1; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & (GS - 1)) == 0) {
      atomicAdd(stats1 + (size_t)plane * 64 + 2 * (c / GS), (double)s1);
      atomicAdd(stats1 + (size_t)plane * 64 + 2 * (c / GS) + 1, (double)s2);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pass B: conv1 -> GN1 + LeakyReLU -> conv2 (+ b2) [-> GN2 + LeakyReLU] -> out, GroupNorm-2 statistics
// ------------------------------------------------------------------------------------------------
// FINAL: the output epilogue applies GroupNorm-2 + LeakyReLU (coef2) and always stores; otherwise it writes the raw y2
// (+ b2) when A.store is set and accumulates the GroupNorm-2 statistics.  SLOPE01: 0 <= slope <= 1, where
// LeakyReLU(y) == max(y, slope * y) exactly.
template <bool SLOPE01>
__device__ __forceinline__ float lrelu(float y, float slope) {
  if (SLOPE01) return fmaxf(y, y * slope);
  return y > 0.f ? y : y * slope;
}

template <int C, bool FINAL, bool SLOPE01, bool TWO, bool MC>
__global__ void __launch_bounds__(kThreadsB, 1)
fusion_b2b_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                  const __grid_constant__ XMaps xmaps, const FArgs A) {
  using S = Smem<C>;
  static_assert(!(TWO && MC), "pairs (cta_group::2) and multicast weights are alternatives");
  constexpr bool CL = TWO || MC;
  constexpr int kStages = Ring<TWO>::kStages;
  constexpr int N2 = C < 128 ? C : 128;
  constexpr uint32_t kCl = MC ? (uint32_t)kMc : (TWO ? 2u : 1u);
  const uint32_t rank = blockIdx.x % kCl;
  const int tile0 = (int)(blockIdx.x - rank);
  constexpr int C2 = 2 * C;
  constexpr int NCH = C2 / kChunk;
  constexpr int GS2 = C / 32;               // channels per GroupNorm-2 group
  constexpr int kColsPerWarp = C / 2;       // output epilogue: two warps per lane quadrant, 16 groups each
  constexpr int kTmaW = 16, kTmaX = 17, kMmaWarp = 18;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sX = smem_base + S::x_off, sW = smem_base + S::w_off;
  float2* sCoef1 = reinterpret_cast<float2*>(gen_base + S::aux_off);            // [2][C2]
  float* sB2 = reinterpret_cast<float*>(gen_base + S::aux_off + 2 * C2 * 8);    // [C]
  float2* sCoef2 = reinterpret_cast<float2*>(gen_base + S::aux_off + 2 * C2 * 8 + C * 4);   // [2][C]
  const Bars bar = make_bars(smem_base + S::bar_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  constexpr uint32_t kPair = TWO ? 2u : 1u;
  if (CL) cluster_sync_all();
  if (warp == kTmaW && lane == 0) {
    tma_prefetch_desc(&tmap_w1);
    tma_prefetch_desc(&tmap_w2);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar.w_full + 8u * s, kPair);
      mbar_init(bar.w_empty + 8u * s, MC ? kMc : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar.x_full + 8u * s, kPair);        // (expect_tx) arrival of the TMA warp of each CTA
      mbar_init(bar.x_empty + 8u * s, 1);
      mbar_init(bar.d1_full + 8u * s, 1);
      mbar_init(bar.d1_done + 8u * s, 8 * kPair);   // chunk epilogue warps
    }
    mbar_init(bar.d2_full, 1);
    mbar_init(bar.d2_empty, 8 * kPair);             // output epilogue warps
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (TWO) tmem_alloc_2sm(bar.tmem_slot, 512);
    else tmem_alloc(bar.tmem_slot, 512);
  }
  tc_fence_before();
  if (CL) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (bar.tmem_slot - smem_base));

  // Per tile the tensor pipe runs   C1(0) C1(1) | C2(0) C1(2) | C2(1) C1(3) | C2(2) | C2(3)   (conv1 one chunk ahead of
  // conv2; the next tile's C1(0) C1(1) follow at once, which is the window in which the output epilogue drains D2).
  if (warp == kTmaW) {
    // ===================== TMA: weight stages in consumption order =====================
    if (lane == 0) {
      uint32_t wc = 0;
      for (int base = tile0; base < A.total_tiles; base += gridDim.x) {
        auto w1 = [&](int j) {
          for (int kc = 0; kc < C / kStageK; ++kc)
            load_weight_stage<TWO, MC>(bar, sW, &tmap_w1, kc * kStageK, j * kChunk, 128 * 128, 64, rank, wc);
        };
        // one CTA: the box is always 128 rows (rows past C read as zero); pairs: N2/2 rows per CTA
        auto w2 = [&](int j) {
          for (int h = 0; h < (C + 127) / 128; ++h)
            for (int kc2 = 0; kc2 < kChunk / kStageK; ++kc2)
              load_weight_stage<TWO, MC>(bar, sW, &tmap_w2, j * kChunk + kc2 * kStageK, h * 128, (CL ? N2 : 128) * 128,
                                         N2 / 2, rank, wc);
        };
        w1(0);
        if (NCH > 1) w1(1);
        for (int j = 0; j < NCH; ++j) {
          w2(j);
          if (j + 2 < NCH) w1(j + 2);
        }
      }
      if (CL) drain_ring<TWO>(bar, wc);
    }
  } else if (warp == kTmaX) {
    // ===================== TMA: activation tiles (bf16 copy written by pass A) =====================
    if (lane == 0) {
      for (int l = 0; l < A.nl; ++l) tma_prefetch_desc(&xmaps.m[l]);
      uint32_t it = 0;
      for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it) {
        const uint32_t buf = it & 1u;
        const TileInfo t = decode_tile(A, base + (int)rank);
        mbar_wait_relaxed(bar.x_empty + 8u * buf, ((it >> 1) & 1u) ^ 1u);
        // two boxes of 64 pixels x C rows: exactly the two MN-major swizzle atoms of the A operand
        const uint32_t dst = sX + buf * S::x_bytes;
        if (TWO) {
          if (rank == 0) mbar_expect_tx(bar.x_full + 8u * buf, 2 * S::x_bytes);   // this CTA's tile and the peer's
          else mbar_arrive_leader(bar.x_full + 8u * buf);
          tma_load_2d_2sm(dst, &xmaps.m[t.level], t.px0, t.img * C, bar.x_full + 8u * buf);
          tma_load_2d_2sm(dst + S::x_bytes / 2, &xmaps.m[t.level], t.px0 + 64, t.img * C, bar.x_full + 8u * buf);
        } else {
          mbar_expect_tx(bar.x_full + 8u * buf, S::x_bytes);
          tma_load_2d(dst, &xmaps.m[t.level], t.px0, t.img * C, bar.x_full + 8u * buf);
          tma_load_2d(dst + S::x_bytes / 2, &xmaps.m[t.level], t.px0 + 64, t.img * C, bar.x_full + 8u * buf);
        }
      }
      if (TWO) drain_x(bar, it);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (pairs: the leader CTA issues for both) =====================
    if (!TWO || rank == 0) {
      uint32_t wc = 0, g0 = 0, it = 0;
      long long pr[6] = {0, 0, 0, 0, 0, 0};   // x_full, w_full (conv1), d1_done, d2_empty, w_full (conv2), total
      long long* P = A.prof ? pr : nullptr;
      const long long tstart = clock64();
      for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it, g0 += NCH) {
        const uint32_t xb = it & 1u;
        const uint32_t sXt = sX + xb * S::x_bytes;
        mbar_wait_timed(bar.x_full + 8u * xb, (it >> 1) & 1u, P);
        tc_fence_after();
        // chunk g uses TMEM buffer g & 1.  The buffer's previous user is chunk g-2: its conv2 block was issued before
        // this conv1 block (the tensor pipe executes in issue order) and this thread waited for that chunk's epilogue
        // before issuing it, so no further wait is needed here.
        auto c1 = [&](int j) {
          const uint32_t b = (g0 + j) & 1u;
          issue_conv1<C, TWO, MC>(bar, sXt, sW, tmem_base + b * (uint32_t)kChunk, wc, P ? P + 1 : nullptr);
          commit_elect<TWO>(bar.d1_full + 8u * b);
          if (j == NCH - 1) commit_elect<TWO>(bar.x_empty + 8u * xb);   // the activation tile may be refilled
        };
        c1(0);
        if (NCH > 1) c1(1);
        for (int j = 0; j < NCH; ++j) {
          const uint32_t g = g0 + j, b = g & 1u;
          mbar_wait_timed(bar.d1_done + 8u * b, (g >> 1) & 1u, P ? P + 2 : nullptr);          // y1 chunk written to TMEM
          if (j == 0) mbar_wait_timed(bar.d2_empty, (it & 1u) ^ 1u, P ? P + 3 : nullptr);     // previous tile's output drained
          tc_fence_after();
          issue_conv2<C, TWO, MC>(bar, sW, tmem_base, b, j == 0, wc, P ? P + 4 : nullptr);
          if (j == NCH - 1) commit_elect<TWO>(bar.d2_full);
          if (j + 2 < NCH) c1(j + 2);
        }
      }
      if (P && lane == 0) {
        pr[5] = clock64() - tstart;
        for (int i = 0; i < 6; ++i) A.prof[blockIdx.x * 16 + i] = pr[i];
      }
    }
  } else if (warp >= 8) {
    // ===================== chunk epilogue (warps 8-15; thread = pixel): GN1 + LeakyReLU, y1 -> TMEM =====================
    // warp = (lane quadrant q, column half hh): channels [64 hh, 64 hh + 64) of the chunk, written back as bf16 into
    // columns [64 hh, 64 hh + 32) of the same buffer -- columns this warp alone reads, and has read before it writes
    const int q = warp & 3, hh = (warp - 8) >> 2;
    const int tid = threadIdx.x - 256;
    const float slope = A.slope;
    uint32_t it = 0, g = 0;
    long long e1wait = 0;
    long long* PE = (A.prof && warp == 8) ? &e1wait : nullptr;
    const long long e1start = clock64();
    for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it) {
      const TileInfo t = decode_tile(A, base + (int)rank);
      const size_t plane = (size_t)t.level * A.B + t.img;
      float2* sc = sCoef1 + (it & 1u) * C2;
      for (int c = tid; c < C2; c += 256) sc[c] = __ldg(A.coef1 + plane * C2 + c);
      named_bar_sync(1, 256);
#pragma unroll 1
      for (int j = 0; j < NCH; ++j, ++g) {
        const uint32_t b = g & 1u;
        mbar_wait_timed(bar.d1_full + 8u * b, (g >> 1) & 1u, PE);
        tc_fence_after();
        const uint32_t tbuf = tmem_base + ((uint32_t)(q * 32) << 16) + b * (uint32_t)kChunk + (uint32_t)(hh * 64);
#pragma unroll
        for (int cb = 0; cb < 64; cb += 32) {
          uint32_t r[32];
          tmem_ld32(tbuf + (uint32_t)cb, r);
          tmem_ld_wait();
          const float4* cf = reinterpret_cast<const float4*>(sc + j * kChunk + hh * 64 + cb);   // (scale, shift) x 2 channels
          uint32_t p[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 k = cf[i];
            const float a = lrelu<SLOPE01>(fmaf(__uint_as_float(r[2 * i]), k.x, k.y), slope);
            const float c = lrelu<SLOPE01>(fmaf(__uint_as_float(r[2 * i + 1]), k.z, k.w), slope);
            p[i] = pack_bf16x2(a, c);
          }
          tmem_st16(tbuf + (uint32_t)(cb >> 1), p);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_gate<TWO>(bar.d1_done + 8u * b);
      }
    }
    if (PE && lane == 0) {
      A.prof[blockIdx.x * 16 + 6] = e1wait;
      A.prof[blockIdx.x * 16 + 7] = clock64() - e1start;
    }
  } else {
    // ===================== output epilogue (warps 0-7; thread = pixel) =====================
    const int q = warp & 3, hh = warp >> 2;   // lane quadrant, channel half
    const float slope = A.slope;
    for (int c = threadIdx.x; c < C; c += 256) sB2[c] = __ldg(A.b2 + c);
    named_bar_sync(2, 256);
    uint32_t it = 0;
    long long e2wait = 0;
    long long* PE2 = (A.prof && warp == 0) ? &e2wait : nullptr;
    const long long e2start = clock64();
    for (int base = tile0; base < A.total_tiles; base += gridDim.x, ++it) {
      const TileInfo t = decode_tile(A, base + (int)rank);
      const FLevel& L = A.lv[t.level];
      const int pl = q * 32 + lane;
      const bool valid = pl < t.nvalid;
      const bool do_store = valid && (FINAL || A.store);
      const size_t plane = (size_t)t.level * A.B + t.img;
      float2* sc2 = sCoef2 + (it & 1u) * C;
      if (FINAL) {
        for (int c = threadIdx.x; c < C; c += 256) sc2[c] = __ldg(A.coef2 + plane * C + c);
        named_bar_sync(2, 256);
      }
      char* op = reinterpret_cast<char*>(L.out + ((size_t)t.img * C + (size_t)hh * kColsPerWarp) * L.hw + t.px0 + pl);
      const size_t ostride = (size_t)L.hw * 4;
      float acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      mbar_wait_timed(bar.d2_full, it & 1u, PE2);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < kColsPerWarp; cb += 32) {
        const int c0 = hh * kColsPerWarp + cb;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kD2Col + (uint32_t)c0, r);
        tmem_ld_wait();
        if (cb == kColsPerWarp - 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_gate<TWO>(bar.d2_empty);   // D2 may be overwritten by the next tile's conv2
        }
        const float4* bp = reinterpret_cast<const float4*>(sB2 + c0);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bv = bp[i4];
          const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = 4 * i4 + e;
            float y = __uint_as_float(r[i]) + bb[e];
            if (!FINAL) {
              const int gi = (cb + i) / GS2;   // thread-local group index (16 groups per warp)
              acc[2 * gi] += y;
              acc[2 * gi + 1] = fmaf(y, y, acc[2 * gi + 1]);
            } else {
              const float2 k = sc2[c0 + i];
              y = lrelu<SLOPE01>(fmaf(y, k.x, k.y), slope);
            }
            if (do_store) *reinterpret_cast<float*>(op) = y;
            op += ostride;
          }
        }
      }
      if (!FINAL && A.stats2) {
        if (!valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        }
        warp_transpose_reduce32(acc, lane);
        const int group = (hh * kColsPerWarp) / GS2 + (lane >> 1);
        atomicAdd(A.stats2 + plane * 64 + 2 * group + (lane & 1), (double)acc[0]);
      }
    }
    if (PE2 && lane == 0) {
      A.prof[blockIdx.x * 16 + 8] = e2wait;
      A.prof[blockIdx.x * 16 + 9] = clock64() - e2start;
    }
  }

  tc_fence_before();
  if (CL) cluster_sync_all();
  else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (TWO) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// launch with an optional cluster of 2 CTAs along x
template <typename... KArgs, typename... Args>
int launch_fused(void (*kernel)(KArgs...), int grid, int threads, size_t smem, int cluster, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  OSD_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
  return OSD_OK;
}

template <int C, bool TWO, bool MC>
int run_full(const osd_fusion_desc* d, const FusionWorkspace& ws, cudaStream_t stream) {
  using S = Smem<C>;
  constexpr bool CL = TWO || MC;
  static_assert(S::total <= 227 * 1024, "shared-memory plan exceeds 227 KB");
  constexpr int N2 = C < 128 ? C : 128;
  const int C2 = 2 * C, B = d->batch, nl = d->num_levels;
  const size_t per = (size_t)nl * B;
  OSD_CUDA(cudaMemsetAsync(ws.stats1, 0, sizeof(double) * per * 64, stream));
  OSD_CUDA(cudaMemsetAsync(ws.stats2, 0, sizeof(double) * per * 64, stream));
  timeline_mark("fusion_begin", stream);

  // weight boxes: one CTA per tile loads whole 128-row stages; a CTA of a pair loads its half of the rows
  CUtensorMap map1, map2;
  XMaps xm;
  memset(&xm, 0, sizeof(xm));
  int rc = make_bf16_map(d->w1x_bf16, C2, C, C, kStageK, MC ? 128 / kMc : (TWO ? 64 : 128), &map1);
  if (rc != OSD_OK) return rc;
  rc = make_bf16_map(d->w2_bf16, C, C2, C2, kStageK, MC ? N2 / kMc : (TWO ? N2 / 2 : 128), &map2);
  if (rc != OSD_OK) return rc;

  FArgs A{};
  A.nl = nl; A.B = B; A.slope = d->lrelu_slope;
  A.bias1 = ws.bias_eff; A.coef1 = ws.coef1; A.b2 = d->b2; A.coef2 = ws.coef2;
  A.stats1 = ws.stats1; A.stats2 = ws.stats2;
  int tiles = 0;
  size_t xb_off = 0;
  int32_t hw[OSD_MAX_LEVELS];
  float* outs[OSD_MAX_LEVELS];
  for (int l = 0; l < nl; ++l) {
    FLevel& L = A.lv[l];
    L.in = static_cast<const float*>(d->feat[l]);
    L.out = static_cast<float*>(d->out[l]);
    L.hw = d->hw[l];
    L.pitch = (int)fusion_xb_pitch(d->hw[l]);
    L.xb = static_cast<__nv_bfloat16*>(ws.xb) + xb_off;
    xb_off += (size_t)B * C * L.pitch;
    L.tiles_per_img = (d->hw[l] + kTileM - 1) / kTileM;
    L.tile_begin = tiles;
    tiles += B * L.tiles_per_img;
    hw[l] = d->hw[l];
    outs[l] = L.out;
    // columns = valid pixels (the pad up to the pitch reads as zero), rows = B*C
    rc = make_bf16_map(L.xb, (int64_t)B * C, L.hw, L.pitch, 64, C, &xm.m[l]);
    if (rc != OSD_OK) return rc;
  }
  A.total_tiles = tiles;
  if (tiles <= 0) return OSD_OK;
  // one CTA (or one CTA pair) per SM (pair of SMs), persistent over tiles (pairs of tiles)
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  auto kA = fusion_stats1_kernel<C, TWO, MC>;
  rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kA), S::total);
  if (rc != OSD_OK) return rc;
  if (CL) {
    // persistent clusters: every cluster of the grid must be resident at once (a cluster that has to wait for a free SM
    // pair would run its whole share of the tiles after the others have finished)
    constexpr int kCl = MC ? kMc : 2;
    int max_clusters = kNumSMs / kCl;
    cudaLaunchConfig_t q{};
    q.gridDim = dim3((unsigned)kNumSMs);
    q.blockDim = dim3((unsigned)kThreadsB);
    q.dynamicSmemBytes = S::total;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = kCl; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    q.attrs = qa; q.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, reinterpret_cast<const void*>(kA), &q) == cudaSuccess && n > 0)
      max_clusters = std::min(max_clusters, n);
    else
      (void)cudaGetLastError();
    grid = kCl * std::min((tiles + kCl - 1) / kCl, max_clusters);
  }
  static const bool prof_on = [] { const char* e = getenv("OSD_FUSION_PROF"); return e && e[0] == '1'; }();
  static long long* prof_dev = nullptr;
  if (prof_on) {
    if (!prof_dev) OSD_CUDA(cudaMalloc(&prof_dev, sizeof(long long) * 16 * kNumSMs));
    OSD_CUDA(cudaMemsetAsync(prof_dev, 0, sizeof(long long) * 16 * kNumSMs, stream));
    A.prof = prof_dev;
  }

  const bool s01 = d->lrelu_slope >= 0.f && d->lrelu_slope <= 1.f;
  auto kB = s01 ? fusion_b2b_kernel<C, false, true, TWO, MC> : fusion_b2b_kernel<C, false, false, TWO, MC>;
  auto kBfinal = s01 ? fusion_b2b_kernel<C, true, true, TWO, MC> : fusion_b2b_kernel<C, true, false, TWO, MC>;
  rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kB), S::total);
  if (rc != OSD_OK) return rc;

  // ---- pass A: GroupNorm-1 statistics (+ bf16 copy of the features)
  // opt-in (OSD_FUSION_GRAM=1): correct (parity-tested) but measured slower than running conv1 -- DESIGN section 6
  static const bool gram_off = [] { const char* e = getenv("OSD_FUSION_GRAM"); return !(e && e[0] == '1'); }();
  bool gram_done = false;
  if constexpr (C >= 128) {
    if (d->w1x_gram != nullptr && ws.gseg != nullptr && !gram_off) {
      // consecutive tile ranges per CTA (the accumulator follows a plane), (CTA, plane) segments numbered in tile order
      const int gg = std::min(tiles, kNumSMs);
      GramArgs Q{};
      auto plane_of = [&](int tile) {
        int li = 0;
        for (int k = 1; k < nl; ++k)
          if (tile >= A.lv[k].tile_begin) li = k;
        return li * B + (tile - A.lv[li].tile_begin) / A.lv[li].tiles_per_img;
      };
      int nseg = 0;
      for (int i = 0; i < gg; ++i) {
        Q.tile_first[i] = (int)((int64_t)i * tiles / gg);
        const int t1 = (int)((int64_t)(i + 1) * tiles / gg);
        Q.seg_first[i] = nseg;
        nseg += plane_of(t1 - 1) - plane_of(Q.tile_first[i]) + 1;
      }
      Q.tile_first[gg] = tiles;
      Q.gseg = ws.gseg; Q.rowsum = ws.rowsum; Q.seg_plane = ws.seg_plane;
      if (nseg > ws.max_segments) {
        set_error("osd_fusion: %d Gram segments exceed the workspace's %d", nseg, ws.max_segments);
        return OSD_ERR_WORKSPACE;
      }
      auto kG = fusion_gram_kernel<C>;
      rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kG), S::total);
      if (rc != OSD_OK) return rc;
      kG<<<gg, kThreadsG, S::total, stream>>>(A, Q);
      OSD_LAUNCH_CHECK("fusion_gram_kernel");
      timeline_mark("fusion_gram_kernel", stream);
      if (nl * B > kMaxPlanes) {
        set_error("osd_fusion: %d (level, episode) planes exceed the Gram statistics' %d; split the batch", nl * B, kMaxPlanes);
        return OSD_ERR_INVALID;
      }
      fusion_gram_frob_kernel<C><<<C * C / 256, 256, 0, stream>>>(ws.gseg, ws.seg_plane, nseg, nl * B, d->w1x_gram, ws.stats1);
      OSD_LAUNCH_CHECK("fusion_gram_frob_kernel");
      fusion_gram_linear_kernel<C><<<nl * B, 256, 0, stream>>>(ws.rowsum, ws.seg_plane, nseg,
                                                               static_cast<const __nv_bfloat16*>(d->w1x_bf16), ws.bias_eff, A, ws.stats1);
      OSD_LAUNCH_CHECK("fusion_gram_linear_kernel");
      timeline_mark("fusion_gram_stats", stream);
      gram_done = true;
    }
  }
  if (!gram_done) {
    rc = launch_fused(kA, grid, kThreadsA, S::total, MC ? kMc : (TWO ? 2 : 1), stream, map1, A);
    if (rc != OSD_OK) return rc;
    OSD_LAUNCH_CHECK("fusion_stats1_kernel");
    timeline_mark("fusion_stats1_kernel", stream);
  }
  rc = fusion_launch_gn_coef(nl, B, C2, d->gn_eps, ws.stats1, d->gn1_w, d->gn1_b, ws.bias_eff, ws.coef1, hw, stream);
  if (rc != OSD_OK) return rc;

  // ---- pass B: fused conv1 -> GN1 -> LeakyReLU -> conv2; raw y2 (+ b2) to out, GroupNorm-2 statistics
  static const bool recompute = [] { const char* e = getenv("OSD_FUSION_RECOMPUTE"); return e && e[0] == '1'; }();
  A.store = recompute ? 0 : 1;
  A.final_xform = 0;
  rc = launch_fused(kB, grid, kThreadsB, S::total, MC ? kMc : (TWO ? 2 : 1), stream, map1, map2, xm, A);
  if (rc != OSD_OK) return rc;
  OSD_LAUNCH_CHECK("fusion_b2b_kernel");
  timeline_mark("fusion_b2b_kernel", stream);
  rc = fusion_launch_gn_coef(nl, B, C, d->gn_eps, ws.stats2, d->gn2_w, d->gn2_b, nullptr, ws.coef2, hw, stream);
  if (rc != OSD_OK) return rc;

  if (recompute) {
    // ---- pass C': the fused chain once more with GroupNorm-2 + LeakyReLU in the output epilogue (no y2 round trip)
    A.store = 1;
    A.final_xform = 1;
    A.stats2 = nullptr;
    rc = ensure_dynamic_smem(reinterpret_cast<const void*>(kBfinal), S::total);
    if (rc != OSD_OK) return rc;
    rc = launch_fused(kBfinal, grid, kThreadsB, S::total, MC ? kMc : (TWO ? 2 : 1), stream, map1, map2, xm, A);
    if (rc != OSD_OK) return rc;
    OSD_LAUNCH_CHECK("fusion_b2b_kernel");
    timeline_mark("fusion_b2b_kernel(final)", stream);
    return OSD_OK;
  }
  // ---- pass C: GroupNorm-2 + LeakyReLU in place
  rc = fusion_launch_gn_lrelu(nl, B, C, d->lrelu_slope, outs, ws.coef2, hw, stream);
  timeline_mark("fusion_gn_lrelu_kernel", stream);
  if (prof_on && rc == OSD_OK) {
    static long long host[16 * kNumSMs];
    OSD_CUDA(cudaStreamSynchronize(stream));
    OSD_CUDA(cudaMemcpy(host, prof_dev, sizeof(host), cudaMemcpyDeviceToHost));
    const char* names[14] = {"B.mma x_full", "B.mma w_full(c1)", "B.mma d1_done", "B.mma d2_empty", "B.mma w_full(c2)", "B.mma total",
                             "B.e1 wait", "B.e1 total", "B.e2 wait", "B.e2 total", "A.mma x_full", "A.mma w_full", "A.mma d1_done",
                             "A.mma total"};
    const int step = TWO ? 2 : 1;   // pairs: only the leader CTAs issue MMAs
    for (int i = 0; i < 14; ++i) {
      double sum = 0, mx = 0;
      int n = 0;
      const bool mma_row = i < 6 || i >= 10;
      for (int c = 0; c < grid; c += (mma_row ? step : 1), ++n) { sum += (double)host[c * 16 + i]; mx = std::max(mx, (double)host[c * 16 + i]); }
      fprintf(stderr, "[fusion prof] %-18s mean %10.0f clk  max %10.0f clk\n", names[i], sum / n, mx);
    }
  }
  return rc;
}

}  // namespace

int fusion_full_forward(const osd_fusion_desc* d, const FusionWorkspace& ws, cudaStream_t stream) {
  // OSD_FUSION_MODE: "single" = one CTA per tile; "mc" = clusters of 8 CTAs, one tile each, every weight stage
  // fetched once from L2 (an eighth of its rows per CTA) and multicast into all eight rings; "pair" (or OSD_FUSION_2CTA=1) = CTA pairs on tcgen05
  // cta_group::2 (M = 256 across two SMs)
  static const int mode = [] {
    const char* e = getenv("OSD_FUSION_MODE");
    const char* p2 = getenv("OSD_FUSION_2CTA");
    if (e && !strcmp(e, "pair")) return 1;
    if (e && !strcmp(e, "mc")) return 2;
    if (e && !strcmp(e, "single")) return 0;
    if (p2 && p2[0] == '1') return 1;
    return kDefaultFusionMode;
  }();
#define OSD_FUSION_DISPATCH(CH)                                         \
  case CH:                                                              \
    if (mode == 1) return run_full<CH, true, false>(d, ws, stream);     \
    if (mode == 2) return run_full<CH, false, true>(d, ws, stream);     \
    return run_full<CH, false, false>(d, ws, stream);
  switch (d->channels) {
    OSD_FUSION_DISPATCH(64)
    OSD_FUSION_DISPATCH(128)
    OSD_FUSION_DISPATCH(256)
  }
#undef OSD_FUSION_DISPATCH
  set_error("osd_fusion: channels must be 64, 128 or 256 (got %d)", d->channels);
  return OSD_ERR_INVALID;
}

}  // namespace osd
