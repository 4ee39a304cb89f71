// Batched greedy NMS for sm_100a: per-episode sort -> upper-triangle IoU bitmask -> warp-cooperative
// sweep -> emit.  Everything stays on the device (no host round-trip, cf. the D2H mask copy of
// maskrcnn_benchmark/csrc/cuda/nms.cu:100-123 in the reference), one launch covers all episodes, and
// results are bit-identical to maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp:5-64.
//
// Bit-exactness notes
//  * every IoU operation is an explicit round-to-nearest intrinsic (__fsub_rn, __fadd_rn, __fmul_rn,
//    __fdiv_rn): nvcc may not contract them into FMAs, so they round exactly like the x86-64 build of
//    nms_cpu.cpp (which has no FMA either).
//  * the common case avoids the division: with u = area_i + area_j - inter > 0 and thr > 0,
//    fl(inter/u) >= thr is decided by comparing inter against thr*(1 +- 2^-19)*u; only pairs inside
//    that 4e-6-wide band take the IEEE division.  Episodes containing non-finite or inverted boxes
//    (u may be <= 0) take the division for every pair.
//  * visiting order is (score desc, index asc): the canonical stand-in for ATen's unstable sort at
//    nms_cpu.cpp:24; identical whenever scores are pairwise distinct.
#include <cfloat>
#include <climits>
#include <cmath>

#include <cstdlib>

#include "osd_common.cuh"
#include "osd_device_utils.cuh"

namespace osd {

namespace {

constexpr int kChunk = 2048;         // keys per sort chunk (one CTA each)
constexpr int kChunkThreads = 512;
constexpr int kMergeStage = 6 * kChunk;   // keys staged per window by the merge kernel (96 KB)
constexpr int kMergeThreads = 1024;
constexpr int kMergeKeys = kChunk / kMergeThreads;   // keys of the own chunk per thread
constexpr int kMaskRows = 128;       // rows per mask tile (one thread per row)
constexpr int kMaskCols = 128;       // columns per mask tile (2 words of 64)
constexpr int kSweepThreads = 512;
constexpr int kSweepPre = 8;         // rows per thread whose column words the sweep prefetches (small passes)
constexpr int kSweepDepth = 4;       // columns in flight in the sweep's cp.async ring
constexpr int kEdgeCap = 16384;      // suppression edges per episode handled by the sparse sweep (128 KB)
constexpr int kEdgeRounds = 48;      // relaxation rounds before the sparse sweep gives up

typedef unsigned long long u64;

struct IouTest {
  float thr, thr_lo, thr_hi;
  float c_lo;   // (thr / (1 + thr)) * (1 - 2^-19): inter < c_lo * (area_a + area_b)  =>  IoU < thr for certain
  int strict;
  int force_exact;
};

__device__ __forceinline__ uint32_t score_key_desc(float s) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone in float order
  return ~u;                                       // larger score -> smaller key
}

__device__ __forceinline__ int episode_count(const CandLayout& L, int e) {
  if (L.seg) return (int)(L.seg[e + 1] - L.seg[e]);
  int n = 0;
  for (int l = 0; l < L.nl; ++l) n += L.level_count[e * L.nl + l];
  return n;
}

// compact candidate index -> row in L.boxes / L.scores
__device__ __forceinline__ int64_t cand_row(const CandLayout& L, int e, int idx, const int* lvl_prefix) {
  if (L.seg) return L.seg[e] + idx;
  int l = 0;
#pragma unroll
  for (int k = 1; k < OSD_MAX_LEVELS; ++k)
    if (k < L.nl && idx >= lvl_prefix[k]) l = k;
  return (int64_t)e * L.cap + L.slot[l] + (idx - lvl_prefix[l]);
}

// exclusive prefix of the per-level candidate counts of episode e (mode B); contains a barrier
__device__ __forceinline__ void fill_level_prefix(const CandLayout& L, int e, int* lvl_prefix) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < OSD_MAX_LEVELS; ++l) {
      lvl_prefix[l] = acc;
      if (!L.seg && l < L.nl) acc += L.level_count[e * L.nl + l];
    }
    lvl_prefix[OSD_MAX_LEVELS] = acc;
  }
  __syncthreads();
}

__device__ __forceinline__ float box_area(const float4& b) {
  // nms_cpu.cpp:22  (x2 - x1 + 1) * (y2 - y1 + 1), each op rounded
  return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

__device__ __forceinline__ bool box_is_regular(const float4& b, float area) {
  return isfinite(b.x) && isfinite(b.y) && isfinite(b.z) && isfinite(b.w) && b.z >= b.x && b.w >= b.y &&
         isfinite(area);
}

// ------------------------------------------------------------------------------------------------
// 1. sort: (a) every 2048-key chunk of an episode is sorted by its own CTA (bitonic network in shared memory,
//    64-bit keys = score key << 32 | candidate index, so keys are unique), (b) every key finds its final
//    position as  own rank + sum over the other chunks of #keys smaller  (branch-free binary searches) and
//    scatters its box there.  E x ceil(n/2048) CTAs per kernel: the whole chip works on the sort.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_sorted(const CandLayout& L, const NmsWorkspace& W, int e, int pos,
                                             int idx, const int* lvl_prefix, bool& regular) {
  int64_t row = cand_row(L, e, idx, lvl_prefix);
  float4 b = L.boxes[row];
  float a = box_area(b);
  size_t o = (size_t)e * W.NP + pos;
  W.sbox[o] = b;
  W.sarea[o] = a;
  W.sscore[o] = L.scores[row];
  W.sidx[o] = idx;
  regular = regular && box_is_regular(b, a);
}

// Bitonic network over the 2048 keys of a chunk; thread t holds elements 4t..4t+3 in registers.  Exchange
// distances 1-2 stay inside the thread, 4-64 are warp shuffles, only distances >= 128 go through shared memory
// (10 of the 66 stages).
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int lane_mask) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, lane_mask);
  const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), lane_mask);
  return ((u64)hi << 32) | lo;
}

__global__ void __launch_bounds__(kChunkThreads) nms_chunk_sort_kernel(CandLayout L, NmsWorkspace W) {
  __shared__ u64 sk[kChunk];
  __shared__ int lvl_prefix[OSD_MAX_LEVELS + 1];
  const int e = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  fill_level_prefix(L, e, lvl_prefix);
  const int n = min(episode_count(L, e), W.NP);
  if (c == 0 && tid == 0) {
    W.n[e] = n;
    W.flags[e] = 1;
    W.done[e] = 0;
    W.kcount[e] = 0;
    W.ecount[e] = 0;
    if (e == 0) {
      W.sched[0] = 0;  // episodes finished
#pragma unroll
      for (int k = 1; k < 8; ++k) W.sched[k] = 0;  // mask tile counters of the passes
    }
  }
  const int base = c * kChunk;
  if (base >= n) return;
  const int len = min(kChunk, n - base);
  u64 r[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int i = 4 * tid + s;
    u64 k = ~0ull;
    if (i < len) {
      const float sc = L.scores[cand_row(L, e, base + i, lvl_prefix)];
      k = ((u64)score_key_desc(sc) << 32) | (uint32_t)(base + i);
    }
    r[s] = k;
  }
  for (int k = 2; k <= kChunk; k <<= 1) {
    int j = k >> 1;
    if (j >= 128) {
      // wide exchanges through shared memory
#pragma unroll
      for (int s = 0; s < 4; ++s) sk[4 * tid + s] = r[s];
      __syncthreads();
      for (; j >= 128; j >>= 1) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int t = tid + u * kChunkThreads;
          const int i = 2 * t - (t & (j - 1));
          const int l = i + j;
          const u64 a = sk[i], b = sk[l];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            sk[i] = b;
            sk[l] = a;
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) r[s] = sk[4 * tid + s];
      __syncthreads();  // sk is rewritten by the next wide phase
    }
    for (; j >= 4; j >>= 1) {
      // partner element lives in lane ^ (j/4), same register slot
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int i = 4 * tid + s;
        const u64 other = shfl_xor_u64(r[s], j >> 2);
        const bool up = (i & k) == 0;
        const bool lower = (i & j) == 0;
        const bool take_min = (lower == up);
        // keys are unique: keep mine iff it is the one this position wants
        r[s] = ((r[s] < other) == take_min) ? r[s] : other;
      }
    }
    // distances 2 and 1 stay inside the thread (static register indices)
    if (k >= 4) {
      const bool up = ((4 * tid) & k) == 0;
      if ((r[0] > r[2]) == up) { const u64 t0 = r[0]; r[0] = r[2]; r[2] = t0; }
      if ((r[1] > r[3]) == up) { const u64 t1 = r[1]; r[1] = r[3]; r[3] = t1; }
    }
    {
      const bool up01 = ((4 * tid) & k) == 0;      // element 4t   (bit 1 clear)
      const bool up23 = ((4 * tid + 2) & k) == 0;  // element 4t+2
      if ((r[0] > r[1]) == up01) { const u64 t0 = r[0]; r[0] = r[1]; r[1] = t0; }
      if ((r[2] > r[3]) == up23) { const u64 t1 = r[2]; r[2] = r[3]; r[3] = t1; }
    }
  }
  u64* out = W.sortkeys + (size_t)e * W.NP + base;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int i = 4 * tid + s;
    if (i < len) out[i] = r[s];
  }
}

// lower_bound (number of keys < key) in a sorted run of `len2` <= kChunk keys; fixed 12 probes, no branches
__device__ __forceinline__ int count_less(const u64* __restrict__ run, int len2, u64 key) {
  int pos = 0;
#pragma unroll
  for (int step = kChunk; step >= 1; step >>= 1) {
    const int probe = pos + step;
    if (probe <= len2 && run[probe - 1] < key) pos = probe;
  }
  return pos;
}

__global__ void __launch_bounds__(kMergeThreads) nms_merge_kernel(CandLayout L, NmsWorkspace W) {
  extern __shared__ u64 staged[];  // up to kMergeStage sorted keys of the other chunks
  __shared__ int lvl_prefix[OSD_MAX_LEVELS + 1];
  const int e = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  const int n = W.n[e];
  const int base = c * kChunk;
  if (base >= n) return;
  fill_level_prefix(L, e, lvl_prefix);
  const int len = min(kChunk, n - base);
  const u64* keys = W.sortkeys + (size_t)e * W.NP;
  // own keys (kMergeKeys per thread) and their running ranks
  u64 mine[kMergeKeys];
  int rank[kMergeKeys];
#pragma unroll
  for (int s = 0; s < kMergeKeys; ++s) {
    const int i = tid + s * kMergeThreads;
    mine[s] = (i < len) ? keys[base + i] : ~0ull;
    rank[s] = i;
  }
  // walk the episode's keys in windows that fit in shared memory; every window holds whole chunks
  for (int w0 = 0; w0 < n; w0 += kMergeStage) {
    const int wn = min(kMergeStage, n - w0);
    __syncthreads();
    for (int i = tid; i < wn; i += 4 * kMergeThreads) {  // 4 independent loads in flight
      u64 t4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i + u * kMergeThreads;
        t4[u] = idx < wn ? keys[w0 + idx] : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = i + u * kMergeThreads;
        if (idx < wn) staged[idx] = t4[u];
      }
    }
    __syncthreads();
    for (int c2off = 0; c2off < wn; c2off += kChunk) {
      if (w0 + c2off == base) continue;  // own chunk
      const int len2 = min(kChunk, wn - c2off);
#pragma unroll
      for (int s = 0; s < kMergeKeys; ++s) rank[s] += count_less(staged + c2off, len2, mine[s]);
    }
  }
  bool regular = true;
#pragma unroll
  for (int s = 0; s < kMergeKeys; ++s) {
    const int i = tid + s * kMergeThreads;
    if (i < len) write_sorted(L, W, e, rank[s], (int)(mine[s] & 0xffffffffu), lvl_prefix, regular);
  }
  if (!regular) atomicAnd(&W.flags[e], 0);
}

// ------------------------------------------------------------------------------------------------
// 2. IoU bitmask.  Persistent CTAs walk (row block of 128) x (column group of 256) tiles of all episodes;
//    thread = one row, 64 pairs per 64-bit word, column boxes broadcast from shared memory.
//    The inner loop is branch-free and builds one mask per word -- the pairs a conservative, division-free test
//    cannot rule out (see pair_bit); only those take the reference's full arithmetic afterwards.
// ------------------------------------------------------------------------------------------------
struct PairTerms {
  float inter, uni;
};

__device__ __forceinline__ PairTerms pair_terms(const float4& a, float aa, const float4& b, float ba) {
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.0f), 0.0f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.0f), 0.0f);
  PairTerms t;
  t.inter = __fmul_rn(w, h);
  t.uni = __fsub_rn(__fadd_rn(aa, ba), t.inter);
  return t;
}

// std::max / std::min semantics of nms_cpu.cpp:51-57, `a` is the earlier (suppressing) box
__device__ __forceinline__ bool iou_hit_exact(const float4& a, float aa, const float4& b, float ba,
                                              const IouTest& T) {
  float xx1 = (a.x < b.x) ? b.x : a.x, yy1 = (a.y < b.y) ? b.y : a.y;
  float xx2 = (b.z < a.z) ? b.z : a.z, yy2 = (b.w < a.w) ? b.w : a.w;
  float w = __fadd_rn(__fsub_rn(xx2, xx1), 1.0f);
  w = (0.0f < w) ? w : 0.0f;
  float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.0f);
  h = (0.0f < h) ? h : 0.0f;
  float inter = __fmul_rn(w, h);
  float u = __fsub_rn(__fadd_rn(aa, ba), inter);
  float q = __fdiv_rn(inter, u);
  return T.strict ? (q > T.thr) : (q >= T.thr);
}

// Conservative filter, one mask per word.  IoU >= thr  <=>  inter >= (thr / (1 + thr)) * (area_a + area_b) in exact
// arithmetic; the reference's three roundings (the sum, the subtraction, the division) move that boundary by less than
// 2^-22 relative, so  inter < c_lo * fl(area_a + area_b)  with c_lo = thr/(1+thr) * (1 - 2^-19) proves "not suppressed"
// (inter and the sum are computed with the reference's own operations).  Everything else -- the true hits and the
// pairs within 2^-19 of the threshold, well under 1 % of the pairs -- is decided by the reference's arithmetic
// including the IEEE division.  The compare is an integer subtraction whose sign bit is funnel-shifted into the mask
// (every operand is a non-negative finite float in a regular episode, so float order = bit-pattern order); bits
// arrive MSB-first and are reversed per 32 pairs.  13 + 3 instructions per pair instead of 13 + 7 for two masks.
__device__ __forceinline__ void pair_bit(const float4& rb, float ra, const float4& cb, float ca, const IouTest& T,
                                         uint32_t& sup) {
  const float xx1 = fmaxf(rb.x, cb.x), yy1 = fmaxf(rb.y, cb.y);
  const float xx2 = fminf(rb.z, cb.z), yy2 = fminf(rb.w, cb.w);
  const float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.0f), 0.0f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.0f), 0.0f);
  const float inter = __fmul_rn(w, h);
  const float sum = __fadd_rn(ra, ca);
  const int d = __float_as_int(__fmul_rn(T.c_lo, sum)) - __float_as_int(inter) - 1;   // < 0  <=>  inter >= c_lo * sum
  sup = __funnelshift_l((uint32_t)d, sup, 1);
}

__device__ __forceinline__ u64 row_word_fast(const float4& rb, float ra, const float4* __restrict__ cb,
                                             const float* __restrict__ ca, const IouTest& T) {
  uint32_t sup_lo = 0, sup_hi = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) pair_bit(rb, ra, cb[j], ca[j], T, sup_lo);
#pragma unroll
  for (int j = 0; j < 32; ++j) pair_bit(rb, ra, cb[32 + j], ca[32 + j], T, sup_hi);
  u64 cand = ((u64)__brev(sup_hi) << 32) | __brev(sup_lo);
  u64 sub = 0ull;
  while (cand) {  // the candidates: decide with the reference's arithmetic (IEEE division)
    const int j = __ffsll((long long)cand) - 1;
    cand &= cand - 1ull;
    const PairTerms t = pair_terms(rb, ra, cb[j], ca[j]);
    const float q = __fdiv_rn(t.inter, t.uni);
    if (T.strict ? (q > T.thr) : (q >= T.thr)) sub |= (1ull << j);
  }
  return sub;
}

struct MaskArgs {
  int row_begin, row_end;  // rows (visiting positions) this launch covers, multiples of 64
  int col_begin, col_end;  // columns this launch covers, multiples of 64
  int tiles_r, tiles_c;    // tile grid per episode
  int pass;                // which tile counter to use
  int compact_rows;        // rows [0, compact_rows) (a multiple of the tile height) have FINAL kept bits from earlier passes:
                           // only kept boxes can suppress, so their row tiles walk W.klist instead of every row
  IouTest test;
};

__global__ void __launch_bounds__(kMaskRows) nms_mask_kernel(NmsWorkspace W, MaskArgs A) {
  __shared__ float4 cbox[kMaskCols];
  __shared__ float carea[kMaskCols];
  __shared__ int s_tile;
  if (W.sched[0] >= W.E) return;  // every episode already finished (early exit): nothing to do
  const int total = W.E * A.tiles_r * A.tiles_c;
  int* counter = W.sched + 1 + A.pass;
  while (true) {
    // dynamic tile scheduler: tiles differ in cost (diagonal / skipped / full), CTAs pull the next one
    __syncthreads();  // the previous tile's readers are done with cbox and s_tile
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const int t = s_tile;
    if (t >= total) break;
    // episode fastest: concurrently running CTAs work on tiles of equal cost
    const int e = t % W.E;
    const int rc = t / W.E;
    const int rbi = rc / A.tiles_c, cgi = rc - rbi * A.tiles_c;
    if (W.done[e]) continue;
    const int n = W.n[e];
    const int row0 = A.row_begin + rbi * kMaskRows;
    const int col0 = A.col_begin + cgi * kMaskCols;
    const int row_end = min(n, A.row_end);
    const int col_end = min(n, A.col_end);
    if (row0 >= row_end || col0 >= col_end) continue;
    if (col0 + kMaskCols <= row0) continue;  // every column precedes every row of this tile
    // compacted row tile: its 128 threads take the rbi-th group of 128 KEPT rows (the list is ascending, so the entries
    // below compact_rows are a prefix of it); a group past them has nothing to do
    const bool compact = row0 < A.compact_rows;
    const int32_t* kl = W.klist + (size_t)e * W.NP;
    int kc = 0;
    if (compact) {
      kc = W.kcount[e];
      const int k0 = rbi * kMaskRows;
      if (k0 >= kc || kl[k0] >= A.compact_rows) continue;
    }
    const float4* sbox = W.sbox + (size_t)e * W.NP;
    const float* sarea = W.sarea + (size_t)e * W.NP;
    const int ncols = min(kMaskCols, col_end - col0);
    for (int c = threadIdx.x; c < kMaskCols; c += kMaskRows) {
      // columns past the end are padded with the first column (finite, regular; their bits are masked off)
      const int cc = col0 + (c < ncols ? c : 0);
      cbox[c] = sbox[cc];
      carea[c] = sarea[cc];
    }
    __syncthreads();
    int i = row0 + threadIdx.x;
    if (compact) {
      const int k = rbi * kMaskRows + (int)threadIdx.x;
      i = k < kc ? kl[k] : INT_MAX;
      if (i >= A.compact_rows) continue;
    }
    if (i >= row_end) continue;
    const float4 rb = sbox[i];
    const float ra = sarea[i];
    const bool exact = A.test.force_exact || !(W.flags[e] & 1);
    const int row_tile0 = i & ~63;
    const int il = i & 63;
    u64* mcol = W.mask + (size_t)e * W.NW * W.NP + i;  // word-major: mask[e][w][row], coalesced over rows
    for (int tt = 0; tt < kMaskCols / 64; ++tt) {
      const int ct0 = col0 + 64 * tt;
      if (ct0 >= col_end) break;
      if (ct0 < row_tile0) continue;  // warp-uniform: 32 consecutive rows share a 64-row tile
      const int nc = min(64, col_end - ct0);
      const bool diag = (ct0 == row_tile0);
      u64 bits = 0ull;
      if (!exact) {
        bits = row_word_fast(rb, ra, cbox + 64 * tt, carea + 64 * tt, A.test);
      } else {
        for (int j = 0; j < nc; ++j) {
          const float4 cb = cbox[64 * tt + j];
          const float ca = carea[64 * tt + j];
          // below the diagonal the column box is the earlier one
          const bool hit = (diag && j < il) ? iou_hit_exact(cb, ca, rb, ra, A.test)
                                            : iou_hit_exact(rb, ra, cb, ca, A.test);
          if (hit) bits |= (1ull << j);
        }
      }
      if (nc < 64) bits &= (1ull << nc) - 1ull;
      if (diag) {
        const u64 below = (il == 0) ? 0ull : (bits & ((1ull << il) - 1ull));
        const u64 above = (il == 63) ? 0ull : (bits & ~((2ull << il) - 1ull));
        W.diagcol[(size_t)e * W.NP + i] = below;
        bits = above;
      }
      mcol[(size_t)(ct0 >> 6) * W.NP] = bits;
      if (bits) {
        // sparse view of the same information for the edge sweep: (suppressed box, suppressing box)
        const int cntb = __popcll(bits);
        const int base = atomicAdd(&W.ecount[e], cntb);
        if (base + cntb <= W.ecap) {
          u64* ed = W.edges + (size_t)e * W.ecap + base;
          u64 m = bits;
          int k = 0;
          while (m) {
            const int b = __ffsll((long long)m) - 1;
            m &= m - 1ull;
            ed[k++] = ((u64)(uint32_t)(ct0 + b) << 32) | (uint32_t)i;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 4. emit (device function, called by the sweep CTA of an episode the moment the episode finishes): turns the
//    kept bits (visiting order, in shared memory) into the caller's output layout.
// ------------------------------------------------------------------------------------------------
// exclusive prefix of popcounts over `nw` 64-bit words (smem in, smem out)
__device__ void prefix_popc(const u64* words, int* prefix, int nw, int* warp_tot) {
  int base = 0;
  for (int w0 = 0; w0 < nw; w0 += blockDim.x) {
    int w = w0 + threadIdx.x;
    int v = (w < nw) ? __popcll(words[w]) : 0;
    int total;
    int ex = block_exclusive_scan(v, warp_tot, total);
    if (w < nw) prefix[w] = base + ex;
    base += total;
  }
  __syncthreads();
}

// vbits: kept bits in visiting order (smem, NW words, zero beyond the swept blocks); scratch: NW u64 + NW int
__device__ void emit_episode(const CandLayout& L, const NmsWorkspace& W, const NmsOutputs& O, int post_top_n, int e,
                             int total, const u64* vbits, u64* cbits, int* prefix, int* warp_tot) {
  const int tid = threadIdx.x;
  const int NW = W.NW;
  const int n = W.n[e];
  const size_t so = (size_t)e * W.NP;
  for (int w = tid; w < NW; w += blockDim.x) cbits[w] = 0ull;
  __syncthreads();
  const bool cut = post_top_n > 0 && total > post_top_n;
  const int64_t seg0 = L.seg ? L.seg[e] : 0;
  if (cut) {
    // best post_top_n by score = the first post_top_n kept boxes in visiting order
    prefix_popc(vbits, prefix, NW, warp_tot);
    for (int i = tid; i < n; i += blockDim.x) {
      u64 wv = vbits[i >> 6];
      if ((wv >> (i & 63)) & 1ull) {
        int p = prefix[i >> 6] + __popcll(wv & ((1ull << (i & 63)) - 1ull));
        if (p < post_top_n) {
          if (O.keep_out) O.keep_out[seg0 + p] = seg0 + W.sidx[so + i];
          if (O.out_boxes) {
            size_t oo = (size_t)e * O.K + p;
            reinterpret_cast<float4*>(O.out_boxes)[oo] = W.sbox[so + i];
            O.out_scores[oo] = W.sscore[so + i];
            O.out_index[oo] = W.sidx[so + i];
          }
        }
      }
    }
  } else {
    // every kept box, ascending candidate index (nms_cpu.cpp:64 nonzero(suppressed == 0))
    for (int i = tid; i < n; i += blockDim.x) {
      if ((vbits[i >> 6] >> (i & 63)) & 1ull) {
        int c = W.sidx[so + i];
        atomicOr(&cbits[c >> 6], 1ull << (c & 63));
      }
    }
    __syncthreads();
    prefix_popc(cbits, prefix, NW, warp_tot);
    for (int i = tid; i < n; i += blockDim.x) {
      if ((vbits[i >> 6] >> (i & 63)) & 1ull) {
        int c = W.sidx[so + i];
        int p = prefix[c >> 6] + __popcll(cbits[c >> 6] & ((1ull << (c & 63)) - 1ull));
        if (O.keep_out) O.keep_out[seg0 + p] = seg0 + c;
        if (O.out_boxes && p < O.K) {
          size_t oo = (size_t)e * O.K + p;
          reinterpret_cast<float4*>(O.out_boxes)[oo] = W.sbox[so + i];
          O.out_scores[oo] = W.sscore[so + i];
          O.out_index[oo] = c;
        }
      }
    }
  }
  if (tid == 0) {
    int cnt = cut ? post_top_n : total;
    if (O.keep_counts) O.keep_counts[e] = cnt;
    if (O.out_count) O.out_count[e] = O.out_boxes ? min(cnt, O.K) : cnt;
    if (O.kept_total) O.kept_total[e] = total;
  }
}

// ------------------------------------------------------------------------------------------------
// 3. sweep: one CTA per episode walks the 64-box blocks in visiting order.  The suppression word of block b is
//    *pulled*: OR over all kept earlier boxes i of mask[b][i] -- a coalesced column read (the mask is word-major)
//    and a tree reduction, no atomics.  The column of block b+1 is loaded before block b is resolved (it does not
//    depend on the outcome; the kept bits are applied as a predicate afterwards), so the serial chain per block is
//    reduce -> barrier -> warp-0 ballot fixpoint on the transposed diagonal tile -> barrier.
// ------------------------------------------------------------------------------------------------
struct SweepArgs {
  int blk_begin;  // first 64-box block of this pass
  int blk_end;    // blocks [blk_begin, blk_end) are swept (clipped to the episode)
  int stop;       // finish the episode once this many boxes are kept
  int passthrough;
  int post_top_n; // emit: keep the best post_top_n when more survive (<= 0: all)
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ u64 warp_or_u64(u64 v) {
  const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)v);
  const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(v >> 32));
  return ((u64)hi << 32) | lo;
}

__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(CandLayout L, NmsWorkspace W, NmsOutputs O, SweepArgs A) {
  extern __shared__ u64 kw[];  // [NW] kept bits of the blocks swept so far | scratch (ring / emit)
  __shared__ u64 partial[kSweepThreads / 32];
  __shared__ int s_nk;
  __shared__ int warp_tot[33];
  const int e = blockIdx.x;
  if (W.done[e]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = W.n[e];
  const int nblk = (n + 63) >> 6;
  u64* kb = W.keptbits + (size_t)e * W.NW;
  if (A.passthrough) {
    // boxlist_nms with nms_thresh <= 0: nothing is suppressed
    for (int w = tid; w < W.NW; w += kSweepThreads) {
      u64 v = 0;
      if (w < nblk) {
        int nv = min(64, n - 64 * w);
        v = nv == 64 ? ~0ull : ((1ull << nv) - 1ull);
      }
      kb[w] = v;
      kw[w] = v;
    }
    __syncthreads();
    emit_episode(L, W, O, A.post_top_n, e, n, kw, kw + W.NW, reinterpret_cast<int*>(kw + 2 * W.NW), warp_tot);
    if (tid == 0) {
      W.kcount[e] = n;
      W.done[e] = 1;
      atomicAdd(&W.sched[0], 1);
    }
    return;
  }
  const u64* __restrict__ maskT = W.mask + (size_t)e * W.NW * W.NP;  // [word][row]
  const u64* __restrict__ dc = W.diagcol + (size_t)e * W.NP;
  const int blim = min(A.blk_end, nblk);
  const int nb = blim - A.blk_begin;  // blocks swept by this pass (may be <= 0)
  // A pass that ends within 4096 boxes ("small": the early-exit passes) keeps kSweepDepth columns in flight: every
  // thread cp.asyncs the words of its own rows (tid + v*512) into a private slice of a shared-memory ring and reads
  // them back itself, so no barrier is needed for the ring, and the L2 latency is off the serial chain.  The
  // transposed diagonal tiles of the whole pass are preloaded as well.
  const bool small = 64 * blim <= kSweepThreads * kSweepPre;
  u64* ring = kw + W.NW;                                           // [kSweepDepth][kSweepThreads * kSweepPre]
  u64* dcs = ring + kSweepDepth * kSweepThreads * kSweepPre;       // [kSweepThreads * kSweepPre]
  constexpr int kRingCol = kSweepThreads * kSweepPre;
  auto prefetch_col = [&](int col, int slot) {
    const u64* src = maskT + (size_t)col * W.NP;
    u64* dst = ring + slot * kRingCol;
    for (int row = tid; row < 64 * col; row += kSweepThreads) cp_async8(dst + row, src + row);
  };
  for (int w = tid; w < W.NW; w += kSweepThreads) kw[w] = (w < A.blk_begin) ? kb[w] : 0ull;
  int count = W.kcount[e];
  __syncthreads();
  int it = 0;
  // ---- sparse sweep.  When few pairs overlap enough to suppress (the usual case at nms_thresh 0.6-0.8) the mask
  //      kernel's edge list (i, j) = "j suppresses i if j is kept" is tiny.  kept[] is the unique fixpoint of
  //      kept[i] = !exists (i, j): kept[j]; starting from all-kept, parallel relaxation over the edges reaches it in
  //      (longest suppression chain + 1) rounds -- no per-block serial walk at all.  Dense cases (edge list over
  //      capacity or long chains) fall through to the block sweep below, which is exact for any input.
  const int ne = W.ecount[e];
  bool sparse_done = false;
  if (nb > 0 && ne <= W.ecap) {
    u64* sup = ring;             // [NW]
    u64* eds = ring + W.NW;      // [ne] staged edges (ne <= kEdgeCap fits the ring)
    const u64* ge = W.edges + (size_t)e * W.ecap;
    for (int k = tid; k < ne; k += kSweepThreads) eds[k] = ge[k];
    for (int w = A.blk_begin + tid; w < blim; w += kSweepThreads) {
      const int nv = min(64, n - 64 * w);
      kw[w] = nv == 64 ? ~0ull : ((1ull << nv) - 1ull);
    }
    __syncthreads();
    const uint32_t lo_box = 64u * (uint32_t)A.blk_begin, hi_box = 64u * (uint32_t)blim;
    int rounds = 0;
    int changed = 1;
    while (changed && rounds < kEdgeRounds) {
      for (int w = A.blk_begin + tid; w < blim; w += kSweepThreads) sup[w] = 0ull;
      __syncthreads();
      for (int k = tid; k < ne; k += kSweepThreads) {
        const u64 ed = eds[k];
        const uint32_t bi = (uint32_t)(ed >> 32), bj = (uint32_t)ed;
        if (bi >= lo_box && bi < hi_box && ((kw[bj >> 6] >> (bj & 63)) & 1ull)) atomicOr(&sup[bi >> 6], 1ull << (bi & 63));
      }
      __syncthreads();
      int ch = 0;
      for (int w = A.blk_begin + tid; w < blim; w += kSweepThreads) {
        const int nv = min(64, n - 64 * w);
        const u64 vm = nv == 64 ? ~0ull : ((1ull << nv) - 1ull);
        const u64 nk = vm & ~sup[w];
        if (nk != kw[w]) {
          ch = 1;
          kw[w] = nk;
        }
      }
      changed = __syncthreads_or(ch);
      ++rounds;
    }
    if (!changed) {
      int c = 0;
      for (int w = A.blk_begin + tid; w < blim; w += kSweepThreads) {
        c += __popcll(kw[w]);
        kb[w] = kw[w];
      }
      count += block_sum(c, warp_tot);
      it = nb;
      sparse_done = true;
    } else {
      // did not converge within the round budget: redo this pass with the block sweep
      for (int w = A.blk_begin + tid; w < blim; w += kSweepThreads) kw[w] = 0ull;
      __syncthreads();
    }
  }
  if (!sparse_done && small && nb > 0) {
#pragma unroll
    for (int d = 0; d < kSweepDepth - 1; ++d) {
      if (d < nb) prefetch_col(A.blk_begin + d, d);
      cp_async_commit();
    }
    for (int i = tid; i < 64 * nb; i += kSweepThreads) dcs[i] = dc[64 * A.blk_begin + i];
    __syncthreads();
  }
  for (; it < nb && !sparse_done; ++it) {
    const int blk = A.blk_begin + it;
    u64 c_lo = 0ull, c_hi = 0ull;
    u64 acc = 0ull;
    if (small) {
      if (it + kSweepDepth - 1 < nb) prefetch_col(blk + kSweepDepth - 1, (it + kSweepDepth - 1) % kSweepDepth);
      cp_async_commit();
      cp_async_wait<kSweepDepth - 1>();  // this thread's words of column `blk` have landed
      const u64* colw = ring + (it % kSweepDepth) * kRingCol;
      for (int row = tid; row < 64 * blk; row += kSweepThreads)
        if ((kw[row >> 6] >> (row & 63)) & 1ull) acc |= colw[row];
      if (warp == 0) {
        c_lo = dcs[64 * it + lane];
        c_hi = dcs[64 * it + 32 + lane];
      }
    } else {
      if (warp == 0) {
        c_lo = dc[64 * blk + lane];
        c_hi = dc[64 * blk + 32 + lane];
      }
      const u64* col = maskT + (size_t)blk * W.NP;
#pragma unroll 4
      for (int row = tid; row < 64 * blk; row += kSweepThreads)
        if ((kw[row >> 6] >> (row & 63)) & 1ull) acc |= col[row];
    }
    // suppression word of this block: OR over the kept earlier boxes
    acc = warp_or_u64(acc);
    if (lane == 0) partial[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      const u64 rem = warp_or_u64(lane < kSweepThreads / 32 ? partial[lane] : 0ull);
      const int nv = min(64, n - 64 * blk);
      const u64 vmask = nv == 64 ? ~0ull : ((1ull << nv) - 1ull);
      const u64 free_ = ~rem & vmask;
      u64 kept = free_;
      uint32_t lo, hi;
      while (true) {
        const bool a = ((free_ >> lane) & 1ull) && ((c_lo & kept) == 0ull);
        const bool b = ((free_ >> (lane + 32)) & 1ull) && ((c_hi & kept) == 0ull);
        lo = __ballot_sync(0xffffffffu, a);
        hi = __ballot_sync(0xffffffffu, b);
        const u64 nk = ((u64)hi << 32) | lo;
        if (nk == kept) break;
        kept = nk;
      }
      if (lane == 0) {
        s_nk = __popc(lo) + __popc(hi);
        kw[blk] = kept;
        kb[blk] = kept;
      }
    }
    __syncthreads();
    count += s_nk;
    if (count >= A.stop) {
      ++it;
      break;
    }
  }
  cp_async_wait<0>();
  const int blk_next = A.blk_begin + (nb > 0 ? it : 0);
  const bool finished = (count >= A.stop) || (blk_next >= nblk);
  if (finished) {
    for (int w = blk_next + tid; w < W.NW; w += kSweepThreads) kb[w] = 0ull;
    __syncthreads();  // kw[] complete; the ring is free to be reused as emit scratch
    emit_episode(L, W, O, A.post_top_n, e, count, kw, kw + W.NW, reinterpret_cast<int*>(kw + 2 * W.NW), warp_tot);
  }
  if (!finished && !A.passthrough) {
    // the kept rows so far as a compact ascending list for the next pass's mask tiles (kw[] holds blocks [0, blk_next))
    __syncthreads();
    int* prefix = reinterpret_cast<int*>(kw + W.NW);   // the ring is idle now
    prefix_popc(kw, prefix, blk_next, warp_tot);
    int32_t* kl = W.klist + (size_t)e * W.NP;
    for (int w = tid; w < blk_next; w += kSweepThreads) {
      u64 m = kw[w];
      int p = prefix[w];
      while (m) {
        const int b = __ffsll((long long)m) - 1;
        m &= m - 1ull;
        kl[p++] = 64 * w + b;
      }
    }
  }
  if (tid == 0) {
    W.kcount[e] = count;
    if (finished) {
      W.done[e] = 1;
      atomicAdd(&W.sched[0], 1);
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t nms_workspace_carve(Carver& c, int64_t E, int64_t max_len, NmsWorkspace* ws) {
  const int64_t NP = (int64_t)align_up((size_t)(max_len > 0 ? max_len : 1), 64);
  const int64_t NW = NP / 64;
  NmsWorkspace w{};
  w.E = (int32_t)E;
  w.NP = (int32_t)NP;
  w.NW = (int32_t)NW;
  w.sbox = c.take<float4>(E * NP);
  w.sarea = c.take<float>(E * NP);
  w.sscore = c.take<float>(E * NP);
  w.sidx = c.take<int32_t>(E * NP);
  w.n = c.take<int32_t>(E);
  w.flags = c.take<int32_t>(E);
  w.done = c.take<int32_t>(E);
  w.kcount = c.take<int32_t>(E);
  w.sched = c.take<int32_t>(8);
  w.diagcol = c.take<unsigned long long>(E * NP);
  w.keptbits = c.take<unsigned long long>(E * NW);
  w.sortkeys = c.take<unsigned long long>(E * NP);
  w.edges = c.take<unsigned long long>(E * kEdgeCap);
  w.ecount = c.take<int32_t>(E);
  w.ecap = kEdgeCap;
  w.mask = c.take<unsigned long long>(E * NP * NW);
  w.klist = c.take<int32_t>(E * NP);
  if (ws) *ws = w;
  return c.total();
}

int nms_run(const CandLayout& L, const NmsWorkspace& W, const NmsParams& P, const NmsOutputs& O,
            cudaStream_t stream) {
  {
    const void* ks[4] = {reinterpret_cast<const void*>(nms_chunk_sort_kernel), reinterpret_cast<const void*>(nms_merge_kernel),
                         reinterpret_cast<const void*>(nms_mask_kernel), reinterpret_cast<const void*>(nms_sweep_kernel)};
    for (const void* k : ks) {
      int rc2 = ensure_max_shared_carveout(k);
      if (rc2 != OSD_OK) return rc2;
    }
  }
  const int E = W.E;
  if (E <= 0) return OSD_OK;
  const int max_len = P.max_len;
  OSD_REQUIRE(max_len <= W.NP, "nms: max_len %d exceeds the planned capacity %d", max_len, W.NP);
  OSD_REQUIRE(E <= 65535, "nms: at most 65535 segments per call (got %d)", E);

  // ---- sort: chunk sort + merge by rank
  {
    dim3 g((unsigned)ceil_div(max_len > 0 ? max_len : 1, kChunk), (unsigned)E);
    nms_chunk_sort_kernel<<<g, kChunkThreads, 0, stream>>>(L, W);
    OSD_LAUNCH_CHECK("nms_chunk_sort_kernel");
    timeline_mark("nms_chunk_sort_kernel", stream);
    const size_t merge_smem = (size_t)std::min<int64_t>(kMergeStage, (int64_t)align_up((size_t)(max_len > 0 ? max_len : 1), kChunk)) * sizeof(u64);
    {
      int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(nms_merge_kernel), kMergeStage * sizeof(u64));
      if (rc2 != OSD_OK) return rc2;
    }
    nms_merge_kernel<<<g, kMergeThreads, merge_smem, stream>>>(L, W);
    OSD_LAUNCH_CHECK("nms_merge_kernel");
    timeline_mark("nms_merge_kernel", stream);
  }

  // ---- mask + sweep, in passes over growing prefixes of the visiting order.  Without early exit there is one
  //      pass; with it, the first pass covers just enough boxes to keep post_top_n + 1 if little is suppressed,
  //      the second a 30 % larger prefix, the last everything.  Finished episodes skip later passes on the device.
  const size_t sweep_smem = std::max(((size_t)W.NW + (size_t)(kSweepDepth + 1) * kSweepThreads * kSweepPre) * sizeof(u64),
                                     (size_t)W.NW * (2 * sizeof(u64) + sizeof(int)) + 64);
  {
    OSD_REQUIRE(sweep_smem <= 220 * 1024, "nms: %d candidates per episode exceed the sweep capacity", max_len);
    int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(nms_sweep_kernel), sweep_smem);
    if (rc2 != OSD_OK) return rc2;
  }
  SweepArgs S{};
  S.passthrough = P.passthrough;
  S.post_top_n = P.post_top_n;
  if (P.passthrough) {
    S.blk_begin = 0;
    S.blk_end = INT_MAX;
    S.stop = INT_MAX;
    nms_sweep_kernel<<<E, kSweepThreads, sweep_smem, stream>>>(L, W, O, S);
    OSD_LAUNCH_CHECK("nms_sweep_kernel");
    timeline_mark("nms_sweep_kernel", stream);
  } else {
    MaskArgs M{};
    M.test.thr = P.thr;
    M.test.strict = P.strict;
    M.test.force_exact = !(P.thr > 0.0f && std::isfinite(P.thr));
    M.test.thr_lo = (float)((double)P.thr * (1.0 - 1.0 / 524288.0));
    M.test.thr_hi = (float)((double)P.thr * (1.0 + 1.0 / 524288.0));
    M.test.c_lo = (float)((double)P.thr / (1.0 + (double)P.thr) * (1.0 - 1.0 / 524288.0));
    const int NPu = (int)align_up((size_t)(max_len > 0 ? max_len : 1), 64);
    const bool early = P.early_exit && P.post_top_n > 0;
    S.stop = early ? P.post_top_n + 1 : INT_MAX;
    int bounds[7];
    int npass = 0;
    if (early) {
      // enough boxes to keep post_top_n + 1 when at most ~5 % of the best-scored boxes are suppressed, then prefixes
      // that grow by 1.5x (2112 -> 3200 -> 4800 -> 7200 -> all for post_top_n = 2000): when the best-scored boxes
      // suppress each other heavily (a trained detector's clusters around objects) the exit is reached after a fraction
      // of the 67 M pairs of the full problem instead of after all of them, and the rows of the earlier passes enter the
      // later ones through the compact list of their kept boxes only.  Measured on the clustered bench workload
      // (post-processing chain alone): growth 2.0 0.63 ms, 1.5 0.45 ms, 1.3 0.55 ms.  A pass whose episodes have all
      // finished returns at once (W.sched[0]).
      const int b1 = (int)align_up((size_t)P.post_top_n + 1, 64) + 64;
      if (b1 < NPu) {
        bounds[npass++] = b1;
        // short lists grow by 4x; no pass within 25 % of the full list
        // OSD_NMS_MAX_PASSES (diagnosis): cap on the number of passes, the last one always covers everything
        static const int max_passes = [] { const char* e = getenv("OSD_NMS_MAX_PASSES"); return e ? atoi(e) : 7; }();
        // OSD_NMS_GROWTH (percent, diagnosis): growth of the prefix from pass to pass for lists of 2048+ rows
        static const int growth = [] { const char* e = getenv("OSD_NMS_GROWTH"); const int g = e ? atoi(e) : 150; return g < 110 ? 110 : g; }();
        auto next = [&](int b) {
          if (b < 2048) return 4 * (b - 64);
          if (growth == 200) return 2 * (b - 64);
          return (int)align_up((size_t)((int64_t)b * growth / 100), 128);
        };
        for (int b = next(b1); npass < 6 && npass + 1 < max_passes && (int64_t)4 * b <= (int64_t)3 * NPu; b = next(b))
          bounds[npass++] = b;
      }
    }
    bounds[npass++] = NPu;
    int prev = 0;
    for (int p = 0; p < npass; ++p) {
      const int hi = bounds[p];
      // boxes [0, hi) against the new columns [prev, hi)
      M.row_begin = 0;
      M.row_end = hi;
      M.col_begin = prev;
      M.col_end = hi;
      M.tiles_r = (int)ceil_div(hi, kMaskRows);
      M.tiles_c = (int)ceil_div(hi - prev, kMaskCols);
      M.pass = p;
      // OSD_NMS_COMPACT=0 (diagnosis): every row of the earlier passes again, kept or not
      static const bool compact_on = [] { const char* e = getenv("OSD_NMS_COMPACT"); return !(e && e[0] == '0'); }();
      M.compact_rows = (p > 0 && compact_on) ? (prev / kMaskRows) * kMaskRows : 0;
      const int64_t tiles = (int64_t)E * M.tiles_r * M.tiles_c;
      OSD_REQUIRE(tiles < (1ll << 31), "nms: too many mask tiles");
      const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * 12);
      nms_mask_kernel<<<grid, kMaskRows, 0, stream>>>(W, M);
      OSD_LAUNCH_CHECK("nms_mask_kernel");
    timeline_mark("nms_mask_kernel", stream);
      S.blk_begin = prev / 64;
      S.blk_end = hi / 64;
      nms_sweep_kernel<<<E, kSweepThreads, sweep_smem, stream>>>(L, W, O, S);
      OSD_LAUNCH_CHECK("nms_sweep_kernel");
    timeline_mark("nms_sweep_kernel", stream);
      prev = hi;
    }
  }

  return OSD_OK;
}

}  // namespace osd

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int osd_batched_nms_plan(int64_t num_segments, int64_t max_seg_len, osd_nms_plan* plan) {
  OSD_REQUIRE(plan != nullptr, "osd_batched_nms_plan: plan is null");
  OSD_REQUIRE(num_segments >= 0 && max_seg_len >= 0, "osd_batched_nms_plan: negative size");
  OSD_REQUIRE(max_seg_len < (1ll << 24), "osd_batched_nms_plan: segment too long (%lld)", (long long)max_seg_len);
  osd::Carver c(nullptr);
  osd::NmsWorkspace w{};
  osd::nms_workspace_carve(c, num_segments > 0 ? num_segments : 1, max_seg_len, &w);
  plan->workspace_bytes = c.total();
  plan->padded_len = w.NP;
  plan->mask_words = w.NW;
  return OSD_OK;
}

extern "C" int osd_batched_nms(const float* boxes, const float* scores, const int64_t* seg_offsets,
                               int64_t num_segments, int64_t max_seg_len, float threshold, int strict,
                               void* workspace, size_t workspace_bytes, int64_t* keep_out,
                               int32_t* keep_counts, void* stream) {
  if (num_segments == 0) return OSD_OK;
  OSD_REQUIRE(num_segments > 0 && max_seg_len >= 0, "osd_batched_nms: negative size");
  OSD_REQUIRE(seg_offsets && keep_counts, "osd_batched_nms: null seg_offsets / keep_counts");
  OSD_REQUIRE(max_seg_len == 0 || (boxes && scores && keep_out), "osd_batched_nms: null boxes / scores / keep_out");
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(boxes) & 15) == 0, "osd_batched_nms: boxes must be 16-byte aligned");
  OSD_REQUIRE(max_seg_len < (1ll << 24), "osd_batched_nms: segment too long");
  osd::Carver c(workspace);
  osd::NmsWorkspace W{};
  osd::nms_workspace_carve(c, num_segments, max_seg_len, &W);
  if (c.total() > workspace_bytes || (!workspace && c.total() > 0)) {
    osd::set_error("osd_batched_nms: workspace of %zu bytes needed, %zu given", c.total(), workspace_bytes);
    return OSD_ERR_WORKSPACE;
  }
  osd::CandLayout L{};
  L.boxes = reinterpret_cast<const float4*>(boxes);
  L.scores = scores;
  L.seg = seg_offsets;
  osd::NmsParams P{};
  P.thr = threshold;
  P.strict = strict ? 1 : 0;
  P.post_top_n = 0;
  P.early_exit = 0;
  P.max_len = (int)max_seg_len;
  P.passthrough = 0;
  osd::NmsOutputs O{};
  O.keep_out = keep_out;
  O.keep_counts = keep_counts;
  return osd::nms_run(L, W, P, O, static_cast<cudaStream_t>(stream));
}
