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

#include "osd_common.cuh"
#include "osd_device_utils.cuh"

namespace osd {

namespace {

constexpr int kSortThreads = 1024;
constexpr int kSortMaxSmemKeys = 16384;  // 128 KB of 64-bit keys
constexpr int kMaskRows = 256;           // rows per CTA (one thread per row)
constexpr int kMaskCols = 512;           // columns per CTA (8 tiles of 64)
constexpr int kSweepThreads = 1024;

typedef unsigned long long u64;

struct IouTest {
  float thr, thr_lo, thr_hi;
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
// 1. sort: one CTA per episode, bitonic network over 64-bit keys (score key << 32 | index) in smem
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

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(CandLayout L, NmsWorkspace W) {
  extern __shared__ u64 skeys[];
  __shared__ int lvl_prefix[OSD_MAX_LEVELS + 1];
  const int e = blockIdx.x;
  const int tid = threadIdx.x;
  int n = episode_count(L, e);
  n = min(n, W.NP);
  fill_level_prefix(L, e, lvl_prefix);
  int P = 1;
  while (P < n) P <<= 1;
  for (int i = tid; i < P; i += kSortThreads) {
    u64 k = ~0ull;
    if (i < n) {
      float s = L.scores[cand_row(L, e, i, lvl_prefix)];
      k = ((u64)score_key_desc(s) << 32) | (uint32_t)i;
    }
    skeys[i] = k;
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (P >> 1); t += kSortThreads) {
        int i = 2 * t - (t & (j - 1));  // index with bit j clear
        int l = i + j;
        u64 a = skeys[i], b = skeys[l];
        bool up = (i & k) == 0;
        if ((a > b) == up) {
          skeys[i] = b;
          skeys[l] = a;
        }
      }
      __syncthreads();
    }
  }
  bool regular = true;
  for (int i = tid; i < n; i += kSortThreads)
    write_sorted(L, W, e, i, (int)(skeys[i] & 0xffffffffu), lvl_prefix, regular);
  int all_regular = __syncthreads_and(regular ? 1 : 0);
  if (tid == 0) {
    W.n[e] = n;
    W.flags[e] = all_regular ? 1 : 0;
    W.done[e] = 0;
    W.kcount[e] = 0;
  }
}

// large-N fallback (n > 16384): keys to global, rank by counting (keys are unique, so ranks are a
// permutation).  O(n^2) compares spread over the whole chip.
__global__ void __launch_bounds__(256) nms_make_keys_kernel(CandLayout L, NmsWorkspace W) {
  __shared__ int lvl_prefix[OSD_MAX_LEVELS + 1];
  const int e = blockIdx.y;
  fill_level_prefix(L, e, lvl_prefix);
  int n = min(episode_count(L, e), W.NP);
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    float s = L.scores[cand_row(L, e, i, lvl_prefix)];
    W.sortkeys[(size_t)e * W.NP + i] = ((u64)score_key_desc(s) << 32) | (uint32_t)i;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    W.n[e] = n;
    W.flags[e] = 1;
    W.done[e] = 0;
    W.kcount[e] = 0;
  }
}

__global__ void __launch_bounds__(256) nms_rank_sort_kernel(CandLayout L, NmsWorkspace W) {
  __shared__ u64 tile[256];
  __shared__ int lvl_prefix[OSD_MAX_LEVELS + 1];
  const int e = blockIdx.y;
  const int n = min(episode_count(L, e), W.NP);
  if (blockIdx.x * 256 >= n) return;
  fill_level_prefix(L, e, lvl_prefix);
  const int i = blockIdx.x * 256 + threadIdx.x;
  const u64* keys = W.sortkeys + (size_t)e * W.NP;
  u64 mine = (i < n) ? keys[i] : ~0ull;
  int rank = 0;
  for (int j0 = 0; j0 < n; j0 += 256) {
    int j = j0 + threadIdx.x;
    tile[threadIdx.x] = (j < n) ? keys[j] : ~0ull;
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 256; ++k) rank += (tile[k] < mine) ? 1 : 0;
    __syncthreads();
  }
  bool regular = true;
  if (i < n) write_sorted(L, W, e, rank, i, lvl_prefix, regular);
  if (!regular) atomicAnd(&W.flags[e], 0);
}

// ------------------------------------------------------------------------------------------------
// 2. IoU bitmask: CTA = 256 rows x 512 columns; thread = one row, 64 pairs per 64-bit word
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool iou_hit_fast(const float4& a, float aa, const float4& b, float ba,
                                             const IouTest& T) {
  float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.0f), 0.0f);
  float h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.0f), 0.0f);
  float inter = __fmul_rn(w, h);
  float u = __fsub_rn(__fadd_rn(aa, ba), inter);
  if (inter > __fmul_rn(T.thr_hi, u)) return true;
  if (inter < __fmul_rn(T.thr_lo, u)) return false;
  float q = __fdiv_rn(inter, u);
  return T.strict ? (q > T.thr) : (q >= T.thr);
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

struct MaskArgs {
  int row_begin, row_end;  // rows (visiting positions) this launch covers, multiples of 64
  int col_begin, col_end;  // columns this launch covers, multiples of 64
  IouTest test;
};

__global__ void __launch_bounds__(kMaskRows) nms_mask_kernel(NmsWorkspace W, MaskArgs A) {
  __shared__ float4 cbox[kMaskCols];
  __shared__ float carea[kMaskCols];
  const int e = blockIdx.z;
  if (W.done[e]) return;
  const int n = W.n[e];
  const int row0 = A.row_begin + blockIdx.y * kMaskRows;
  const int col0 = A.col_begin + blockIdx.x * kMaskCols;
  const int row_end = min(n, A.row_end);
  const int col_end = min(n, A.col_end);
  if (row0 >= row_end || col0 >= col_end) return;
  if (col0 + kMaskCols <= row0) return;  // every column precedes every row of this CTA
  const float4* sbox = W.sbox + (size_t)e * W.NP;
  const float* sarea = W.sarea + (size_t)e * W.NP;
  const int ncols = min(kMaskCols, col_end - col0);
  for (int c = threadIdx.x; c < ncols; c += kMaskRows) {
    cbox[c] = sbox[col0 + c];
    carea[c] = sarea[col0 + c];
  }
  __syncthreads();
  const int i = row0 + threadIdx.x;
  if (i >= row_end) return;
  const float4 rb = sbox[i];
  const float ra = sarea[i];
  const bool exact = A.test.force_exact || !(W.flags[e] & 1);
  const int row_tile0 = i & ~63;
  const int il = i & 63;
  u64* mrow = W.mask + ((size_t)e * W.NP + i) * W.NW;
  for (int t = 0; t < kMaskCols / 64; ++t) {
    const int ct0 = col0 + 64 * t;
    if (ct0 >= col_end) break;
    if (ct0 < row_tile0) continue;  // warp-uniform: 32 consecutive rows share a 64-row tile
    const int nc = min(64, col_end - ct0);
    const bool diag = (ct0 == row_tile0);
    uint32_t lo = 0, hi = 0;
    if (!exact) {
#pragma unroll 16
      for (int j = 0; j < 32; ++j) {
        if (iou_hit_fast(rb, ra, cbox[64 * t + j], carea[64 * t + j], A.test)) lo |= (1u << j);
      }
#pragma unroll 16
      for (int j = 0; j < 32; ++j) {
        if (iou_hit_fast(rb, ra, cbox[64 * t + 32 + j], carea[64 * t + 32 + j], A.test)) hi |= (1u << j);
      }
    } else {
      for (int j = 0; j < 64; ++j) {
        const float4 cb = cbox[64 * t + j];
        const float ca = carea[64 * t + j];
        // below the diagonal the column box is the earlier one
        bool hit = (diag && j < il) ? iou_hit_exact(cb, ca, rb, ra, A.test)
                                    : iou_hit_exact(rb, ra, cb, ca, A.test);
        if (hit) {
          if (j < 32) lo |= (1u << j);
          else hi |= (1u << (j - 32));
        }
      }
    }
    u64 bits = ((u64)hi << 32) | lo;
    if (nc < 64) bits &= (1ull << nc) - 1ull;  // smem beyond ncols is stale
    if (diag) {
      const u64 below = (il == 0) ? 0ull : (bits & ((1ull << il) - 1ull));
      const u64 above = (il == 63) ? 0ull : (bits & ~((2ull << il) - 1ull));
      W.diagcol[(size_t)e * W.NP + i] = below;
      bits = above;
    }
    mrow[ct0 >> 6] = bits;
  }
}

// ------------------------------------------------------------------------------------------------
// 3. sweep: one CTA per episode walks the 64-box blocks in visiting order
// ------------------------------------------------------------------------------------------------
struct SweepArgs {
  int blk_begin;  // first 64-box block of this pass
  int blk_end;    // blocks [blk_begin, blk_end) are swept (clipped to the episode)
  int word_end;   // mask words [.., word_end) are valid for the rows touched by this pass
  int stop;       // finish the episode once this many boxes are kept
  int passthrough;
};

__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(NmsWorkspace W, SweepArgs A) {
  extern __shared__ u64 rem[];  // NW words: boxes already suppressed
  __shared__ int s_rows[64];
  __shared__ int s_nk;
  const int e = blockIdx.x;
  if (W.done[e]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = W.n[e];
  const int nblk = (n + 63) >> 6;
  u64* kb = W.keptbits + (size_t)e * W.NW;
  if (A.passthrough) {
    for (int w = tid; w < W.NW; w += kSweepThreads) {
      u64 v = 0;
      if (w < nblk) {
        int nv = min(64, n - 64 * w);
        v = nv == 64 ? ~0ull : ((1ull << nv) - 1ull);
      }
      kb[w] = v;
    }
    if (tid == 0) {
      W.kcount[e] = n;
      W.done[e] = 1;
    }
    return;
  }
  const u64* mask = W.mask + (size_t)e * W.NP * W.NW;
  const u64* dc = W.diagcol + (size_t)e * W.NP;
  const int wlim = min(A.word_end, nblk);
  const int blim = min(A.blk_end, nblk);
  for (int w = tid; w < W.NW; w += kSweepThreads) rem[w] = 0ull;
  __syncthreads();
  int count = W.kcount[e];
  if (A.blk_begin > 0) {
    // rebuild the suppression state of columns >= blk_begin from the boxes kept by the earlier pass
    for (int r = warp; r < A.blk_begin * 64; r += kSweepThreads / 32) {
      if ((kb[r >> 6] >> (r & 63)) & 1ull) {
        const u64* mrow = mask + (size_t)r * W.NW;
        for (int w = A.blk_begin + lane; w < wlim; w += 32) {
          u64 v = mrow[w];
          if (v) atomicOr(&rem[w], v);
        }
      }
    }
    __syncthreads();
  }
  int blk = A.blk_begin;
  for (; blk < blim; ++blk) {
    if (warp == 0) {
      const int nv = min(64, n - 64 * blk);
      const u64 vmask = nv == 64 ? ~0ull : ((1ull << nv) - 1ull);
      const u64 free_ = ~rem[blk] & vmask;
      const u64 c_lo = dc[64 * blk + lane];
      const u64 c_hi = dc[64 * blk + 32 + lane];
      u64 kept = free_;
      uint32_t lo, hi;
      while (true) {
        bool a = ((free_ >> lane) & 1ull) && ((c_lo & kept) == 0ull);
        bool b = ((free_ >> (lane + 32)) & 1ull) && ((c_hi & kept) == 0ull);
        lo = __ballot_sync(0xffffffffu, a);
        hi = __ballot_sync(0xffffffffu, b);
        u64 nk = ((u64)hi << 32) | lo;
        if (nk == kept) break;
        kept = nk;
      }
      const uint32_t lt = (1u << lane) - 1u;
      if ((lo >> lane) & 1u) s_rows[__popc(lo & lt)] = lane;
      if ((hi >> lane) & 1u) s_rows[__popc(lo) + __popc(hi & lt)] = lane + 32;
      if (lane == 0) {
        s_nk = __popc(lo) + __popc(hi);
        kb[blk] = kept;
      }
    }
    __syncthreads();
    const int nk = s_nk;
    count += nk;
    {
      const int g = tid >> 8, wl = tid & 255;
      for (int w = blk + 1 + wl; w < wlim; w += 256) {
        u64 acc = 0ull;
#pragma unroll 4
        for (int r = g; r < nk; r += 4) acc |= mask[(size_t)(64 * blk + s_rows[r]) * W.NW + w];
        if (acc) atomicOr(&rem[w], acc);
      }
    }
    __syncthreads();
    if (count >= A.stop) {
      ++blk;
      break;
    }
  }
  const bool finished = (count >= A.stop) || (blk >= nblk);
  if (finished) {
    for (int w = blk + tid; w < W.NW; w += kSweepThreads) kb[w] = 0ull;
  }
  if (tid == 0) {
    W.kcount[e] = count;
    if (finished) W.done[e] = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// 4. emit: one CTA per episode turns kept bits into the caller's output layout
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

__global__ void __launch_bounds__(1024) nms_emit_kernel(CandLayout L, NmsWorkspace W, NmsOutputs O,
                                                        int post_top_n) {
  extern __shared__ u64 sm_words[];  // [NW] kept bits (visiting order) | [NW] kept bits (candidate order)
  __shared__ int warp_tot[33];
  const int e = blockIdx.x;
  const int tid = threadIdx.x;
  const int NW = W.NW;
  u64* vbits = sm_words;
  u64* cbits = sm_words + NW;
  int* prefix = reinterpret_cast<int*>(sm_words + 2 * NW);  // [NW]
  const int n = W.n[e];
  const int total = W.kcount[e];
  const u64* kb = W.keptbits + (size_t)e * NW;
  const size_t so = (size_t)e * W.NP;
  for (int w = tid; w < NW; w += blockDim.x) {
    vbits[w] = kb[w];
    cbits[w] = 0ull;
  }
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
  w.diagcol = c.take<unsigned long long>(E * NP);
  w.keptbits = c.take<unsigned long long>(E * NW);
  w.sortkeys = (max_len > kSortMaxSmemKeys) ? c.take<unsigned long long>(E * NP) : nullptr;
  w.mask = c.take<unsigned long long>(E * NP * NW);
  if (ws) *ws = w;
  return c.total();
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

int nms_run(const CandLayout& L, const NmsWorkspace& W, const NmsParams& P, const NmsOutputs& O,
            cudaStream_t stream) {
  const int E = W.E;
  if (E <= 0) return OSD_OK;
  const int max_len = P.max_len;
  OSD_REQUIRE(max_len <= W.NP, "nms: max_len %d exceeds the planned capacity %d", max_len, W.NP);
  OSD_REQUIRE(E <= 65535, "nms: at most 65535 segments per call (got %d)", E);

  // ---- sort
  if (max_len <= kSortMaxSmemKeys) {
    const int Pmax = next_pow2(max_len > 1 ? max_len : 1);
    const size_t smem = (size_t)Pmax * sizeof(u64);
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      OSD_CUDA(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(kSortMaxSmemKeys * sizeof(u64))));
      configured = kSortMaxSmemKeys * sizeof(u64);
    }
    nms_sort_kernel<<<E, kSortThreads, smem, stream>>>(L, W);
    OSD_LAUNCH_CHECK("nms_sort_kernel");
  } else {
    dim3 g((unsigned)ceil_div(max_len, 256), (unsigned)E);
    nms_make_keys_kernel<<<g, 256, 0, stream>>>(L, W);
    OSD_LAUNCH_CHECK("nms_make_keys_kernel");
    nms_rank_sort_kernel<<<g, 256, 0, stream>>>(L, W);
    OSD_LAUNCH_CHECK("nms_rank_sort_kernel");
  }

  // ---- mask + sweep (one pass, or a short first pass when an early exit is likely)
  const size_t sweep_smem = (size_t)W.NW * sizeof(u64);
  {
    static thread_local size_t configured = 48 * 1024;
    if (sweep_smem > configured) {
      OSD_REQUIRE(sweep_smem <= 200 * 1024, "nms: %d candidates per episode exceed the sweep capacity",
                  max_len);
      OSD_CUDA(cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sweep_smem));
      configured = sweep_smem;
    }
  }
  SweepArgs S{};
  S.passthrough = P.passthrough;
  if (P.passthrough) {
    S.blk_begin = 0;
    S.blk_end = INT_MAX;
    S.word_end = INT_MAX;
    S.stop = INT_MAX;
    nms_sweep_kernel<<<E, kSweepThreads, sweep_smem, stream>>>(W, S);
    OSD_LAUNCH_CHECK("nms_sweep_kernel");
  } else {
    MaskArgs M{};
    M.test.thr = P.thr;
    M.test.strict = P.strict;
    M.test.force_exact = !(P.thr > 0.0f && std::isfinite(P.thr));
    M.test.thr_lo = (float)((double)P.thr * (1.0 - 1.0 / 524288.0));
    M.test.thr_hi = (float)((double)P.thr * (1.0 + 1.0 / 524288.0));
    const int NPu = (int)align_up((size_t)(max_len > 0 ? max_len : 1), 64);
    const bool early = P.early_exit && P.post_top_n > 0;
    S.stop = early ? P.post_top_n + 1 : INT_MAX;
    int R1 = NPu;
    if (early) {
      int64_t guess = (int64_t)P.post_top_n + P.post_top_n / 4 + 128;
      R1 = (int)std::min<int64_t>(NPu, (int64_t)align_up((size_t)guess, 64));
    }
    // pass 1: boxes [0, R1) against each other
    M.row_begin = 0;
    M.row_end = R1;
    M.col_begin = 0;
    M.col_end = R1;
    dim3 g1((unsigned)ceil_div(R1, kMaskCols), (unsigned)ceil_div(R1, kMaskRows), (unsigned)E);
    nms_mask_kernel<<<g1, kMaskRows, 0, stream>>>(W, M);
    OSD_LAUNCH_CHECK("nms_mask_kernel");
    S.blk_begin = 0;
    S.blk_end = R1 / 64;
    S.word_end = R1 / 64;
    nms_sweep_kernel<<<E, kSweepThreads, sweep_smem, stream>>>(W, S);
    OSD_LAUNCH_CHECK("nms_sweep_kernel");
    if (R1 < NPu) {
      // pass 2 (skipped on the device for episodes that already finished): all rows against columns >= R1
      M.row_begin = 0;
      M.row_end = NPu;
      M.col_begin = R1;
      M.col_end = NPu;
      dim3 g2((unsigned)ceil_div(NPu - R1, kMaskCols), (unsigned)ceil_div(NPu, kMaskRows), (unsigned)E);
      nms_mask_kernel<<<g2, kMaskRows, 0, stream>>>(W, M);
      OSD_LAUNCH_CHECK("nms_mask_kernel");
      S.blk_begin = R1 / 64;
      S.blk_end = INT_MAX;
      S.word_end = INT_MAX;
      nms_sweep_kernel<<<E, kSweepThreads, sweep_smem, stream>>>(W, S);
      OSD_LAUNCH_CHECK("nms_sweep_kernel");
    }
  }

  // ---- emit
  const size_t emit_smem = (size_t)W.NW * (2 * sizeof(u64) + sizeof(int));
  {
    static thread_local size_t configured = 48 * 1024;
    if (emit_smem > configured) {
      OSD_REQUIRE(emit_smem <= 200 * 1024, "nms: %d candidates per episode exceed the emit capacity", max_len);
      OSD_CUDA(cudaFuncSetAttribute(nms_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)emit_smem));
      configured = emit_smem;
    }
  }
  nms_emit_kernel<<<E, 1024, emit_smem, stream>>>(L, W, O, P.post_top_n);
  OSD_LAUNCH_CHECK("nms_emit_kernel");
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
