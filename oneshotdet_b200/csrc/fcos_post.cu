// FCOS post-processing for sm_100a: one fused pass per (episode, level) that scores every location,
// radix-selects the pre-NMS top-k in shared memory and writes decoded, clipped candidates in location
// order -- then hands the candidates to the batched NMS pipeline (nms.cu).  Replaces the ~25 ATen ops
// per level, the per-image Python loop and its .item() syncs of
// maskrcnn_benchmark/modeling/rpn/fcos/inference.py:46-137 and :251-323.
#include <climits>

#include "osd_common.cuh"
#include <cooperative_groups.h>

#include "osd_device_utils.cuh"

namespace cg = cooperative_groups;

namespace osd {
// Debug aid (-DOSD_DEBUG_TS): %globaltimer stamps at the phase boundaries of one P3 CTA, read back with
// osd_debug_select_stamps() -- how the phase times next to the matching stream in DESIGN section 4 were measured.
#ifdef OSD_DEBUG_TS
__device__ unsigned long long g_sel_ts[8];
#define OSD_STAMP(k)                                                                       \
  do {                                                                                     \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == (gridDim.z - 1) && threadIdx.x == 0) { \
      unsigned long long _t;                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                               \
      g_sel_ts[k] = _t;                                                                    \
    }                                                                                      \
  } while (0)
#else
#define OSD_STAMP(k) do {} while (0)
#endif

namespace {

constexpr int kCl = 8;              // CTAs per cluster = per (episode, level)
constexpr int kSelThreads = 256;
constexpr int kSelLoads = 9;          // loads in flight per thread and array while scoring (9 x 256 >= a 2100-location slice)
constexpr int kSelMinBlocks = 5;      // launch bound: <= 48 registers, 5 CTAs per SM next to the matching kernel's CTA
constexpr int kRadixBins = 256;     // radix-select digits: 8 + 8 + 8 + 7 bits (== kSelThreads: thread t owns bin t)
constexpr int kMaxRounds = 3;       // 63 * 256 locations per round and CTA
constexpr int kMaxSlice = kMaxRounds * 63 * kSelThreads;   // 48 384 locations per CTA (193 KB of keys)

struct SelectArgs {
  int nl, B;
  int H[OSD_MAX_LEVELS], W[OSD_MAX_LEVELS], stride[OSD_MAX_LEVELS];
  const float* cls[OSD_MAX_LEVELS];
  const float* reg[OSD_MAX_LEVELS];
  const float* ctr[OSD_MAX_LEVELS];
  int slot[OSD_MAX_LEVELS];
  FastDiv div_w[OSD_MAX_LEVELS];   // location index -> (row, col)
  const int32_t* image_hw;  // [B,2] (h, w)
  float pre_thr;
  int top_n;
  float min_size;
  int cap;
  float4* cand_boxes;     // [B, cap]
  float* cand_scores;     // [B, cap]
  int32_t* cand_loc;      // [B, cap]
  int32_t* level_count;   // [B, nl]
  int reg_exp;            // OSD_REG_RAW_EXP_SCALE: reg holds the raw bbox_pred output
  float reg_scale[OSD_MAX_LEVELS];
};

__device__ __forceinline__ float sigmoidf_precise(float x) {
  // inference.py:58 / :72  torch.sigmoid in fp32: 1 / (1 + exp(-x))
  return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

// fcos.py:95-97 with layers/scale.py:10-11: torch.exp(x * scale) -- the product is rounded to fp32 before the exp
__device__ __forceinline__ float reg_distance(float v, bool raw, float scale) {
  return raw ? expf(__fmul_rn(v, scale)) : v;
}

struct Decoded {
  float4 box;
  bool ok;
};

// inference.py:104-109 decode, bounding_box.py:214-224 clip, boxlist_ops.py:202-216 size filter
__device__ __forceinline__ Decoded decode_clip(float dl, float dt, float dr, float db, int i, const FastDiv& div_w, int Wl,
                                               int stride, float xmax, float ymax, float min_size, bool raw,
                                               float scale) {
  dl = reg_distance(dl, raw, scale);
  dt = reg_distance(dt, raw, scale);
  dr = reg_distance(dr, raw, scale);
  db = reg_distance(db, raw, scale);
  const int row = (int)fdiv((uint32_t)i, div_w), col = i - row * Wl;
  const float px = (float)(col * stride + stride / 2);  // fcos.py:220-234
  const float py = (float)(row * stride + stride / 2);
  Decoded d;
  d.box.x = fminf(fmaxf(__fsub_rn(px, dl), 0.f), xmax);
  d.box.y = fminf(fmaxf(__fsub_rn(py, dt), 0.f), ymax);
  d.box.z = fminf(fmaxf(__fadd_rn(px, dr), 0.f), xmax);
  d.box.w = fminf(fmaxf(__fadd_rn(py, db), 0.f), ymax);
  const float ws = __fadd_rn(__fsub_rn(d.box.z, d.box.x), 1.0f);
  const float hs = __fadd_rn(__fsub_rn(d.box.w, d.box.y), 1.0f);
  d.ok = (ws >= min_size) && (hs >= min_size);
  return d;
}

// One thread-block CLUSTER of kCl CTAs per (episode, level): CTA r scores and keeps the keys of its slice of the
// level in its own shared memory; the radix-select histograms are combined through distributed shared memory
// (every CTA sums the kCl histograms and runs the same bin search), and the ordered compaction uses the slice
// totals exchanged the same way.  The P3 level (16 800 locations) is thus worked on by 8 SMs instead of one.
__global__ void __cluster_dims__(kCl, 1, 1) __launch_bounds__(kSelThreads, kSelMinBlocks) fcos_select_kernel(SelectArgs A) {
  OSD_STAMP(0);
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ uint32_t sm_dyn[];
  __shared__ int warp_tot[33];
  __shared__ int xch[4];   // read by the other CTAs of the cluster: [0] candidates, [1] equal keys, [2] survivors
  __shared__ int s_bin, s_kk, s_eq;
  int* hist0 = reinterpret_cast<int*>(sm_dyn);                  // [2][kRadixBins] this CTA's histograms (ping-pong)
  int* red = reinterpret_cast<int*>(sm_dyn) + 2 * kRadixBins;   // [kRadixBins] cluster-wide histogram
  uint32_t* keys = sm_dyn + 3 * kRadixBins;                     // [slice]

  const int r = (int)cluster.block_rank();
  const int l = blockIdx.y, e = blockIdx.z, tid = threadIdx.x;
  const int Wl = A.W[l], HW = A.H[l] * Wl, stride = A.stride[l];
  const bool raw = A.reg_exp != 0;
  const float rscale = A.reg_scale[l];
  const int slice = (HW + kCl - 1) / kCl;
  const int lo = min(HW, r * slice), hi = min(HW, lo + slice);
  const int nloc = hi - lo;
  const float* __restrict__ cls = A.cls[l] + (size_t)e * HW;
  const float* __restrict__ ctr = A.ctr[l] + (size_t)e * HW;
  const float* __restrict__ reg = A.reg[l] + (size_t)e * 4 * HW;

  // ---- 1. scores -> order-preserving keys (0 = not a candidate).  kSelLoads independent loads per array are in flight
  //         per thread: a P3 slice of the 800x1344 geometry (2100 locations) is fetched in ONE round trip, which is what
  //         matters when the HBM queues are full of the matching stream's traffic
  int my_cnt = 0;
  for (int i0 = tid; i0 < nloc; i0 += kSelLoads * kSelThreads) {
    float xc[kSelLoads], xt[kSelLoads];
#pragma unroll
    for (int u = 0; u < kSelLoads; ++u) {
      const int i = i0 + u * kSelThreads;
      xc[u] = (i < nloc) ? __ldg(cls + lo + i) : 0.f;
      xt[u] = (i < nloc) ? __ldg(ctr + lo + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kSelLoads; ++u) {
      const int i = i0 + u * kSelThreads;
      if (i < nloc) {
        const float p = sigmoidf_precise(xc[u]);
        const float c = sigmoidf_precise(xt[u]);
        const float sc = __fmul_rn(p, c);                   // inference.py:79
        const bool cand = p > A.pre_thr;                    // inference.py:74 (tested before the multiply)
        keys[i] = cand ? (__float_as_uint(sc) + 1u) : 0u;   // sc >= 0, so its bit pattern is monotone
        my_cnt += cand ? 1 : 0;
      }
    }
  }
  const int cnt_r = block_sum(my_cnt, warp_tot);        // (contains the barrier that publishes keys)
  if (tid == 0) xch[0] = cnt_r;
  cluster.sync();
  int cnt = 0;
#pragma unroll
  for (int rr = 0; rr < kCl; ++rr) cnt += cluster.map_shared_rank(xch, rr)[0];
  const int k = min(cnt, A.top_n);                      // inference.py:75-76

  OSD_STAMP(1);
  // ---- 2. k-th largest key of the whole level by 3-digit radix select (only when the level overflows top_n)
  const bool take_all = (cnt <= k);
  uint32_t T = 1u;   // threshold key
  int need_eq = 0;   // how many keys equal to T are taken (lowest locations first)
  int eq_total = 0;  // how many keys equal T
  if (!take_all) {
    // keys are (bits of a float in [0, 1]) + 1 < 2^31: digits = exponent (8 bits) and the mantissa in 8/8/7 bits.
    // Each pass: local 256-bin histogram (warp-aggregated shared-memory atomics), cluster barrier, then EVERY CTA
    // sums the kCl histograms through distributed shared memory -- thread t owns bin t, 8 remote reads -- and runs the
    // same suffix-scan bin search, so no result has to be broadcast.
    uint32_t prefix = 0u, pmask = 0u;
    int kk = k;
    const int shifts[4] = {23, 15, 7, 0};
    const int widths[4] = {8, 8, 8, 7};
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int sh = shifts[pass];
      const uint32_t dm = (1u << widths[pass]) - 1u;
      // ping-pong histograms: the buffer zeroed here was last read by the neighbours two passes ago, and this CTA is
      // past the cluster barrier of the previous pass, which they only reach after those reads -- so one cluster
      // barrier per pass is enough (cluster barriers are what slows down next to the matching stream)
      int* hist = hist0 + (pass & 1) * kRadixBins;
      hist[tid] = 0;   // kRadixBins == kSelThreads
      __syncthreads();
      for (int i0 = 0; i0 < nloc; i0 += kSelThreads) {
        const int i = i0 + tid;
        const uint32_t key = (i < nloc) ? keys[i] : 0u;
        const bool on = key != 0u && (key & pmask) == prefix;
        const uint32_t bin = on ? ((key >> sh) & dm) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (on && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[bin], __popc(peers));
      }
      cluster.sync();  // every CTA's histogram is complete
      int v = 0;
#pragma unroll
      for (int rr = 0; rr < kCl; ++rr) v += cluster.map_shared_rank(hist, rr)[tid];
      // suffix sums: thread tid looks at bin q = 255 - tid, so the exclusive prefix over threads = keys in bins > q
      // (the remote read above used bin tid; exchange through shared memory)
      red[tid] = v;
      __syncthreads();
      const int q = kSelThreads - 1 - tid;
      const int vq = red[q];
      int total;
      const int above = block_exclusive_scan(vq, warp_tot, total);
      if (above < kk && kk <= above + vq) {
        s_bin = q;
        s_kk = kk - above;
        s_eq = vq;
      }
      __syncthreads();
      prefix |= ((uint32_t)s_bin) << sh;
      pmask |= dm << sh;
      kk = s_kk;
      eq_total = s_eq;   // after the last pass: number of keys equal to the threshold key
    }
    T = prefix;
    need_eq = kk;
  }

  OSD_STAMP(2);
  // ---- 3. decode + clip + size filter + ordered compaction.  Each thread owns a contiguous run of `per`
  //         locations of the slice (per is odd: conflict-free shared-memory reads), so location order is
  //         (CTA rank, thread) order.
  const int img_h = A.image_hw[2 * e], img_w = A.image_hw[2 * e + 1];
  const float xmax = (float)(img_w - 1), ymax = (float)(img_h - 1);
  const size_t obase = (size_t)e * A.cap + A.slot[l];
  const bool ties = !take_all && need_eq < eq_total;  // ties at the top-k boundary: the lowest locations win
  int eq_before = 0;
  if (ties) {
    // equal keys in the slices of lower-ranked CTAs come first
    int my_eq = 0;
    for (int i = tid; i < nloc; i += kSelThreads) my_eq += (keys[i] == T) ? 1 : 0;
    const int eq_r = block_sum(my_eq, warp_tot);
    if (tid == 0) xch[1] = eq_r;
    cluster.sync();
    for (int rr = 0; rr < r; ++rr) eq_before += cluster.map_shared_rank(xch, rr)[1];
  }
  // per <= 63 so a run's flags fit one 64-bit mask; a slice of up to kMaxRounds * 63 * 256 locations takes
  // kMaxRounds rounds (one for the BASELINE geometry)
  const int per = min(((nloc + kSelThreads - 1) / kSelThreads) | 1, 63);
  const int round = per * kSelThreads;
  unsigned long long okm[kMaxRounds];
  int pos0[kMaxRounds];
  int running = 0, eq_seen = eq_before;
  // A: which locations are selected and survive the size filter; positions inside this CTA's slice
#pragma unroll
  for (int rd = 0; rd < kMaxRounds; ++rd) {
    okm[rd] = 0ull;
    pos0[rd] = 0;
    const int seg0 = rd * round;
    if (seg0 < nloc) {  // uniform over the CTA
      const int first = seg0 + tid * per;
      const int run_end = min(min(nloc, seg0 + round), first + per);  // this thread owns [first, run_end) of the slice
      int eq_rank = 0;
      if (ties) {
        int my_eq = 0;
        for (int i = first; i < run_end; ++i) my_eq += (keys[i] == T) ? 1 : 0;
        int tot;
        eq_rank = eq_seen + block_exclusive_scan(my_eq, warp_tot, tot);
        eq_seen += tot;
      }
      unsigned long long m = 0ull;
      for (int qq = 0; first + qq < run_end; ++qq) {
        const uint32_t key = keys[first + qq];
        bool sel;
        if (take_all) sel = key != 0u;
        else if (key > T) sel = true;
        else if (key == T) sel = !ties || (eq_rank++ < need_eq);
        else sel = false;
        if (sel) m |= 1ull << qq;
      }
      // survivors of the size filter (4 locations' loads in flight at a time)
      unsigned long long ok = 0ull;
      while (m) {
        int qs[4];
        float v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          qs[u] = m ? (__ffsll((long long)m) - 1) : -1;
          if (m) m &= m - 1ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (qs[u] >= 0) {
            const int i = lo + first + qs[u];
            v[u][0] = reg[i];
            v[u][1] = reg[HW + i];
            v[u][2] = reg[2 * HW + i];
            v[u][3] = reg[3 * HW + i];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (qs[u] >= 0) {
            const Decoded d = decode_clip(v[u][0], v[u][1], v[u][2], v[u][3], lo + first + qs[u], A.div_w[l], Wl, stride,
                                          xmax, ymax, A.min_size, raw, rscale);
            if (d.ok) ok |= 1ull << qs[u];
          }
        }
      }
      int tot;
      pos0[rd] = running + block_exclusive_scan(__popcll(ok), warp_tot, tot);
      running += tot;
      okm[rd] = ok;
    }
  }
  OSD_STAMP(3);
  // survivors in the slices of lower-ranked CTAs come first
  if (tid == 0) xch[2] = running;
  cluster.sync();
  int base = 0;
  for (int rr = 0; rr < r; ++rr) base += cluster.map_shared_rank(xch, rr)[2];
  if (r == kCl - 1 && tid == 0) A.level_count[e * A.nl + l] = base + running;
  // B: write (the loads now hit L1)
#pragma unroll
  for (int rd = 0; rd < kMaxRounds; ++rd) {
    unsigned long long m = okm[rd];
    const int first = rd * round + tid * per;
    int pos = base + pos0[rd];
    while (m) {
      const int qq = __ffsll((long long)m) - 1;
      m &= m - 1ull;
      const int i = lo + first + qq;
      const Decoded d = decode_clip(reg[i], reg[HW + i], reg[2 * HW + i], reg[3 * HW + i], i, A.div_w[l], Wl, stride, xmax,
                                    ymax, A.min_size, raw, rscale);
      A.cand_boxes[obase + pos] = d.box;
      A.cand_scores[obase + pos] = __uint_as_float(keys[first + qq] - 1u);
      A.cand_loc[obase + pos] = i;
      ++pos;
    }
  }
  cluster.sync();  // no CTA exits while its shared memory may still be read by a neighbour
  OSD_STAMP(4);
}

struct FcosBuffers {
  float4* cand_boxes;
  float* cand_scores;
  int32_t* cand_loc;
  int32_t* level_count;
  int32_t* kept_total;
  NmsWorkspace nms;
};

int validate(const osd_fcos_config* cfg) {
  OSD_REQUIRE(cfg != nullptr, "fcos: config is null");
  OSD_REQUIRE(cfg->num_levels >= 1 && cfg->num_levels <= OSD_MAX_LEVELS, "fcos: num_levels %d out of range",
              cfg->num_levels);
  OSD_REQUIRE(cfg->batch >= 0 && cfg->batch <= 65535, "fcos: batch %d out of range", cfg->batch);
  OSD_REQUIRE(cfg->pre_nms_top_n >= 1, "fcos: pre_nms_top_n must be >= 1");
  OSD_REQUIRE(cfg->reg_transform == OSD_REG_DISTANCES || cfg->reg_transform == OSD_REG_RAW_EXP_SCALE,
              "fcos: unknown reg_transform %d", cfg->reg_transform);
  int64_t cap = 0;
  for (int l = 0; l < cfg->num_levels; ++l) {
    OSD_REQUIRE(cfg->height[l] >= 1 && cfg->width[l] >= 1 && cfg->stride[l] >= 1, "fcos: bad level %d geometry", l);
    const int64_t hw = (int64_t)cfg->height[l] * cfg->width[l];
    OSD_REQUIRE(hw <= (int64_t)kCl * kMaxSlice, "fcos: level %d has %lld locations; at most %d are supported", l,
                (long long)hw, kCl * kMaxSlice);
    OSD_REQUIRE(((int64_t)cfg->width[l] + 1) * cfg->stride[l] < (1 << 24) &&
                ((int64_t)cfg->height[l] + 1) * cfg->stride[l] < (1 << 24), "fcos: level %d coordinates overflow fp32 integers", l);
    cap += hw < cfg->pre_nms_top_n ? hw : cfg->pre_nms_top_n;
  }
  OSD_REQUIRE(cap < (1 << 24), "fcos: %lld candidates per episode is out of range", (long long)cap);
  return OSD_OK;
}

void carve(const osd_fcos_config* cfg, Carver& c, FcosBuffers* buf, osd_fcos_plan* plan) {
  const int B = cfg->batch > 0 ? cfg->batch : 1;
  int cap = 0;
  int slot[OSD_MAX_LEVELS] = {0};
  FcosBuffers b{};
  for (int l = 0; l < cfg->num_levels; ++l) {
    const int hw = cfg->height[l] * cfg->width[l];
    slot[l] = cap;
    cap += hw < cfg->pre_nms_top_n ? hw : cfg->pre_nms_top_n;
  }
  const size_t o_boxes = c.offset_of_next();
  b.cand_boxes = c.take<float4>((size_t)B * cap);
  const size_t o_scores = c.offset_of_next();
  b.cand_scores = c.take<float>((size_t)B * cap);
  const size_t o_loc = c.offset_of_next();
  b.cand_loc = c.take<int32_t>((size_t)B * cap);
  const size_t o_lc = c.offset_of_next();
  b.level_count = c.take<int32_t>((size_t)B * cfg->num_levels);
  const size_t o_kt = c.offset_of_next();
  b.kept_total = c.take<int32_t>(B);
  nms_workspace_carve(c, B, cap, &b.nms);
  if (buf) *buf = b;
  if (plan) {
    plan->workspace_bytes = c.total();
    plan->cand_capacity = cap;
    plan->out_capacity = (cfg->post_nms_top_n > 0 && cfg->post_nms_top_n < cap) ? cfg->post_nms_top_n : cap;
    for (int l = 0; l < OSD_MAX_LEVELS; ++l) plan->level_slot[l] = l < cfg->num_levels ? slot[l] : 0;
    plan->off_cand_boxes = o_boxes;
    plan->off_cand_scores = o_scores;
    plan->off_cand_loc = o_loc;
    plan->off_level_count = o_lc;
    plan->off_kept_count = o_kt;
  }
}

}  // namespace
}  // namespace osd

#ifdef OSD_DEBUG_TS
extern "C" int osd_debug_select_stamps(unsigned long long* out8) {
  return (int)cudaMemcpyFromSymbol(out8, osd::g_sel_ts, sizeof(unsigned long long) * 8);
}
#endif

extern "C" int osd_fcos_postprocess_plan(const osd_fcos_config* cfg, osd_fcos_plan* plan) {
  OSD_REQUIRE(plan != nullptr, "osd_fcos_postprocess_plan: plan is null");
  int rc = osd::validate(cfg);
  if (rc != OSD_OK) return rc;
  osd::Carver c(nullptr);
  osd::carve(cfg, c, nullptr, plan);
  return OSD_OK;
}

extern "C" int osd_fcos_postprocess(const osd_fcos_config* cfg, const float* const* cls, const float* const* reg,
                                    const float* const* ctr, const int32_t* image_hw, void* workspace,
                                    size_t workspace_bytes, float* out_boxes, float* out_scores,
                                    int32_t* out_index, int32_t* out_count, void* stream_) {
  using namespace osd;
  int rc = validate(cfg);
  if (rc != OSD_OK) return rc;
  if (cfg->batch == 0) return OSD_OK;
  OSD_REQUIRE(cls && reg && ctr && image_hw, "osd_fcos_postprocess: null input array");
  OSD_REQUIRE(out_boxes && out_scores && out_index && out_count, "osd_fcos_postprocess: null output");
  OSD_REQUIRE((reinterpret_cast<uintptr_t>(out_boxes) & 15) == 0, "osd_fcos_postprocess: out_boxes must be 16-byte aligned");
  OSD_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "osd_fcos_postprocess: workspace must be 256-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver c(workspace);
  FcosBuffers buf{};
  osd_fcos_plan plan{};
  carve(cfg, c, &buf, &plan);
  if (plan.workspace_bytes > workspace_bytes) {
    set_error("osd_fcos_postprocess: workspace of %zu bytes needed, %zu given", plan.workspace_bytes, workspace_bytes);
    return OSD_ERR_WORKSPACE;
  }

  SelectArgs A{};
  A.nl = cfg->num_levels;
  A.B = cfg->batch;
  int max_slice = 0;
  for (int l = 0; l < cfg->num_levels; ++l) {
    OSD_REQUIRE(cls[l] && reg[l] && ctr[l], "osd_fcos_postprocess: null level %d input", l);
    A.H[l] = cfg->height[l];
    A.W[l] = cfg->width[l];
    A.stride[l] = cfg->stride[l];
    A.cls[l] = cls[l];
    A.reg[l] = reg[l];
    A.ctr[l] = ctr[l];
    A.slot[l] = plan.level_slot[l];
    A.div_w[l] = make_fastdiv((uint32_t)cfg->width[l]);
    const int slice = (cfg->height[l] * cfg->width[l] + kCl - 1) / kCl;
    if (slice > max_slice) max_slice = slice;
  }
  A.image_hw = image_hw;
  A.pre_thr = cfg->pre_nms_thresh;
  A.top_n = cfg->pre_nms_top_n;
  A.min_size = cfg->min_size;
  A.reg_exp = cfg->reg_transform == OSD_REG_RAW_EXP_SCALE;
  for (int l = 0; l < cfg->num_levels; ++l) A.reg_scale[l] = cfg->reg_scale[l];
  A.cap = plan.cand_capacity;
  A.cand_boxes = buf.cand_boxes;
  A.cand_scores = buf.cand_scores;
  A.cand_loc = buf.cand_loc;
  A.level_count = buf.level_count;

  const size_t smem = (size_t)3 * kRadixBins * sizeof(int) + (size_t)max_slice * sizeof(uint32_t);
  {
    const size_t cap_bytes = (size_t)3 * kRadixBins * sizeof(int) + (size_t)kMaxSlice * sizeof(uint32_t);
    int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(fcos_select_kernel), cap_bytes);
    if (rc2 != OSD_OK) return rc2;
    rc2 = ensure_max_shared_carveout(reinterpret_cast<const void*>(fcos_select_kernel));
    if (rc2 != OSD_OK) return rc2;
  }
  // cluster of kCl CTAs along x per (level, episode); __cluster_dims__ on the kernel makes <<<>>> launch clusters
  dim3 grid((unsigned)kCl, (unsigned)cfg->num_levels, (unsigned)cfg->batch);
  timeline_mark("post_begin", stream);
  fcos_select_kernel<<<grid, kSelThreads, smem, stream>>>(A);
  OSD_LAUNCH_CHECK("fcos_select_kernel");
    timeline_mark("fcos_select_kernel", stream);

  CandLayout L{};
  L.boxes = buf.cand_boxes;
  L.scores = buf.cand_scores;
  L.seg = nullptr;
  L.level_count = buf.level_count;
  L.nl = cfg->num_levels;
  L.cap = plan.cand_capacity;
  for (int l = 0; l < cfg->num_levels; ++l) L.slot[l] = plan.level_slot[l];
  NmsParams P{};
  P.thr = cfg->nms_thresh;
  P.strict = cfg->strict ? 1 : 0;
  P.post_top_n = cfg->post_nms_top_n;
  P.early_exit = cfg->early_exit ? 1 : 0;
  P.max_len = plan.cand_capacity;
  P.passthrough = !(cfg->nms_thresh > 0.0f);  // boxlist_ops.py:22-23
  NmsOutputs O{};
  O.out_boxes = out_boxes;
  O.out_scores = out_scores;
  O.out_index = out_index;
  O.out_count = out_count;
  O.K = plan.out_capacity;
  O.kept_total = buf.kept_total;
  return nms_run(L, buf.nms, P, O, stream);
}
