// Internal helpers shared by the translation units of libosd_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "osd_b200.h"

namespace osd {

// ---- host-side error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
bool timeline_on();
void timeline_mark(const char* name, cudaStream_t stream);  // no-op unless OSD_TIMELINE=1 / osd_timeline_enable(1)

#define OSD_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::osd::set_error(__VA_ARGS__);      \
      return OSD_ERR_INVALID;             \
    }                                     \
  } while (0)

#define OSD_CUDA(call)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::osd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__,    \
                       __LINE__);                                                           \
      return OSD_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define OSD_LAUNCH_CHECK(name)                                                              \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::osd::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
      return OSD_ERR_CUDA;                                                                  \
    }                                                                                       \
    ::osd::count_launch();                                                                  \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Every kernel of the two concurrently running streams (matching || post-processing) asks for the same shared
// memory carve-out.  The L1/shared split is per-SM state: a CTA whose kernel needs a different split cannot become
// resident on an SM until the CTAs already there have drained -- which, next to a persistent kernel, means "after it
// ends" (measured: the post-processing chain did not start until the matching kernel finished).
template <typename K>
inline cudaError_t prefer_max_shared_carveout(K kernel) {
  return cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                              (int)cudaSharedmemCarveoutMaxShared);
}

// Kernel attributes are per-DEVICE state: these remember what has been configured per (device, kernel), so a process
// (or thread) that drives several GPUs configures every one of them.  `bytes` is the dynamic shared memory of the
// launch; the opt-in limit is raised whenever it exceeds what was set before (static shared memory on top of it is the
// reason there is no "only above 48 KB" shortcut).
int ensure_dynamic_smem(const void* kernel, size_t bytes);
int ensure_max_shared_carveout(const void* kernel);

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t offset_of_next() { return align_up(off, 256); }
  size_t total() const { return align_up(off, 256); }
};

// ---- NMS pipeline (nms.cu), shared by osd_batched_nms and osd_fcos_postprocess ---------------------
// Where the candidates of episode e live.
struct CandLayout {
  const float4* boxes;       // xyxy
  const float* scores;
  // mode A (seg != nullptr): compact segments, episode e = rows [seg[e], seg[e+1])
  const int64_t* seg;
  // mode B (seg == nullptr): per-level slots inside a fixed-capacity episode block
  const int32_t* level_count;  // [E, nl]
  int32_t nl;
  int32_t cap;                 // rows per episode block
  int32_t slot[OSD_MAX_LEVELS];
};

struct NmsWorkspace {
  int32_t E;
  int32_t NP;   // padded rows per episode (multiple of 64)
  int32_t NW;   // NP / 64
  float4* sbox;       // [E, NP] boxes in visiting order
  float* sarea;       // [E, NP]
  float* sscore;      // [E, NP]
  int32_t* sidx;      // [E, NP] compact candidate index of the i-th visited box
  int32_t* n;         // [E] candidates per episode
  int32_t* flags;     // [E] bit0: every box is finite with x2>=x1, y2>=y1 (fast IoU test is safe)
  int32_t* done;      // [E] episode finished (early exit or all blocks swept)
  int32_t* kcount;    // [E] boxes kept so far
  int32_t* sched;     // [8] [0]: episodes finished; [1..3]: dynamic tile counters of the mask passes
  unsigned long long* mask;     // [E, NW, NP] word-major: bit j of mask[e][w][i]: box 64w+j suppressed by box i (64w+j > i)
  unsigned long long* diagcol;  // [E, NP] for box i: which earlier boxes of its own 64-block suppress it
  unsigned long long* keptbits; // [E, NW]
  unsigned long long* sortkeys; // [E, NP] sorted 64-bit keys of the chunks
  unsigned long long* edges;    // [E, kEdgeCap] suppression edges (i << 32 | j): box j (earlier) overlaps box i enough to suppress it
  int32_t* ecount;              // [E] edges found so far (may exceed the capacity: then the bitmask sweep is used)
  int32_t ecap;
  int32_t* klist;               // [E, NP] visiting positions of the boxes kept so far (ascending), written by a sweep pass that
                                //          leaves its episode unfinished: the next pass's mask tiles walk these rows only
};

size_t nms_workspace_carve(Carver& c, int64_t E, int64_t max_len, NmsWorkspace* ws);

struct NmsParams {
  float thr;
  int strict;         // 0: >=, 1: >
  int post_top_n;     // <= 0: unlimited
  int early_exit;     // stop at post_top_n + 1 kept
  int max_len;        // host bound on candidates per episode
  int passthrough;    // 1: no suppression at all (boxlist_nms with nms_thresh <= 0, boxlist_ops.py:22-23)
};

struct NmsOutputs {
  // generic mode (keep_out != nullptr): ascending global row indices at keep_out[seg[e] + p]
  int64_t* keep_out;
  int32_t* keep_counts;
  // fcos mode
  float* out_boxes;    // [E, K, 4]
  float* out_scores;   // [E, K]
  int32_t* out_index;  // [E, K]
  int32_t* out_count;  // [E]
  int32_t K;
  int32_t* kept_total; // [E] optional: kept-before-cut (saturates at post_top_n + 1 under early exit)
};

int nms_run(const CandLayout& L, const NmsWorkspace& W, const NmsParams& P, const NmsOutputs& O,
            cudaStream_t stream);

}  // namespace osd
