// Support -> target matching on the FPN levels for sm_100a: one launch streams all five levels.
//   product:  out[b,c,h,w] = feat[b,c,h,w] * mean_s supp[b*S+s, c]      (generalized_rcnn.py:100-104, :306-311)
//   concat :  out[b, 0:C] = feat[b], out[b, C:2C] = broadcast(mean_s supp)   (box_head.py:147; reversed :144)
// HBM-bound elementwise stream: 128-bit non-allocating loads / streaming stores, 4 vectors in flight per
// thread and up to 2048 threads per SM, persistent grid of 148 x k CTAs walking 16 KB output chunks, no
// barriers; pooled support scalars come from L1.
// The K-shot mean is the sequential fp32 sum divided by S (what ATen's mean computes on this shape), the
// multiply a single rounded fp32 product: fp32 results are bit-identical to the reference expression.
#include <cuda_bf16.h>

#include <atomic>
#include <cstdlib>
#include <cstring>

#include "osd_common.cuh"
#include "osd_device_utils.cuh"

namespace osd {
namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;                     // independent 128-bit loads in flight per thread
constexpr int kChunkVec = kThreads * kVecPerThread;  // 1024 x 16 B = 16 KB of output per chunk

struct Level {
  const void* feat;
  const void* supp;
  void* out;
  uint32_t hw;
  uint32_t out_elems;    // B * Cout * HW
  uint32_t chunk_begin;  // first global chunk index of this level
  FastDiv div_hw;        // NCHW: plane = g / HW
  FastDiv div_img;       // NHWC: b = g / (HW * Cout)
};

struct Args {
  int nl, B, S, C, Cout, mode;
  uint32_t total_chunks;
  int l2_evict_first;    // bulk kernel: tag the streamed lines evict-first in L2
  int sched_slot;        // bulk kernel: which pair of dynamic-scheduler counters this launch uses
  FastDiv div_cout;      // q / Cout (NCHW concat: plane -> episode; NHWC: offset -> pixel)
  Level lv[OSD_MAX_LEVELS];
};

template <typename T> struct Vec;
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int N = 8; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f(float& d, float v) { d = v; }
__device__ __forceinline__ void from_f(__nv_bfloat16& d, float v) { d = __float2bfloat16_rn(v); }

__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// mean over the S shots of channel c of episode b: sequential fp32 sum, one IEEE division
template <typename T>
__device__ __forceinline__ float pooled_at(const T* __restrict__ supp, int S, int C, uint32_t q /* b*C + c */,
                                           const FastDiv& div_c) {
  if (S == 1) return to_f(supp[q]);
  const uint32_t b = fdiv(q, div_c), c = q - b * C;
  const T* p = supp + (size_t)b * S * C + c;
  float acc = to_f(p[0]);
  for (int s = 1; s < S; ++s) acc = __fadd_rn(acc, to_f(p[(size_t)s * C]));
  const float m = __fdiv_rn(acc, (float)S);
  T rounded;  // the mean is a tensor of the input dtype in the reference (bf16 inputs: rounded to bf16)
  from_f(rounded, m);
  return to_f(rounded);
}

// ---------------------------------------------------------------------------------------------
// NCHW.  Output element g of a level lives in plane po = g / HW at offset r.
//   product: value = feat[g] * pooled(po)
//   concat : b = po / 2C, co = po % 2C; feature half copies feat[(b*C + cc)*HW + r], support half
//            writes pooled(b*C + cc)
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE>
__device__ __forceinline__ float nchw_value(const Args& A, const Level& L, uint32_t po, uint32_t r,
                                            const FastDiv& div_c) {
  const T* feat = static_cast<const T*>(L.feat);
  const T* supp = static_cast<const T*>(L.supp);
  if (MODE == OSD_MATCH_PRODUCT) {
    return __fmul_rn(to_f(feat[(size_t)po * L.hw + r]), pooled_at(supp, A.S, A.C, po, div_c));
  }
  const uint32_t b = fdiv(po, A.div_cout), co = po - b * A.Cout;
  const bool first_half = co < (uint32_t)A.C;
  const uint32_t cc = first_half ? co : co - A.C;
  const bool is_feat = (MODE == OSD_MATCH_CONCAT) ? first_half : !first_half;
  if (is_feat) return to_f(feat[((size_t)b * A.C + cc) * L.hw + r]);
  return pooled_at(supp, A.S, A.C, b * A.C + cc, div_c);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 5) match_nchw_kernel(Args A, FastDiv div_c) {
  constexpr int N = Vec<T>::N;
  // no shared memory, no barriers: every warp streams independently; the pooled support scalar of a plane is a
  // 4-byte read that stays in L1 (B*C*S values per level)
  for (uint32_t chunk = blockIdx.x; chunk < A.total_chunks; chunk += gridDim.x) {
    int li = 0;
#pragma unroll
    for (int k = 1; k < OSD_MAX_LEVELS; ++k)
      if (k < A.nl && chunk >= A.lv[k].chunk_begin) li = k;
    const Level& L = A.lv[li];
    const uint32_t nvec = (L.out_elems + N - 1) / N;
    const uint32_t v0 = (chunk - L.chunk_begin) * kChunkVec;
    const uint32_t v1 = min(v0 + kChunkVec, nvec);
    T* out = static_cast<T*>(L.out);
    uint4 in[kVecPerThread];
    uint32_t po[kVecPerThread], rr[kVecPerThread];
    bool whole[kVecPerThread], is_feat[kVecPerThread];
    // phase 1: index math + all loads in flight
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      const uint32_t v = v0 + j * kThreads + threadIdx.x;
      whole[j] = false;
      is_feat[j] = true;
      if (v < v1) {
        const uint32_t g = v * N;
        po[j] = fdiv(g, L.div_hw);
        rr[j] = g - po[j] * L.hw;
        whole[j] = (rr[j] + N <= L.hw) && (g + N <= L.out_elems);
        if (whole[j]) {
          size_t src = g;
          if (MODE != OSD_MATCH_PRODUCT) {
            const uint32_t b = fdiv(po[j], A.div_cout), co = po[j] - b * A.Cout;
            const bool first_half = co < (uint32_t)A.C;
            const uint32_t cc = first_half ? co : co - A.C;
            is_feat[j] = (MODE == OSD_MATCH_CONCAT) ? first_half : !first_half;
            src = ((size_t)b * A.C + cc) * L.hw + rr[j];
            po[j] = b * A.C + cc;  // pooled index for the support half
          }
          if (is_feat[j]) in[j] = ld_stream(static_cast<const T*>(L.feat) + src);
        }
      }
    }
    // phase 2: compute + streaming stores
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      const uint32_t v = v0 + j * kThreads + threadIdx.x;
      if (v >= v1) continue;
      const uint32_t g = v * N;
      if (whole[j]) {
        uint4 o;
        T* ov = reinterpret_cast<T*>(&o);
        const T* iv = reinterpret_cast<const T*>(&in[j]);
        if (MODE == OSD_MATCH_PRODUCT) {
          const float s = pooled_at(static_cast<const T*>(L.supp), A.S, A.C, po[j], div_c);
#pragma unroll
          for (int k = 0; k < N; ++k) from_f(ov[k], __fmul_rn(to_f(iv[k]), s));
        } else if (is_feat[j]) {
          o = in[j];
        } else {
          const float s = pooled_at(static_cast<const T*>(L.supp), A.S, A.C, po[j], div_c);
#pragma unroll
          for (int k = 0; k < N; ++k) from_f(ov[k], s);
        }
        st_stream(out + g, o);
      } else {
        // vector straddles a plane boundary (HW not a multiple of the vector width) or the tensor end
        uint32_t p = fdiv(g, L.div_hw), r = g - p * L.hw;
        for (int k = 0; k < N && g + k < L.out_elems; ++k) {
          while (r >= L.hw) {
            r -= L.hw;
            ++p;
          }
          from_f(out[g + k], nchw_value<T, MODE>(A, L, p, r, div_c));
          ++r;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NHWC (channels_last).  Output element g = ((b*HW + pix) * Cout + co); requires C % vector width == 0,
// so a vector never leaves its pixel nor its half.
// ---------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 5) match_nhwc_kernel(Args A, FastDiv div_c) {
  constexpr int N = Vec<T>::N;
  for (uint32_t chunk = blockIdx.x; chunk < A.total_chunks; chunk += gridDim.x) {
    int li = 0;
#pragma unroll
    for (int k = 1; k < OSD_MAX_LEVELS; ++k)
      if (k < A.nl && chunk >= A.lv[k].chunk_begin) li = k;
    const Level& L = A.lv[li];
    const uint32_t nvec = L.out_elems / N;
    const uint32_t v0 = (chunk - L.chunk_begin) * kChunkVec;
    const uint32_t v1 = min(v0 + kChunkVec, nvec);
    const T* feat = static_cast<const T*>(L.feat);
    const T* supp = static_cast<const T*>(L.supp);
    T* out = static_cast<T*>(L.out);
    uint4 in[kVecPerThread];
    uint32_t bq[kVecPerThread];  // b*C + cc
    bool is_feat[kVecPerThread];
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      const uint32_t v = v0 + j * kThreads + threadIdx.x;
      is_feat[j] = true;
      if (v < v1) {
        const uint32_t g = v * N;
        const uint32_t b = fdiv(g, L.div_img);
        const uint32_t off = g - b * (L.hw * A.Cout);
        const uint32_t pix = fdiv(off, A.div_cout), co = off - pix * A.Cout;
        uint32_t cc = co;
        if (MODE != OSD_MATCH_PRODUCT) {
          const bool first_half = co < (uint32_t)A.C;
          cc = first_half ? co : co - A.C;
          is_feat[j] = (MODE == OSD_MATCH_CONCAT) ? first_half : !first_half;
        }
        bq[j] = b * A.C + cc;
        if (is_feat[j]) in[j] = ld_stream(feat + ((size_t)b * L.hw + pix) * A.C + cc);
      }
    }
#pragma unroll
    for (int j = 0; j < kVecPerThread; ++j) {
      const uint32_t v = v0 + j * kThreads + threadIdx.x;
      if (v >= v1) continue;
      uint4 o;
      T* ov = reinterpret_cast<T*>(&o);
      const T* iv = reinterpret_cast<const T*>(&in[j]);
      if (MODE == OSD_MATCH_PRODUCT) {
#pragma unroll
        for (int k = 0; k < N; ++k)
          from_f(ov[k], __fmul_rn(to_f(iv[k]), pooled_at(supp, A.S, A.C, bq[j] + k, div_c)));
      } else if (is_feat[j]) {
        o = in[j];
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k) from_f(ov[k], pooled_at(supp, A.S, A.C, bq[j] + k, div_c));
      }
      st_stream(out + (size_t)v * N, o);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Bulk-copy (TMA) variant of the NCHW product: same arithmetic, different data movement.
// The LSU kernel above needs ~48 K registers per SM to keep 64 KB of loads in flight (4 CTAs x 256 threads), which
// starves the post-processing chain that runs concurrently on the second stream: its 512/1024-thread CTAs (merge,
// sweep) cannot become resident until the matching CTAs retire (measured with tools/timeline.py: the chain made no
// progress behind fcos_select until the match kernel ended).  Here one small CTA per SM streams through a shared
// memory ring instead: a producer thread issues cp.async.bulk global->shared (16 KB chunks, mbarrier complete_tx),
// two consumer groups scale their chunk in place and hand it to cp.async.bulk shared->global.  In-flight bytes live
// in shared memory, so the kernel holds 288 threads x ~40 registers per SM and leaves the rest to the other stream.
// ---------------------------------------------------------------------------------------------
// Ring depth, measured on one box as (two-stream step / kernel alone): 5 slots 0.183 / 0.161 ms, 6: 0.181 / 0.145,
// 7: 0.192 / 0.135, 8: 0.204 / 0.136 -- deeper queues make the kernel itself faster and its neighbours slower.
#ifndef OSD_BULK_SLOTS
#define OSD_BULK_SLOTS 7   // measured best for the three-stream step; -DOSD_BULK_SLOTS=n builds an A/B variant
#endif
constexpr int kBulkSlots = OSD_BULK_SLOTS;                     // x 16 KB ring
#ifndef OSD_BULK_GROUPS
#define OSD_BULK_GROUPS 2
#endif
constexpr int kBulkGroups = OSD_BULK_GROUPS;
constexpr int kBulkMinBlocks = 5;   // launch bound that caps the kernel at 40 registers/thread: the other stream's CTAs
                                    // need the register file
constexpr unsigned kBulkBackoffNs = 200;
constexpr int kBulkGroupThreads = 128;
constexpr int kBulkThreads = 32 + kBulkGroups * kBulkGroupThreads;
constexpr uint32_t kChunkBytes = kChunkVec * 16;
constexpr uint32_t kBulkSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(kBulkBackoffNs);   // waiting warps should not eat the issue slots of the other stream's CTAs on this SM
    if (++spins > kBulkSpinLimit) __trap();
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// the same with an L2 eviction policy: the feature stream is touched once, the concurrently running post-processing
// chain's working set (head outputs, candidates, sort keys, bitmask) should be what stays in L2
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
               ::"l"(dst), "r"(src), "r"(bytes), "l"(pol) : "memory");
}

struct ChunkRef {
  int level;
  uint32_t v0;      // first vector of the chunk inside the level
  uint32_t bytes;   // 16 KB, or the level's tail
};

template <typename T>
__device__ __forceinline__ ChunkRef locate_chunk(const Args& A, uint32_t chunk) {
  constexpr int N = Vec<T>::N;
  int li = 0;
#pragma unroll
  for (int k = 1; k < OSD_MAX_LEVELS; ++k)
    if (k < A.nl && chunk >= A.lv[k].chunk_begin) li = k;
  ChunkRef c;
  c.level = li;
  c.v0 = (chunk - A.lv[li].chunk_begin) * kChunkVec;
  const uint32_t nvec = A.lv[li].out_elems / N;   // the bulk path requires out_elems % N == 0
  c.bytes = min((uint32_t)kChunkVec, nvec - c.v0) * 16u;
  return c;
}

// Dynamic chunk scheduler: CTAs draw 16 KB chunks from a global counter, so an SM that shares its cycles with the other
// stream's CTAs (post-processing kernels, the NCCL all-gather) simply takes fewer chunks instead of finishing last.
// g_sched[slot] = {next chunk, CTAs finished}; the last CTA to finish resets the pair for the next launch.  Launches
// Launches that may overlap in time must not share a slot: eager launches rotate through the first kEagerSlots, a launch
// recorded into a CUDA graph gets one of the remaining slots for good (the graph replays with it); when those run out
// the launch uses static round-robin chunks (sched_slot < 0).
constexpr int kSchedSlots = 64;
constexpr int kEagerSlots = 32;
constexpr uint32_t kNoChunk = 0xffffffffu;
__device__ unsigned int g_sched[kSchedSlots][2];

template <typename T>
__global__ void __launch_bounds__(kBulkThreads, kBulkMinBlocks) match_product_bulk_kernel(Args A, FastDiv div_c) {
  constexpr int N = Vec<T>::N;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  __shared__ __align__(8) uint64_t full_bar[kBulkSlots], empty_bar[kBulkSlots];
  __shared__ uint32_t slot_chunk[kBulkSlots];   // which chunk the producer put into each ring slot (kNoChunk: stop)
  const uint32_t ring = (smem_addr(ring_raw) + 127u) & ~127u;
  uint8_t* ring_ptr = ring_raw + (ring - smem_addr(ring_raw));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < kBulkSlots; ++s) {
      mbar_init(smem_addr(&full_bar[s]), 1);
      mbar_init(smem_addr(&empty_bar[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    if (tid == 0) {
      const bool dynamic = A.sched_slot >= 0;
      unsigned int* sched = g_sched[dynamic ? A.sched_slot : 0];
      int stops = 0;   // one stop mark per consumer group (group g consumes slots of iterations i = g mod kBulkGroups)
      for (uint32_t i = 0; stops < kBulkGroups; ++i) {
        const uint32_t s = i % kBulkSlots;
        if (i >= (uint32_t)kBulkSlots) mbar_wait(smem_addr(&empty_bar[s]), ((i / kBulkSlots) - 1) & 1);
        const uint32_t chunk = stops ? kNoChunk : (dynamic ? atomicAdd(&sched[0], 1u) : blockIdx.x + i * gridDim.x);
        const uint32_t bar = smem_addr(&full_bar[s]);
        if (chunk >= A.total_chunks) {
          slot_chunk[s] = kNoChunk;
          mbar_arrive(bar);
          ++stops;
          continue;
        }
        slot_chunk[s] = chunk;
        const ChunkRef c = locate_chunk<T>(A, chunk);
        mbar_expect_tx(bar, c.bytes);
        const uint8_t* src = static_cast<const uint8_t*>(A.lv[c.level].feat) + (size_t)c.v0 * 16;
        if (A.l2_evict_first) bulk_load_hint(ring + s * kChunkBytes, src, c.bytes, bar, l2_evict_first_policy());
        else bulk_load(ring + s * kChunkBytes, src, c.bytes, bar);
      }
      // the last CTA out resets the counters (nobody draws from them any more)
      if (dynamic && atomicAdd(&sched[1], 1u) == gridDim.x - 1) {
        sched[0] = 0u;
        sched[1] = 0u;
      }
    }
    return;
  }

  const int g = (warp - 1) / (kBulkGroupThreads / 32);
  const int t = tid - 32 - g * kBulkGroupThreads;
  for (uint32_t i = g;; i += kBulkGroups) {
    const uint32_t s = i % kBulkSlots;
    mbar_wait(smem_addr(&full_bar[s]), (i / kBulkSlots) & 1);
    const uint32_t chunk = slot_chunk[s];
    if (chunk == kNoChunk) break;
    const ChunkRef c = locate_chunk<T>(A, chunk);
    const Level& L = A.lv[c.level];
    const T* supp = static_cast<const T*>(L.supp);
    uint4* slot = reinterpret_cast<uint4*>(ring_ptr + s * kChunkBytes);
    const uint32_t nv = c.bytes >> 4;
    const uint32_t e_first = c.v0 * N;
    if (L.hw >= kChunkVec * N) {
      // large planes (P3, P4: 94 % of the bytes): the chunk touches at most two planes, so the two pooled scalars are
      // fetched once per chunk and a vector only compares its element range with the plane boundary
      const uint32_t p0 = fdiv(e_first, L.div_hw);
      const uint32_t boundary = (p0 + 1) * L.hw;   // first element of the next plane (level-relative index)
      const float s0 = pooled_at(supp, A.S, A.C, p0, div_c);
      const float s1 = (boundary < e_first + nv * N) ? pooled_at(supp, A.S, A.C, p0 + 1, div_c) : s0;
#pragma unroll 8
      for (uint32_t j = t; j < nv; j += kBulkGroupThreads) {
        uint4 v = slot[j];
        const uint32_t e0 = e_first + j * N;
        T* x = reinterpret_cast<T*>(&v);
        if (e0 + N <= boundary || e0 >= boundary) {
          const float sc = (e0 >= boundary) ? s1 : s0;
#pragma unroll
          for (int k = 0; k < N; ++k) from_f(x[k], __fmul_rn(to_f(x[k]), sc));
        } else {
#pragma unroll
          for (int k = 0; k < N; ++k) from_f(x[k], __fmul_rn(to_f(x[k]), (e0 + k < boundary) ? s0 : s1));
        }
        slot[j] = v;
      }
    } else {
#pragma unroll 4
      for (uint32_t j = t; j < nv; j += kBulkGroupThreads) {
        uint4 v = slot[j];
        const uint32_t e0 = e_first + j * N;
        const uint32_t p = fdiv(e0, L.div_hw), r = e0 - p * L.hw;
        const float s0 = pooled_at(supp, A.S, A.C, p, div_c);
        T* x = reinterpret_cast<T*>(&v);
        if (r + N <= L.hw) {
#pragma unroll
          for (int k = 0; k < N; ++k) from_f(x[k], __fmul_rn(to_f(x[k]), s0));
        } else {  // the vector runs into the next plane (HW not a multiple of the vector width; HW >= N: one crossing)
          const float s1 = pooled_at(supp, A.S, A.C, p + 1, div_c);
#pragma unroll
          for (int k = 0; k < N; ++k) from_f(x[k], __fmul_rn(to_f(x[k]), (r + k < L.hw) ? s0 : s1));
        }
        slot[j] = v;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the bulk store
    asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kBulkGroupThreads) : "memory");
    if (t == 0) {
      uint8_t* dst = static_cast<uint8_t*>(L.out) + (size_t)c.v0 * 16;
      if (A.l2_evict_first) bulk_store_hint(dst, ring + s * kChunkBytes, c.bytes, l2_evict_first_policy());
      else bulk_store(dst, ring + s * kChunkBytes, c.bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (i >= (uint32_t)kBulkGroups) {   // the group's previous store has finished reading its slot: hand it back
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        mbar_arrive(smem_addr(&empty_bar[(i - kBulkGroups) % kBulkSlots]));
      }
    }
  }
  if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory stays valid until the stores are done
}

template <typename T>
bool bulk_eligible(const Args& A) {
  constexpr int N = Vec<T>::N;
  for (int l = 0; l < A.nl; ++l)
    if (A.lv[l].out_elems % N != 0 || A.lv[l].hw < (uint32_t)N) return false;
  return true;   // feat / out are 16-byte aligned (checked by osd_match_forward)
}

template <typename T, int MODE>
int launch(const Args& A, int layout, cudaStream_t stream) {
  const FastDiv div_c = make_fastdiv((uint32_t)A.C);
  // persistent grid: a multiple of the SM count, capped by the work
  // CTAs per SM of the persistent grid.  The stream is HBM-bound, a few hundred threads per SM with 4 x 128-bit loads
  // in flight each saturate it; keeping the footprint small leaves registers and thread slots for the
  // post-processing kernels that run concurrently on the second stream (OSD_MATCH_CTAS_PER_SM overrides for tuning).
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* env = getenv("OSD_MATCH_CTAS_PER_SM");
    per_sm = env ? atoi(env) : 4;
    if (per_sm < 1) per_sm = 1;
  }
  int ctas = kNumSMs * per_sm;
  if ((uint32_t)ctas > A.total_chunks) ctas = (int)A.total_chunks;
  if (ctas < 1) return OSD_OK;
  {
    int rc2 = ensure_max_shared_carveout(reinterpret_cast<const void*>(match_product_bulk_kernel<T>));
    if (rc2 != OSD_OK) return rc2;
  }
  timeline_mark("match_begin", stream);
  static int use_bulk = -1;
  if (use_bulk < 0) {
    const char* env = getenv("OSD_MATCH_BULK");
    use_bulk = env ? atoi(env) : 1;
  }
  if (MODE == OSD_MATCH_PRODUCT && layout == OSD_LAYOUT_NCHW && use_bulk && bulk_eligible<T>(A)) {
    const size_t smem = (size_t)kBulkSlots * kChunkBytes + 128;
    {
      int rc2 = ensure_dynamic_smem(reinterpret_cast<const void*>(match_product_bulk_kernel<T>), smem);
      if (rc2 != OSD_OK) return rc2;
    }
    static int bulk_ctas = 0;
    if (bulk_ctas == 0) {
      const char* env = getenv("OSD_MATCH_BULK_CTAS");
      bulk_ctas = env ? atoi(env) : kNumSMs;
      if (bulk_ctas < 1) bulk_ctas = kNumSMs;
    }
    int grid = bulk_ctas;
    if ((uint32_t)grid > A.total_chunks) grid = (int)A.total_chunks;
    static int l2_hint = -1;
    if (l2_hint < 0) {
      const char* env = getenv("OSD_MATCH_L2_EVICT_FIRST");
      l2_hint = env ? atoi(env) : 1;
    }
    Args AB = A;
    AB.l2_evict_first = l2_hint;
    {
      static std::atomic<int> next_eager{0}, next_captured{kEagerSlots};
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      OSD_CUDA(cudaStreamIsCapturing(stream, &cap));
      const char* sched_env = getenv("OSD_MATCH_SCHED");   // "static": round-robin chunks (read per launch: tests toggle it)
      if (sched_env && strcmp(sched_env, "static") == 0) {
        AB.sched_slot = -1;
      } else if (cap == cudaStreamCaptureStatusNone) {
        AB.sched_slot = next_eager.fetch_add(1) % kEagerSlots;
      } else {
        const int slot = next_captured.fetch_add(1);
        AB.sched_slot = slot < kSchedSlots ? slot : -1;   // out of dedicated slots: static chunk assignment
      }
    }
    match_product_bulk_kernel<T><<<grid, kBulkThreads, smem, stream>>>(AB, div_c);
    OSD_LAUNCH_CHECK("match_product_bulk_kernel");
    timeline_mark("match_product_bulk_kernel", stream);
    return OSD_OK;
  }
  if (layout == OSD_LAYOUT_NCHW) {
    match_nchw_kernel<T, MODE><<<ctas, kThreads, 0, stream>>>(A, div_c);
    OSD_LAUNCH_CHECK("match_nchw_kernel");
    timeline_mark("match_nchw_kernel", stream);
  } else {
    match_nhwc_kernel<T, MODE><<<ctas, kThreads, 0, stream>>>(A, div_c);
    OSD_LAUNCH_CHECK("match_nhwc_kernel");
    timeline_mark("match_nhwc_kernel", stream);
  }
  return OSD_OK;
}

template <typename T>
int dispatch_mode(const Args& A, int layout, cudaStream_t stream) {
  switch (A.mode) {
    case OSD_MATCH_PRODUCT: return launch<T, OSD_MATCH_PRODUCT>(A, layout, stream);
    case OSD_MATCH_CONCAT: return launch<T, OSD_MATCH_CONCAT>(A, layout, stream);
    case OSD_MATCH_CONCAT_REVERSED: return launch<T, OSD_MATCH_CONCAT_REVERSED>(A, layout, stream);
  }
  set_error("osd_match_forward: unknown mode %d", A.mode);
  return OSD_ERR_INVALID;
}

}  // namespace
}  // namespace osd

extern "C" int osd_match_forward(const osd_match_desc* d, void* stream) {
  using namespace osd;
  OSD_REQUIRE(d != nullptr, "osd_match_forward: desc is null");
  OSD_REQUIRE(d->num_levels >= 1 && d->num_levels <= OSD_MAX_LEVELS, "osd_match_forward: num_levels %d out of range", d->num_levels);
  OSD_REQUIRE(d->batch >= 0 && d->shots >= 1 && d->channels >= 1, "osd_match_forward: bad batch/shots/channels");
  OSD_REQUIRE(d->mode >= OSD_MATCH_PRODUCT && d->mode <= OSD_MATCH_CONCAT_REVERSED, "osd_match_forward: unknown mode %d", d->mode);
  OSD_REQUIRE(d->layout == OSD_LAYOUT_NCHW || d->layout == OSD_LAYOUT_NHWC, "osd_match_forward: unknown layout %d", d->layout);
  OSD_REQUIRE(d->dtype == OSD_DTYPE_F32 || d->dtype == OSD_DTYPE_BF16, "osd_match_forward: unknown dtype %d", d->dtype);
  if (d->batch == 0) return OSD_OK;
  const int N = d->dtype == OSD_DTYPE_F32 ? 4 : 8;
  Args A{};
  A.nl = d->num_levels;
  A.B = d->batch;
  A.S = d->shots;
  A.C = d->channels;
  A.Cout = d->mode == OSD_MATCH_PRODUCT ? d->channels : 2 * d->channels;
  A.mode = d->mode;
  A.div_cout = make_fastdiv((uint32_t)A.Cout);
  if (d->layout == OSD_LAYOUT_NHWC)
    OSD_REQUIRE(d->channels % N == 0, "osd_match_forward: NHWC needs channels %% %d == 0 (got %d)", N, d->channels);
  uint64_t chunks = 0;
  for (int l = 0; l < d->num_levels; ++l) {
    OSD_REQUIRE(d->hw[l] >= 1, "osd_match_forward: level %d is empty", l);
    OSD_REQUIRE(d->feat[l] && d->supp[l] && d->out[l], "osd_match_forward: null pointer at level %d", l);
    OSD_REQUIRE(((reinterpret_cast<uintptr_t>(d->feat[l]) | reinterpret_cast<uintptr_t>(d->out[l])) & 15) == 0,
                "osd_match_forward: level %d feat/out must be 16-byte aligned", l);
    const uint64_t elems = (uint64_t)d->batch * A.Cout * d->hw[l];
    OSD_REQUIRE(elems < (1ull << 31), "osd_match_forward: level %d has %llu output elements; split the batch", l,
                (unsigned long long)elems);
    Level& L = A.lv[l];
    L.feat = d->feat[l];
    L.supp = d->supp[l];
    L.out = d->out[l];
    L.hw = (uint32_t)d->hw[l];
    L.out_elems = (uint32_t)elems;
    L.chunk_begin = (uint32_t)chunks;
    L.div_hw = make_fastdiv(L.hw);
    L.div_img = make_fastdiv((uint32_t)((uint64_t)d->hw[l] * A.Cout));
    OSD_REQUIRE((uint64_t)d->hw[l] * A.Cout < (1ull << 32), "osd_match_forward: level %d image too large", l);
    chunks += (elems + (uint64_t)N * kChunkVec - 1) / ((uint64_t)N * kChunkVec);
  }
  OSD_REQUIRE(chunks < (1ull << 32), "osd_match_forward: too much work for one call");
  A.total_chunks = (uint32_t)chunks;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d->dtype == OSD_DTYPE_F32) return dispatch_mode<float>(A, d->layout, s);
  return dispatch_mode<__nv_bfloat16>(A, d->layout, s);
}
