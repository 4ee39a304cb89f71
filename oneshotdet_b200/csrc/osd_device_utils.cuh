// Block-level primitives shared by the kernels (blockDim.x must be a multiple of 32, <= 1024).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace osd {

// Exclusive prefix sum of `v` over the thread block; `total` receives the block sum.
// warp_tot: 33 ints of shared memory.  Contains __syncthreads(): call from uniform control flow.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tot, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int t = (lane < nwarps) ? warp_tot[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    warp_tot[lane] = ti - t;  // exclusive over warps
    if (lane == 31) warp_tot[32] = ti;
  }
  __syncthreads();
  const int res = warp_tot[warp] + inc - v;
  total = warp_tot[32];
  __syncthreads();  // warp_tot may be reused right away
  return res;
}

__device__ __forceinline__ int block_sum(int v, int* warp_tot) {
  int total;
  block_exclusive_scan(v, warp_tot, total);
  return total;
}


// exact n / d for 32-bit n, d (Lemire): q = (M * n) >> 64
struct FastDiv {
  uint64_t M;
  uint32_t d;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  f.M = d > 1 ? (0xFFFFFFFFFFFFFFFFull / d + 1ull) : 0ull;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  return f.d > 1 ? (uint32_t)__umul64hi(f.M, (uint64_t)n) : n;
}

}  // namespace osd
