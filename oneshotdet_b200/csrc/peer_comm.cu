// Detection exchange over peer memory: replaces the reference's pickle-based all_gather
// (maskrcnn_benchmark/utils/comm.py:48-88) for the fixed-shape result block.  See include/osd_b200.h.
#include <algorithm>
#include <cstring>

#include "osd_common.cuh"

namespace osd {
namespace {

constexpr int kMaxPeers = 16;
struct PushArgs {
  void* dst[kMaxPeers];
  int n;
  const uint4* src;
  size_t vecs;   // 16-byte vectors
};

// grid (ctas_per_dst, num_dst): each CTA streams its slice of the block to one destination; the loads hit L2 after the
// first destination, the stores go out through NVLink (or stay local for the rank's own buffer)
__global__ void __launch_bounds__(256) peer_push_kernel(PushArgs A) {
  uint4* d = static_cast<uint4*>(A.dst[blockIdx.y]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.vecs; i += (size_t)gridDim.x * blockDim.x)
    d[i] = __ldg(A.src + i);
}

}  // namespace
}  // namespace osd

extern "C" int osd_comm_alloc(size_t bytes, void** ptr) {
  using namespace osd;
  OSD_REQUIRE(ptr != nullptr && bytes > 0, "osd_comm_alloc: bad arguments");
  OSD_CUDA(cudaMalloc(ptr, bytes));
  OSD_CUDA(cudaMemset(*ptr, 0, bytes));
  return OSD_OK;
}

extern "C" int osd_comm_free(void* ptr) {
  using namespace osd;
  if (ptr) OSD_CUDA(cudaFree(ptr));
  return OSD_OK;
}

extern "C" int osd_comm_export(void* ptr, unsigned char handle[OSD_IPC_HANDLE_BYTES]) {
  using namespace osd;
  static_assert(sizeof(cudaIpcMemHandle_t) == OSD_IPC_HANDLE_BYTES, "IPC handle size");
  OSD_REQUIRE(ptr != nullptr && handle != nullptr, "osd_comm_export: null pointer");
  cudaIpcMemHandle_t h;
  OSD_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return OSD_OK;
}

extern "C" int osd_comm_import(const unsigned char handle[OSD_IPC_HANDLE_BYTES], void** ptr) {
  using namespace osd;
  OSD_REQUIRE(ptr != nullptr && handle != nullptr, "osd_comm_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  OSD_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return OSD_OK;
}

extern "C" int osd_comm_close(void* ptr) {
  using namespace osd;
  if (ptr) OSD_CUDA(cudaIpcCloseMemHandle(ptr));
  return OSD_OK;
}

extern "C" int osd_comm_push(void* const* dst, int32_t num_dst, const void* src, size_t bytes, void* stream_) {
  using namespace osd;
  OSD_REQUIRE(dst != nullptr && src != nullptr && num_dst >= 0, "osd_comm_push: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int i = 0; i < num_dst; ++i) {
    OSD_REQUIRE(dst[i] != nullptr, "osd_comm_push: destination %d is null", i);
    OSD_CUDA(cudaMemcpyAsync(dst[i], src, bytes, cudaMemcpyDefault, stream));   // peer copy: DMA engines, no SM
  }
  return OSD_OK;
}

extern "C" int osd_comm_push_kernel(void* const* dst, int32_t num_dst, const void* src, size_t bytes, void* stream_) {
  using namespace osd;
  OSD_REQUIRE(dst != nullptr && src != nullptr, "osd_comm_push_kernel: bad arguments");
  OSD_REQUIRE(num_dst >= 0 && num_dst <= kMaxPeers, "osd_comm_push_kernel: at most %d destinations", kMaxPeers);
  OSD_REQUIRE((bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "osd_comm_push_kernel: 16-byte granularity");
  if (num_dst == 0 || bytes == 0) return OSD_OK;
  PushArgs A{};
  A.n = num_dst;
  A.src = static_cast<const uint4*>(src);
  A.vecs = bytes / 16;
  for (int i = 0; i < num_dst; ++i) {
    OSD_REQUIRE(dst[i] != nullptr && (reinterpret_cast<uintptr_t>(dst[i]) & 15) == 0, "osd_comm_push_kernel: destination %d", i);
    A.dst[i] = dst[i];
  }
  const int per = (int)std::min<size_t>(4, (A.vecs + 255) / 256);
  peer_push_kernel<<<dim3((unsigned)per, (unsigned)num_dst), 256, 0, static_cast<cudaStream_t>(stream_)>>>(A);
  OSD_LAUNCH_CHECK("peer_push_kernel");
  return OSD_OK;
}
