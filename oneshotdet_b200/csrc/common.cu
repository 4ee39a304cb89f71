// Error plumbing, version and device checks of libosd_b200.so.
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "osd_common.cuh"

namespace osd {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches += n; }

// ---- kernel timeline (profiling aid): one CUDA event after every launch, on the launching stream ----------------
namespace {
struct Mark {
  const char* name;
  cudaEvent_t ev;
};
std::mutex g_tl_mutex;
std::vector<Mark> g_marks;
int g_timeline = -1;
}  // namespace

bool timeline_on() {
  if (g_timeline < 0) {
    const char* env = getenv("OSD_TIMELINE");
    g_timeline = (env && atoi(env) > 0) ? 1 : 0;
  }
  return g_timeline > 0;
}

void timeline_mark(const char* name, cudaStream_t stream) {
  if (!timeline_on()) return;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, stream);
  std::lock_guard<std::mutex> lock(g_tl_mutex);
  g_marks.push_back({name, ev});
}

namespace {
std::mutex g_attr_mutex;
std::map<std::pair<int, const void*>, size_t> g_smem_set;
std::map<std::pair<int, const void*>, bool> g_carveout_set;
}  // namespace

int ensure_dynamic_smem(const void* kernel, size_t bytes) {
  int dev = 0;
  OSD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  auto key = std::make_pair(dev, kernel);
  auto it = g_smem_set.find(key);
  if (it != g_smem_set.end() && it->second >= bytes) return OSD_OK;
  OSD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  g_smem_set[key] = bytes;
  return OSD_OK;
}

int ensure_max_shared_carveout(const void* kernel) {
  int dev = 0;
  OSD_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  auto key = std::make_pair(dev, kernel);
  if (g_carveout_set.count(key)) return OSD_OK;
  OSD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
  g_carveout_set[key] = true;
  return OSD_OK;
}

}  // namespace osd

extern "C" int osd_version(void) { return 100; }

extern "C" const char* osd_last_error(void) { return osd::g_err; }

extern "C" int64_t osd_launch_count(void) { return osd::g_launches; }

extern "C" void osd_reset_launch_count(void) { osd::g_launches = 0; }

extern "C" void osd_timeline_enable(int on) { osd::g_timeline = on ? 1 : 0; }

extern "C" int osd_timeline_read(char* buf, size_t cap) {
  using namespace osd;
  OSD_REQUIRE(buf != nullptr && cap > 0, "osd_timeline_read: null buffer");
  OSD_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(g_tl_mutex);
  std::string out;
  char line[160];
  for (size_t i = 0; i < g_marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_marks[0].ev, g_marks[i].ev);
    snprintf(line, sizeof(line), "%s %.6f\n", g_marks[i].name, ms);
    out += line;
  }
  for (auto& m : g_marks) cudaEventDestroy(m.ev);
  const int n = (int)g_marks.size();
  g_marks.clear();
  snprintf(buf, cap, "%s", out.c_str());
  return n;
}

extern "C" int osd_check_device(void) {
  int dev = 0;
  OSD_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  OSD_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  OSD_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    osd::set_error("libosd_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return OSD_ERR_UNSUPPORTED;
  }
  return OSD_OK;
}
