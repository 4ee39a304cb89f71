// Error plumbing, version and device checks of libosd_b200.so.
#include <cstring>

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

}  // namespace osd

extern "C" int osd_version(void) { return 100; }

extern "C" const char* osd_last_error(void) { return osd::g_err; }

extern "C" int64_t osd_launch_count(void) { return osd::g_launches; }

extern "C" void osd_reset_launch_count(void) { osd::g_launches = 0; }

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
