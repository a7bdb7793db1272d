// runtime.cu -- error state, version, device check.
#include "common.cuh"
#include <cstdarg>

namespace acan {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
}  // namespace acan

extern "C" int acan_abi_version(void) { return ACAN_ABI_VERSION; }
extern "C" const char* acan_last_error(void) { return acan::g_err; }

extern "C" int acan_device_check(void) {
  int dev = 0;
  ACAN_CUDA(cudaGetDevice(&dev));
  int major = 0;
  ACAN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return acan::fail(ACAN_ERR_ARCH, "device %d has compute capability %d.x; libacan_b200 is sm_100a only", dev, major);
  return 0;
}
