// Error reporting and library-level entry points of libmdm_sm100.so.
#include <stdarg.h>

#include "common.cuh"

namespace mdm {
static thread_local char g_err[1024] = "";
long long g_launch_count = 0;
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mdm

extern "C" {
const char* mdm_last_error(void) { return mdm::g_err; }
int mdm_version(void) { return 100; }
long long mdm_launch_count(void) { return mdm::g_launch_count; }
int mdm_device_available(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n > 0 ? 1 : 0;
}
}
