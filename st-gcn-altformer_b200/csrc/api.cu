// Error plumbing and device checks for the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace afb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
}  // namespace afb

extern "C" int afb_version(void) { return 100; }
extern "C" const char* afb_last_error(void) { return afb::g_err; }
extern "C" int afb_device_ok(int dev) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    afb::set_error("cudaGetDeviceProperties(%d): %s", dev, cudaGetErrorString(e));
    return (int)e;
  }
  if (prop.major != 10) {
    afb::set_error("device %d is sm_%d%d; altformer_b200 kernels are built for sm_100a only", dev, prop.major, prop.minor);
    return AFB_ERR_UNSUPPORTED;
  }
  return 0;
}
