// Error plumbing and device checks for the C ABI.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

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
// Ping-pong traversal order of the streaming kernels: every kernel of the step reads what its predecessor wrote, each tensor
// (92-370 MB) is larger than the 126 MB L2, and all kernels used to walk their rows in ascending order -- so a consumer's first
// rows had long been evicted while the LAST third its producer wrote (still in L2) was overwritten before the consumer got
// there.  Alternating the direction launch by launch makes every consumer start where its producer stopped.
// AFB_PINGPONG=0 keeps every kernel ascending (read per call).
// AFB_PDL: 0 = never, 1 = every supported launch, unset = the short launches only (read per call)
bool pdl_enabled() {
  const char* v = getenv("AFB_PDL");
  return v == nullptr || v[0] != '0';
}
bool pdl_force_all() {
  const char* v = getenv("AFB_PDL");
  return v != nullptr && v[0] == '1';
}
static int g_dir = 0;
int next_stream_dir() {
  const char* v = getenv("AFB_PINGPONG");
  if (v != nullptr && v[0] == '0') return 0;
  g_dir ^= 1;
  return g_dir;
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
