#include <stdlib.h>
// Library runtime: thread-local error string, device attribute cache, version.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace b200rl {

static thread_local char g_last_error[1024] = "";
long long g_launch_count = 0;
static int pdl_default() {
  const char* e = getenv("B200RL_PDL");
  return (e && e[0] == '0') ? 0 : 1;
}
int g_pdl_enabled = pdl_default();

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200rl

extern "C" const char* b200rl_last_error(void) { return b200rl::g_last_error; }

extern "C" int b200rl_version(void) { return 100; }

extern "C" long long b200rl_launch_count(void) { return b200rl::g_launch_count; }

// Returns 0 when the current device is an sm_100 part (the only target of this library).
extern "C" int b200rl_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  B200RL_CUDA_OK(cudaGetDevice(&dev));
  B200RL_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B200RL_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10)
    return b200rl::set_error(b200rl::B200RL_ERR_UNSUPPORTED,
                             "libb200rl is built for sm_100a only; device %d is sm_%d%d", dev, major,
                             minor);
  return 0;
}

// 1 (default; env B200RL_PDL=0 disables): hot-path kernels use programmatic dependent launch (common.cuh)
extern "C" int b200rl_set_pdl(int enable) {
  b200rl::g_pdl_enabled = enable ? 1 : 0;
  return 0;
}
