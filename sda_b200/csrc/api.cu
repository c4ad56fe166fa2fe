// api.cu -- error reporting, device check and launch accounting of libsdab.
#include "common.cuh"

namespace sdab {

namespace {
thread_local std::string g_error;
thread_local long long g_launches = 0;
}  // namespace

void set_error(const std::string& msg) { g_error = msg; }

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

void count_launch(int n) { g_launches += n; }

}  // namespace sdab

extern "C" {

const char* sdab_last_error(void) { return sdab::g_error.c_str(); }

int sdab_version(void) { return 100; }

long long sdab_launch_count(int reset) {
  const long long v = sdab::g_launches;
  if (reset) sdab::g_launches = 0;
  return v;
}

int sdab_device_check(void) {
  static thread_local int cached_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return sdab::fail(SDAB_ERR_DEVICE, "no CUDA device is available: libsdab has no CPU fallback");
  }
  if (dev == cached_dev) return SDAB_OK;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    cudaGetLastError();
    return sdab::fail(SDAB_ERR_DEVICE, "cannot query the CUDA device: libsdab has no CPU fallback");
  }
  if (major != 10)
    return sdab::fail(SDAB_ERR_DEVICE, "libsdab is built for sm_100a (B200) only; current device has compute capability " +
                                           std::to_string(major) + ".x");
  cached_dev = dev;
  return SDAB_OK;
}

}  // extern "C"
