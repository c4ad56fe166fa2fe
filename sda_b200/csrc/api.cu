// api.cu -- error reporting, device check and launch accounting of libsdab.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace sdab {

namespace {
thread_local std::string g_error;
// process-wide: autograd runs backward passes on its own thread
std::atomic<long long> g_launches{0};
}  // namespace

void set_error(const std::string& msg) { g_error = msg; }

int fail(int code, const std::string& msg) {
  g_error = msg;
  return code;
}

void count_launch(int n) { g_launches += n; }

// ---- optional per-launch timing of the convolution engine (bench.py's live roofline)
namespace {
struct ConvProfiler {
  bool enabled = false;
  std::vector<cudaEvent_t> events;  // pairs (begin, end)
  size_t used = 0;
  double flops = 0.0;
  long long launches = 0;
};
ConvProfiler g_prof;
std::mutex g_prof_mutex;
}  // namespace

void conv_profile_before(cudaStream_t st) {
  if (!g_prof.enabled) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  if (g_prof.used + 2 > g_prof.events.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      g_prof.events.push_back(e);
    }
  }
  cudaEventRecord(g_prof.events[g_prof.used], st);
}

void conv_profile_after(cudaStream_t st, double flops) {
  if (!g_prof.enabled) return;
  std::lock_guard<std::mutex> lock(g_prof_mutex);
  cudaEventRecord(g_prof.events[g_prof.used + 1], st);
  g_prof.used += 2;
  g_prof.flops += flops;
  g_prof.launches += 1;
}

}  // namespace sdab

extern "C" {

const char* sdab_last_error(void) { return sdab::g_error.c_str(); }

int sdab_version(void) { return 100; }

long long sdab_launch_count(int reset) {
  return reset ? sdab::g_launches.exchange(0) : sdab::g_launches.load();
}

int sdab_conv_profile(int enable) {
  std::lock_guard<std::mutex> lock(sdab::g_prof_mutex);
  sdab::g_prof.enabled = enable != 0;
  sdab::g_prof.used = 0;
  sdab::g_prof.flops = 0.0;
  sdab::g_prof.launches = 0;
  return SDAB_OK;
}

int sdab_conv_profile_read(double* ms, double* flops, long long* launches) {
  std::lock_guard<std::mutex> lock(sdab::g_prof_mutex);
  double total = 0.0;
  for (size_t i = 0; i + 1 < sdab::g_prof.used; i += 2) {
    if (cudaEventSynchronize(sdab::g_prof.events[i + 1]) != cudaSuccess)
      return sdab::fail(SDAB_ERR_DEVICE, "cudaEventSynchronize failed while reading the conv profile");
    float t = 0.f;
    cudaEventElapsedTime(&t, sdab::g_prof.events[i], sdab::g_prof.events[i + 1]);
    total += t;
  }
  if (ms) *ms = total;
  if (flops) *flops = sdab::g_prof.flops;
  if (launches) *launches = sdab::g_prof.launches;
  return SDAB_OK;
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
