// vaw_core.cu — library-wide plumbing: last-error string, version, device queries.
#include "vaw_common.cuh"
#include <stdarg.h>

#include <atomic>
namespace {
thread_local char g_err[1024] = "";
std::atomic<unsigned long long> g_launches{0};
}

void vaw_note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// number of kernels this library has launched in this process (all threads)
extern "C" unsigned long long vaw_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void vaw_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int vaw_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    sms = n;
  }
  return sms;
}

extern "C" const char* vaw_last_error(void) { return g_err; }

extern "C" int vaw_version(void) { return 100; }  // 0.1.0

// 0 if the current device can run the library's kernels (compute capability 10.x), negative otherwise.
extern "C" int vaw_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  VAW_CUDA_TRY(cudaGetDevice(&dev));
  VAW_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  VAW_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    vaw_set_error("vaw_b200 is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return VAW_ERR_UNSUPPORTED;
  }
  return VAW_OK;
}
