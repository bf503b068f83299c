// api.cu -- error reporting and device queries of libb200woq.
#include <atomic>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace b200woq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int num_sms() {
  static int cached[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(mu);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200woq

namespace b200woq { long long launches(); }
extern "C" int64_t b200woq_launch_count(void) { return (int64_t)b200woq::launches(); }

extern "C" int b200woq_version(void) { return B200WOQ_VERSION; }

extern "C" const char* b200woq_last_error(void) { return b200woq::g_err; }

extern "C" int b200woq_device_arch(char* out, int out_len) {
  int dev = 0, major = 0, minor = 0;
  WOQ_CUDA(cudaGetDevice(&dev));
  WOQ_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  WOQ_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (out && out_len > 0) snprintf(out, out_len, "sm_%d%d", major, minor);
  return 0;
}
