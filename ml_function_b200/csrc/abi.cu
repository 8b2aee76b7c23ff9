// libkon_b200 C-ABI plumbing: version, thread-local error string, device attribute cache.
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace kon {

char* tls_error_buf() {
  static thread_local char buf[kErrLen] = {0};
  return buf;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count_of(int device_id) {
  static std::mutex mu;
  static int cache[64] = {0};
  if (device_id < 0 || device_id >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cache[device_id] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device_id) != cudaSuccess || v <= 0)
      v = 148;
    cache[device_id] = v;
  }
  return cache[device_id];
}

}  // namespace kon

extern "C" int kon_abi_version(void) { return KON_ABI_VERSION; }

extern "C" long long kon_launch_count(void) {
  return kon::g_launches.load(std::memory_order_relaxed);
}

extern "C" const char* kon_last_error(void) { return kon::tls_error_buf(); }

extern "C" int kon_device_info(int device_id, int* sm_count, int* cc_major, int* cc_minor) {
  int sm = 0, maj = 0, min = 0;
  KON_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device_id));
  KON_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device_id));
  KON_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, device_id));
  if (sm_count) *sm_count = sm;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return KON_OK;
}
