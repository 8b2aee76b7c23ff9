// libkon_b200 C-ABI plumbing: version, thread-local error string, device attribute cache.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace kon {

char* tls_error_buf() {
  static thread_local char buf[kErrLen] = {0};
  return buf;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-kernel timing ---------------------------------------------------------------------
namespace {
struct ProfRec { cudaEvent_t e0, e1; };
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::map<std::string, std::vector<ProfRec>> g_prof;
struct ProfTok { std::string name; cudaEvent_t e0; };
}  // namespace
bool profile_on() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
void profile_begin(const char* name, cudaStream_t st, void** tok) {
  auto* t = new ProfTok{name, nullptr};
  if (cudaEventCreate(&t->e0) != cudaSuccess) { delete t; return; }
  cudaEventRecord(t->e0, st);
  *tok = t;
}
void profile_end(void* tok, cudaStream_t st) {
  auto* t = static_cast<ProfTok*>(tok);
  cudaEvent_t e1;
  if (cudaEventCreate(&e1) == cudaSuccess) {
    cudaEventRecord(e1, st);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof[t->name].push_back(ProfRec{t->e0, e1});
  }
  delete t;
}

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

extern "C" int kon_profile_enable(int on) {
  kon::g_prof_on.store(on ? 1 : 0);
  return KON_OK;
}

extern "C" int kon_profile_reset(void) {
  std::lock_guard<std::mutex> lock(kon::g_prof_mu);
  for (auto& kv : kon::g_prof)
    for (auto& r : kv.second) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  kon::g_prof.clear();
  return KON_OK;
}

extern "C" int kon_profile_read(const char* kernel, double* total_ms, long long* launches) {
  KON_REQUIRE(kernel && total_ms && launches, KON_EINVAL, "kon_profile_read: NULL argument");
  std::lock_guard<std::mutex> lock(kon::g_prof_mu);
  *total_ms = 0.0;
  *launches = 0;
  auto it = kon::g_prof.find(kernel);
  if (it == kon::g_prof.end()) return KON_OK;
  for (auto& r : it->second) {
    KON_CUDA(cudaEventSynchronize(r.e1));
    float ms = 0.f;
    KON_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    *total_ms += ms;
    ++*launches;
  }
  return KON_OK;
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
