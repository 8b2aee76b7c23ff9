// 8e: peer-memory plumbing for the sharded embedding exchange over NVLink 5 / NVSwitch.
//
// One process per GPU.  Each rank allocates ONE exchange region with kon_peer_alloc (plain
// cudaMalloc, so the legacy CUDA IPC handle works in any container), the 64-byte handles travel
// through the host-side process group, and every rank maps the other ranks' regions with
// kon_peer_open.  The gather / scatter kernels of embed.cu then store to / load from those
// mappings directly (kon_embed_fwd_peer, kon_embed_bwd_peer): the all-to-all of pooled rows IS
// the gather kernel's store stream, no staging buffer and no separate collective.
//
// kon_peer_barrier is the only synchronisation the exchange needs: a one-CTA kernel in which
// thread q releases a flag into rank q's region and acquires rank q's flag in its own region.
// It lives on the caller's stream, is capturable in a CUDA graph (the epoch is kept in device
// memory) and cannot hang the GPU: a rank that never shows up trips a timeout that is
// reported through the flag block instead of spinning forever.
#include <string.h>

#include "common.cuh"

namespace kon {
namespace {

// flag block layout (uint32 words): [0,16) arrival slots, 16 epoch, 17 error
constexpr int kMaxPuts = KON_MAX_PUTS;
constexpr int kEpochWord = 16;
constexpr int kErrorWord = 17;

struct FlagTable {
  uint32_t* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const __grid_constant__ FlagTable ft, int n, int rank,
                                                          unsigned long long timeout_ns) {
  uint32_t* mine = ft.p[rank];
  const int t = threadIdx.x;
  const uint32_t e = mine[kEpochWord] + 1;   // every thread reads it before thread 0 bumps it below
  __syncwarp();
  if (t < n) {
    // everything this rank wrote into peer memory in earlier kernels of the stream happens-before
    // the release below (kernel boundary + cumulativity of the system-scope fence)
    __threadfence_system();
    st_release_sys(ft.p[t] + rank, e);
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(mine + t) - e) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        mine[kErrorWord] = 1u + (uint32_t)t;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  if (t == 0) mine[kEpochWord] = e;
}

// ---- generic all-to-all payload mover: up to kMaxPuts strided 2-D copies, sources local, destinations anywhere
// (peer mappings) -- 16-byte units when every pointer / pitch / width allows it, else 4-byte units.
// Stores are fire-and-forget, so the remote side runs at NVLink bandwidth instead of at load latency.
struct PutTable {
  const char* src[kMaxPuts];
  char* dst[kMaxPuts];
  long long src_pitch[kMaxPuts], dst_pitch[kMaxPuts];
  long long upr[kMaxPuts];        // units per row
  long long units[kMaxPuts];      // rows * upr
  int n;
};

template <typename U>
__global__ void __launch_bounds__(256) peer_put2d_kernel(const __grid_constant__ PutTable t) {
  const int d = blockIdx.y;
  const long long total = t.units[d];
  const long long upr = t.upr[d];
  const char* __restrict__ src = t.src[d];
  char* __restrict__ dst = t.dst[d];
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
    const long long r = u / upr;
    const long long c = u - r * upr;
    const U v = *reinterpret_cast<const U*>(src + r * t.src_pitch[d] + c * (long long)sizeof(U));
    *reinterpret_cast<U*>(dst + r * t.dst_pitch[d] + c * (long long)sizeof(U)) = v;
  }
}

}  // namespace
}  // namespace kon

using namespace kon;

extern "C" int kon_peer_put2d(const KonPut2D* puts, int32_t n, int device_id, void* stream) {
  KON_REQUIRE(puts != nullptr && n >= 1 && n <= kMaxPuts, KON_EINVAL, "kon_peer_put2d: n=%d outside [1,%d]", n, kMaxPuts);
  bool vec = true;
  long long max_units4 = 0;
  for (int i = 0; i < n; ++i) {
    const KonPut2D& q = puts[i];
    KON_REQUIRE(q.rows >= 0 && q.width >= 0, KON_EINVAL, "put %d: negative extent", i);
    if (q.rows == 0 || q.width == 0) continue;
    KON_REQUIRE(q.src != nullptr && q.dst != nullptr, KON_EINVAL, "put %d: NULL pointer", i);
    KON_REQUIRE(((uintptr_t)q.src | (uintptr_t)q.dst | (uintptr_t)q.src_pitch | (uintptr_t)q.dst_pitch | (uintptr_t)q.width) % 4 == 0,
                KON_EINVAL, "put %d: pointers, pitches and width must be multiples of 4 bytes", i);
    KON_REQUIRE(q.src_pitch >= q.width || q.rows == 1, KON_EINVAL, "put %d: src_pitch < width", i);
    if (((uintptr_t)q.src | (uintptr_t)q.dst | (uintptr_t)q.src_pitch | (uintptr_t)q.dst_pitch | (uintptr_t)q.width) % 16 != 0) vec = false;
    max_units4 = std::max<long long>(max_units4, q.rows * (q.width / 4));
  }
  if (max_units4 == 0) return KON_OK;
  const int unit = vec ? 16 : 4;
  PutTable t{};
  t.n = n;
  for (int i = 0; i < n; ++i) {
    const KonPut2D& q = puts[i];
    t.src[i] = static_cast<const char*>(q.src);
    t.dst[i] = static_cast<char*>(q.dst);
    t.src_pitch[i] = q.src_pitch;
    t.dst_pitch[i] = q.dst_pitch;
    t.upr[i] = q.width / unit;
    t.units[i] = (q.rows == 0 || q.width == 0) ? 0 : q.rows * (q.width / unit);
    if (t.upr[i] == 0) t.upr[i] = 1;
  }
  DeviceGuard guard(device_id);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long max_units = max_units4 * 4 / unit;
  const int gx = (int)std::max<long long>(1, std::min<long long>((max_units + 255) / 256, (long long)sm_count_of(device_id) * 8 / std::max(1, std::min(n, 8))));
  dim3 grid(gx, n);
  if (vec) peer_put2d_kernel<uint4><<<grid, 256, 0, st>>>(t);
  else peer_put2d_kernel<uint32_t><<<grid, 256, 0, st>>>(t);
  KON_LAUNCH_CHECK("peer_put2d_kernel");
  return KON_OK;
}

extern "C" int kon_peer_alloc(int device_id, size_t bytes, void** ptr, void* handle64) {
  KON_REQUIRE(ptr != nullptr && handle64 != nullptr && bytes > 0, KON_EINVAL, "kon_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle is 64 bytes");
  DeviceGuard guard(device_id);
  void* p = nullptr;
  KON_CUDA(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(KON_ECUDA, "kon_peer_alloc: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return KON_OK;
}

extern "C" int kon_peer_open(int device_id, const void* handle64, void** ptr) {
  KON_REQUIRE(ptr != nullptr && handle64 != nullptr, KON_EINVAL, "kon_peer_open: bad argument");
  DeviceGuard guard(device_id);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  KON_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return KON_OK;
}

extern "C" int kon_peer_close(int device_id, void* ptr) {
  DeviceGuard guard(device_id);
  if (ptr) KON_CUDA(cudaIpcCloseMemHandle(ptr));
  return KON_OK;
}

extern "C" int kon_peer_free(int device_id, void* ptr) {
  DeviceGuard guard(device_id);
  if (ptr) KON_CUDA(cudaFree(ptr));
  return KON_OK;
}

extern "C" int kon_peer_barrier(void* const* peer_flags, int32_t n_peers, int32_t rank, int device_id,
                                int64_t timeout_ms, void* stream) {
  KON_REQUIRE(peer_flags != nullptr && n_peers >= 1 && n_peers <= kMaxPeers && rank >= 0 && rank < n_peers,
              KON_EINVAL, "kon_peer_barrier: n_peers=%d rank=%d", n_peers, rank);
  FlagTable ft{};
  for (int q = 0; q < n_peers; ++q) {
    KON_REQUIRE(peer_flags[q] != nullptr, KON_EINVAL, "peer_flags[%d] is NULL", q);
    ft.p[q] = static_cast<uint32_t*>(peer_flags[q]);
  }
  DeviceGuard guard(device_id);
  const unsigned long long to = (unsigned long long)(timeout_ms > 0 ? timeout_ms : 10000) * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(ft, n_peers, rank, to);
  KON_LAUNCH_CHECK("peer_barrier_kernel");
  return KON_OK;
}
