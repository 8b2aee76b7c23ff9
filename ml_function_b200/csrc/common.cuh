// Shared host/device helpers for libkon_b200: error reporting across the C-ABI,
// DLTensor validation, small device utilities.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/kon_b200.h"

namespace kon {

// ----------------------------------------------------------------------------------
// thread-local error string (kon_last_error)
// ----------------------------------------------------------------------------------
char* tls_error_buf();                       // defined in abi.cu
constexpr int kErrLen = 512;
constexpr int kMaxPeers = 16;                 // ranks of one NVLink domain a peer table can address (8e)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error_buf(), kErrLen, fmt, ap);
  va_end(ap);
  return code;
}

#define KON_REQUIRE(cond, code, ...)                  \
  do {                                                \
    if (!(cond)) return ::kon::fail((code), __VA_ARGS__); \
  } while (0)

#define KON_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess)                                                          \
      return ::kon::fail(KON_ECUDA, "%s failed: %s (%s:%d)", #expr,                 \
                         cudaGetErrorString(_e), __FILE__, __LINE__);               \
  } while (0)

void count_launch();   // abi.cu: bumps the process-wide launch counter (kon_launch_count)

#define KON_LAUNCH_CHECK(name)                                                      \
  do {                                                                              \
    ::kon::count_launch();                                                          \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess)                                                          \
      return ::kon::fail(KON_ECUDA, "launch of %s failed: %s", name,                \
                         cudaGetErrorString(_e));                                   \
  } while (0)

#define KON_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != KON_OK) return _rc; \
  } while (0)

// ----------------------------------------------------------------------------------
// optional per-kernel device timing (kon_profile_*): the entry points that launch several
// kernels bracket their main kernels with CUDA events on the caller's stream
// ----------------------------------------------------------------------------------
bool profile_on();                                                  // abi.cu
void profile_begin(const char* name, cudaStream_t st, void** tok);  // abi.cu
void profile_end(void* tok, cudaStream_t st);                       // abi.cu
struct ProfileScope {
  void* tok = nullptr;
  cudaStream_t st;
  ProfileScope(const char* name, cudaStream_t s) : st(s) {
    if (profile_on()) profile_begin(name, s, &tok);
  }
  ~ProfileScope() {
    if (tok) profile_end(tok, st);
  }
};

// ----------------------------------------------------------------------------------
// DLTensor checks
// ----------------------------------------------------------------------------------
inline bool is_dtype(const DLTensor* t, int code, int bits) {
  return t->dtype.code == code && t->dtype.bits == bits && t->dtype.lanes == 1;
}
inline bool is_f32(const DLTensor* t) { return is_dtype(t, kDLFloat, 32); }
inline bool is_bf16(const DLTensor* t) { return is_dtype(t, kDLBfloat, 16); }
inline bool is_i32(const DLTensor* t) { return is_dtype(t, kDLInt, 32); }
inline bool is_i64(const DLTensor* t) { return is_dtype(t, kDLInt, 64); }
inline bool is_u8(const DLTensor* t) { return is_dtype(t, kDLUInt, 8); }

inline int64_t stride_of(const DLTensor* t, int d) {
  if (t->strides) return t->strides[d];
  int64_t s = 1;
  for (int i = t->ndim - 1; i > d; --i) s *= t->shape[i];
  return s;
}
inline bool is_compact(const DLTensor* t) {
  int64_t s = 1;
  for (int i = t->ndim - 1; i >= 0; --i) {
    if (t->shape[i] != 1 && stride_of(t, i) != s) return false;
    s *= t->shape[i];
  }
  return true;
}
inline int64_t numel(const DLTensor* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}
template <typename T>
inline T* data_ptr(const DLTensor* t) {
  return reinterpret_cast<T*>(static_cast<char*>(t->data) + t->byte_offset);
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline int check_cuda_tensor(const DLTensor* t, const char* name, int device_id = -1) {
  KON_REQUIRE(t != nullptr, KON_EINVAL, "%s is NULL", name);
  KON_REQUIRE(t->device.device_type == kDLCUDA, KON_EDEVICE,
              "%s is not a CUDA tensor (device_type %d); libkon_b200 has no CPU path", name,
              (int)t->device.device_type);
  KON_REQUIRE(device_id < 0 || t->device.device_id == device_id, KON_EDEVICE,
              "%s lives on cuda:%d, expected cuda:%d", name, t->device.device_id, device_id);
  KON_REQUIRE(t->data != nullptr || numel(t) == 0, KON_EINVAL, "%s has a NULL data pointer", name);
  return KON_OK;
}

// Per-device attribute cache (immutable after first use).
int sm_count_of(int device_id);  // defined in abi.cu

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    ok = cudaGetDevice(&prev) == cudaSuccess;
    if (ok && prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// ----------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map) -------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-B aligned.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
#endif  // __CUDACC__

}  // namespace kon
