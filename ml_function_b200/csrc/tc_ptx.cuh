// Thin inline-PTX wrappers for the Blackwell tensor-core path (tcgen05 / TMEM), sm_100a.
// Only what the CIN kernels need: TMEM alloc, tcgen05.mma (A from TMEM or smem, B from
// smem), commit -> mbarrier, 32x32b TMEM loads/stores, the fences, descriptors.
#pragma once

#include <cuda_bf16.h>

#include "common.cuh"

namespace kon {
namespace tc {

// ---- TMEM allocation (one warp, .sync.aligned) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}

__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void wait_st() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// mbarrier arrive once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---- descriptors -------------------------------------------------------------------------
// Instruction descriptor, kind::f16, bf16 x bf16 -> f32, A and B K-major (bit layout of the
// PTX ISA "instruction descriptor" table: c_format[4,6) a_format[7,10) b_format[10,13)
// a_major[15] b_major[16] n>>3 [17,23) m>>4 [24,29)).
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, int a_mn_major = 0,
                                                  int b_mn_major = 0) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Shared-memory matrix descriptor, no swizzle ("interleave"): 8 x 16-byte core matrices.
//   K-major : lbo = byte stride between the two core matrices along K,
//             sbo = byte stride between 8-row groups along M/N.
//   MN-major: lbo = byte stride between 8-element groups along M/N ... (see callers)
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[tmem] * B[smem]     (one thread issues)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- TMEM <-> registers, shape 32x32b: lane = TMEM lane (row), consecutive columns -------
#define KON_R4(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3])
#define KON_R8(v, o) KON_R4(v, o), KON_R4(v, o + 4)
#define KON_R16(v, o) KON_R8(v, o), KON_R8(v, o + 8)
#define KON_W4(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3])
#define KON_W8(v, o) KON_W4(v, o), KON_W4(v, o + 4)
#define KON_W16(v, o) KON_W8(v, o), KON_W8(v, o + 8)

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : KON_R16(v, 0), KON_R16(v, 16)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : KON_R16(v, 0)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : KON_R8(v, 0)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      KON_W16(v, 0), KON_W16(v, 16)
      : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      KON_W16(v, 0)
      : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               KON_W8(v, 0)
               : "memory");
}
__device__ __forceinline__ void st4(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr),
               KON_W4(v, 0)
               : "memory");
}

// Store NW packed words (NW % 4 == 0) of this lane's row starting at column taddr.
template <int NW>
__device__ __forceinline__ void st_words(uint32_t taddr, const uint32_t* v) {
  static_assert(NW % 4 == 0, "st_words: multiple of 4");
  int o = 0;
#pragma unroll
  for (int i = 0; i < NW / 32; ++i, o += 32) st32(taddr + o, v + o);
  if constexpr ((NW % 32) >= 16) { st16(taddr + o, v + o); o += 16; }
  if constexpr ((NW % 16) >= 8) { st8(taddr + o, v + o); o += 8; }
  if constexpr ((NW % 8) >= 4) { st4(taddr + o, v + o); o += 4; }
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace tc
}  // namespace kon
