// attn_tc5_probe: measured prototype of the AutoInt attention core  S = q' K^T -> P = sigmoid(S) -> O = P K  on
// tcgen05 ("4 samples x 32 padded fields per M = 128 tile"), to settle whether the 5th-gen tensor
// cores beat the warp-level mma.sync kernel (attn_tc.cu) at 26 x 8 per-sample matrices.
//
// One CTA = 128 threads = one 128-row tile = 4 samples x 32 (padded) fields; thread r owns TMEM lane r.
//   S   : ONE tcgen05.mma  M128 N128 K16  (A = q' rows [128 x 8, zero-padded to K = 16] in shared memory,
//         B = the K rows of the same 4 samples) -- the 128 x 128 result holds the four wanted 32 x 32 blocks on its
//         diagonal and 12 cross-sample blocks nobody reads (75 % of the MMA is waste).
//   P   : warp w (= sample w) reads its own 32 columns of its 32 lanes (tcgen05.ld x32), 32 sigmoids per thread,
//         packs to bf16 and stores them as the A operand of the next product (tcgen05.st x16) into the sample's own
//         16 words of a block-diagonal [128 x 128] A (the other 48 words of a lane stay zero).
//   O   : 8 tcgen05.mma  M128 N16 K16  (A = P from tensor memory, B = K^T [16 x 128] in shared memory);
//         tcgen05.ld x8 of the result, 32-byte store per row.
// Persistent CTAs, 2 per SM (256 TMEM columns each), no pipelining inside a CTA: the point is the cost of the
// TMEM round trips (ld 32 + st 16 + ld 8 words per row) around MMAs this small.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o attn_tc5_probe attn_tc5_probe.cu && ./attn_tc5_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ml_function_b200/csrc/tc_ptx.cuh"

namespace kon {
char* tls_error_buf() {
  static thread_local char buf[kErrLen];
  return buf;
}
void count_launch() {}
int sm_count_of(int) { return 148; }
}  // namespace kon

using namespace kon;

__device__ __forceinline__ float sigmoid_tanh(float z) {      // 0.5 + 0.5 tanh(z/2): one MUFU.TANH (as attn_tc.cu)
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}

constexpr uint32_t kColS = 0, kColP = 128, kColO = 192;   // TMEM columns: S fp32 [128], P bf16x2 [64], O fp32 [16]

// q, k: [B, 32, 8] bf16 (rows >= F are zero); out: [B, 32, 8] fp32
__global__ void __launch_bounds__(128) attn_tc5_kernel(const __nv_bfloat16* __restrict__ q,
                                                       const __nv_bfloat16* __restrict__ k, float* __restrict__ out,
                                                       long long n_tiles) {
  __shared__ __align__(1024) unsigned char sA[128 * 16 * 2];      // q' tile, K-major core matrices, K = 16
  __shared__ __align__(1024) unsigned char sBs[128 * 16 * 2];     // K tile as B[n = row][k = e]
  __shared__ __align__(1024) unsigned char sBo[16 * 128 * 2];     // K tile as B[n = e][k = row]
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&s_tmem, 256);
  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    fence_mbar_init();
  }
  // zero everything once: the k = 8..15 halves, the e = 8..15 rows of sBo
  for (int i = tid; i < 128 * 16 * 2 / 16; i += 128) {
    reinterpret_cast<uint4*>(sA)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(sBs)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(sBo)[i] = make_uint4(0, 0, 0, 0);
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  {   // the block-diagonal P operand: zero all 64 words of the lane once, the own 16 are rewritten per tile
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) tc::st16(tmem + lane_base + kColP + 16 * c, z);
    tc::wait_st();
  }
  const uint32_t lbo = 128 * 16, sbo = 128;              // [128 x 16] operands: k-group stride 2048 B, 8-row groups 128 B
  const uint32_t lbo_o = 16 * 16, sbo_o = 128;           // [16 x 128] operand: k-group stride 256 B
  const uint64_t dA = tc::smem_desc(smem_u32(sA), lbo, sbo);
  const uint64_t dBs = tc::smem_desc(smem_u32(sBs), lbo, sbo);
  const uint64_t dBo = tc::smem_desc(smem_u32(sBo), lbo_o, sbo_o);
  const uint32_t idesc_s = tc::idesc_bf16(128, 128, 0, 0), idesc_o = tc::idesc_bf16(128, 16, 0, 0);
  const uint32_t adv_o = (2 * lbo_o) >> 4;               // one k-step (16 rows of K) further in sBo
  uint32_t phase = 0;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1) {
    // ---- stage q', K of the 4 samples: one 16-byte core-matrix row per thread ------------------------------------
    const long long row = tile * 128 + tid;
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(q) + row);
    const uint4 kv = __ldg(reinterpret_cast<const uint4*>(k) + row);
    *reinterpret_cast<uint4*>(sA + (tid / 8) * sbo + (tid % 8) * 16) = qv;
    *reinterpret_cast<uint4*>(sBs + (tid / 8) * sbo + (tid % 8) * 16) = kv;
    {
      const unsigned short* ke = reinterpret_cast<const unsigned short*>(&kv);
#pragma unroll
      for (int e = 0; e < 8; ++e)      // B[n = e][k = tid]
        *reinterpret_cast<unsigned short*>(sBo + (tid / 8) * lbo_o + (e % 8) * 16 + (tid % 8) * 2) = ke[e];
    }
    fence_proxy_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (warp == 0) {
      if (tc::elect_one()) {
        tc::mma_ss(tmem + kColS, dA, dBs, idesc_s, 0u);
        tc::commit(&bar_s);
      }
      __syncwarp();
    }
    mbar_wait(&bar_s, phase);
    tc::fence_after();
    // ---- P = sigmoid(S) on the sample's own 32 columns -> its 16 words of the A operand --------------------------
    uint32_t sv[32], pw[16];
    tc::ld32(tmem + lane_base + kColS + 32 * warp, sv);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float p0 = sigmoid_tanh(__uint_as_float(sv[2 * j]));
      const float p1 = sigmoid_tanh(__uint_as_float(sv[2 * j + 1]));
      pw[j] = tc::pack_bf16(p0, p1);
    }
    tc::st16(tmem + lane_base + kColP + 16 * warp, pw);
    tc::wait_st();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (warp == 0) {
      if (tc::elect_one()) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          tc::mma_ts(tmem + kColO, tmem + kColP + 8 * s, dBo + (uint64_t)(s * adv_o), idesc_o, s > 0 ? 1u : 0u);
        tc::commit(&bar_o);
      }
      __syncwarp();
    }
    mbar_wait(&bar_o, phase);
    tc::fence_after();
    uint32_t ov[8];
    tc::ld8(tmem + lane_base + kColO, ov);
    tc::wait_ld();
    float4* op = reinterpret_cast<float4*>(out + row * 8);
    op[0] = make_float4(__uint_as_float(ov[0]), __uint_as_float(ov[1]), __uint_as_float(ov[2]), __uint_as_float(ov[3]));
    op[1] = make_float4(__uint_as_float(ov[4]), __uint_as_float(ov[5]), __uint_as_float(ov[6]), __uint_as_float(ov[7]));
    tc::fence_before();
    __syncthreads();        // the next tile overwrites sA / sBs / sBo and the accumulators
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

int main(int argc, char** argv) {
  const long long B = argc > 1 ? atoll(argv[1]) : 65536;
  const int F = 26;
  const long long rows = B * 32, n_tiles = rows / 128;
  std::vector<__nv_bfloat16> q(rows * 8), k(rows * 8);
  srand(7);
  for (long long r = 0; r < rows; ++r)
    for (int e = 0; e < 8; ++e) {
      const bool on = (r % 32) < F;
      q[r * 8 + e] = __float2bfloat16(on ? (rand() % 2001 - 1000) / 1000.f : 0.f);
      k[r * 8 + e] = __float2bfloat16(on ? (rand() % 2001 - 1000) / 1000.f : 0.f);
    }
  __nv_bfloat16 *dq, *dk;
  float* dout;
  cudaMalloc(&dq, q.size() * 2);
  cudaMalloc(&dk, k.size() * 2);
  cudaMalloc(&dout, rows * 8 * 4);
  cudaMemcpy(dq, q.data(), q.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dk, k.data(), k.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, rows * 8 * 4);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = (int)std::min<long long>(n_tiles, 2LL * sms);
  attn_tc5_kernel<<<grid, 128>>>(dq, dk, dout, n_tiles);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error %s\n", cudaGetErrorString(e));
    return 1;
  }
  // ---- check a few samples against the CPU (bf16 P like the kernel) -----------------------------------------------
  std::vector<float> out(rows * 8);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (long long b : {0LL, 1LL, 2LL, 3LL, 5LL, B / 2, B - 1}) {
    for (int i = 0; i < 32; ++i) {
      double o[8] = {0};
      for (int j = 0; j < 32; ++j) {
        double s = 0;
        for (int e2 = 0; e2 < 8; ++e2)
          s += (double)__bfloat162float(q[(b * 32 + i) * 8 + e2]) * __bfloat162float(k[(b * 32 + j) * 8 + e2]);
        const float p = __bfloat162float(__float2bfloat16((float)(1.0 / (1.0 + exp(-s)))));
        for (int e2 = 0; e2 < 8; ++e2) o[e2] += (double)p * __bfloat162float(k[(b * 32 + j) * 8 + e2]);
      }
      for (int e2 = 0; e2 < 8; ++e2) {
        maxerr = fmax(maxerr, fabs(o[e2] - out[(b * 32 + i) * 8 + e2]));
        maxref = fmax(maxref, fabs(o[e2]));
      }
    }
  }
  printf("check: max|err| %.4g (max|ref| %.3g) %s\n", maxerr, maxref, maxerr < 2e-2 * maxref ? "OK" : "MISMATCH");
  // ---- timing -----------------------------------------------------------------------------------------------------
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) attn_tc5_kernel<<<grid, 128>>>(dq, dk, dout, n_tiles);
  cudaEventRecord(e0);
  const int reps = 20;
  for (int i = 0; i < reps; ++i) attn_tc5_kernel<<<grid, 128>>>(dq, dk, dout, n_tiles);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("tcgen05 attention core (S, sigmoid, O; one head): B = %lld samples, %d CTAs: %.1f us per pass "
         "(x2 heads = %.1f us per AutoInt layer; the whole mma.sync forward kernel takes ~92 us per layer)\n",
         B, grid, ms / reps * 1e3, 2 * ms / reps * 1e3);
  return 0;
}
