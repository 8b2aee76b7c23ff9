// tc_probe: pins the tcgen05 descriptor / TMEM-operand conventions the CIN kernels rely on,
// on real hardware, before they are used in anger.  One CTA computes D[128,N] = A[128,K] *
// B[N,K]^T (bf16 in, fp32 out) and the host compares with a CPU reference.
//
//   tc_probe <variant> [N] [K]
//     bit0  A operand: 0 = shared memory (K-major, no swizzle), 1 = tensor memory (packed bf16x2)
//     bit1  B operand: 0 = K-major, 1 = MN-major (no swizzle)
//     bit2  swap LBO/SBO in the descriptors (hypothesis test)
//     bit3  TMEM A packing: 0 = low half-word holds the even k, 1 = the odd k
//     bit4  timing mode: issue 16384 MMAs back to back (K = 64) and report cycles per MMA
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

struct Params {
  const __nv_bfloat16* A;   // [128,K] row-major
  const __nv_bfloat16* B;   // [N,K] row-major
  float* D;                 // [128,N]
  int N, K, variant;
  long long* cycles;
};

__global__ void __launch_bounds__(128) probe_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid / 32;
  const int N = p.N, K = p.K;
  const bool a_tmem = p.variant & 1, b_mn = p.variant & 2, swap = p.variant & 4, podd = p.variant & 8;
  const bool timing = p.variant & 16;

  unsigned char* sA = smem;                              // 128*K*2 bytes
  unsigned char* sB = smem + 128 * K * 2;                // N*K*2 bytes
  const uint32_t lboA = 128 * 16, sboA = 128;            // K-major: k-group stride, m-group stride
  const uint32_t lboB = (uint32_t)N * 16, sboB = 128;    // both majors: 8-wide "other dim" groups are contiguous

  if (warp == 0) tc::tmem_alloc(&s_tmem, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  // ---- stage operands -------------------------------------------------------------------
  for (int idx = tid; idx < 128 * K; idx += 128) {
    const int m = idx / K, k = idx % K;
    const uint32_t off = (k / 8) * lboA + (m / 8) * sboA + (m % 8) * 16 + (k % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sA + off) = p.A[idx];
  }
  for (int idx = tid; idx < N * K; idx += 128) {
    const int n = idx / K, k = idx % K;
    uint32_t off;
    if (!b_mn) off = (k / 8) * lboB + (n / 8) * sboB + (n % 8) * 16 + (k % 8) * 2;
    else       off = (k / 8) * lboB + (n / 8) * sboB + (k % 8) * 16 + (n % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = p.B[idx];
  }
  fence_proxy_async_smem();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t d_col = 0, a_col = 256;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (a_tmem) {
    const int m = tid;
    for (int s = 0; s < K / 16; ++s) {
      uint32_t w[8];
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat16 e = p.A[m * K + 16 * s + 2 * j], o = p.A[m * K + 16 * s + 2 * j + 1];
        const uint32_t eb = *reinterpret_cast<const unsigned short*>(&e);
        const uint32_t ob = *reinterpret_cast<const unsigned short*>(&o);
        w[j] = podd ? (ob | (eb << 16)) : (eb | (ob << 16));
      }
      tc::st8(tmem + lane_base + a_col + 8 * s, w);
    }
    tc::wait_st();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
  }
  // ---- MMA ------------------------------------------------------------------------------
  const uint32_t idesc = tc::idesc_bf16(128, N, 0, b_mn ? 1 : 0);
  long long t0 = 0;
  if (warp == 0) {
    // warp-uniform control flow, one elected lane issues (keeps operands in uniform registers)
    const bool leader = tc::elect_one();
    const int reps = timing ? 4096 : 1;
    const uint64_t da0 = swap ? tc::smem_desc(smem_u32(sA), sboA, lboA) : tc::smem_desc(smem_u32(sA), lboA, sboA);
    const uint64_t db0 = swap ? tc::smem_desc(smem_u32(sB), sboB, lboB) : tc::smem_desc(smem_u32(sB), lboB, sboB);
    const uint32_t adv_a = (2 * lboA) >> 4, adv_b = (2 * lboB) >> 4;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (leader) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {            // K == 64 in timing mode
          if (s < K / 16) {
            const uint32_t acc = (s > 0 || r > 0) ? 1u : 0u;
            if (a_tmem) tc::mma_ts(tmem + d_col, tmem + a_col + 8 * s, db0 + (uint64_t)(s * adv_b), idesc, acc);
            else        tc::mma_ss(tmem + d_col, da0 + (uint64_t)(s * adv_a), db0 + (uint64_t)(s * adv_b), idesc, acc);
          }
        }
      }
      __syncwarp();
    }
    if (leader) tc::commit(&bar);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc::fence_after();
  if (tid == 0 && p.cycles) *p.cycles = clock64() - t0;
  // ---- epilogue: lane = row -----------------------------------------------------------------
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    tc::ld8(tmem + lane_base + d_col + c0, v);
    tc::wait_ld();
    for (int j = 0; j < 8; ++j) p.D[tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = argc > 2 ? atoi(argv[2]) : 208;
  const int K = argc > 3 ? atoi(argv[3]) : 64;
  std::vector<__nv_bfloat16> A(128 * K), B(N * K);
  std::vector<float> Af(128 * K), Bf(N * K), D(128 * N), R(128 * N);
  srand(1234);
  for (int i = 0; i < 128 * K; ++i) { A[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Af[i] = __bfloat162float(A[i]); }
  for (int i = 0; i < N * K; ++i) { B[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); Bf[i] = __bfloat162float(B[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)Af[m * K + k] * Bf[n * K + k];
      R[m * N + n] = (float)acc;
    }
  Params p;
  __nv_bfloat16 *dA, *dB;
  float* dD;
  long long* dC;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  p.A = dA; p.B = dB; p.D = dD; p.N = N; p.K = K; p.variant = variant; p.cycles = dC;
  const size_t smem = (size_t)(128 + N) * K * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d N %d K %d: CUDA error %s\n", variant, N, K, cudaGetErrorString(e)); return 1; }
  long long cyc = 0;
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < D.size(); ++i) { maxerr = fmax(maxerr, fabs((double)D[i] - R[i])); maxref = fmax(maxref, fabs((double)R[i])); }
  if (variant & 16) {
    const int n_mma = 4096 * (K / 16 < 4 ? K / 16 : 4);
    printf("variant %d N %d K %d: TIMING %lld cycles for %d MMAs = %.1f cyc/MMA\n", variant, N, K, cyc, n_mma, (double)cyc / n_mma);
  } else {
    printf("variant %d N %d K %d: max|err| %.4g (max|ref| %.3g) %s\n", variant, N, K, maxerr, maxref,
           maxerr < 1e-2 * maxref ? "OK" : "MISMATCH");
  }
  return 0;
}
