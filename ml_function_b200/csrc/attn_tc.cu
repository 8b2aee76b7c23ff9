// a9-a10 on tensor cores (KON_ATTN_BF16): the AutoInt block with bf16 operands / fp32 accumulate.
//
// The per-sample matrices are tiny (F = 26 fields, k_in = 16, d = 8): far below one tcgen05 tile
// (M = 128), and the op is HBM-bound on paper.  The fp32 kernel of attn.cu is instruction-bound on
// the CUDA cores (62 % issue utilisation, ~4000 warp instructions per sample); here one warp owns one
// sample and runs the whole block as 36 warp-level m16n8k16 MMAs, FlashAttention-2 style:
//   * X [32(pad) x k_in] -> A fragments straight from global memory (fp32 -> bf16x2);
//   * Q/K/R = X W: B fragments of W pre-packed once per call (attn_pack_w_kernel);
//   * S = Q K^T: the accumulator fragments of K ARE the col-major B fragments of K^T;
//   * P = sigmoid(S) (BL:286, fp32) -> the accumulator fragments of two adjacent n-tiles ARE the
//     A fragment of the next k-step;  O = P K needs K transposed: K goes through 512 B of shared
//     memory and comes back with ldmatrix.trans;
//   * LayerNorm (eps 1e-3) / residual / ReLU on the accumulator fragments, 8-byte stores of y.
// Legacy mma.sync on purpose: at 26x8 tiles tcgen05 would waste 75 % of every tile, and the goal is
// only to get the math off the critical path of an HBM-bound kernel.
#include <cuda_bf16.h>

#include "attn_common.cuh"

namespace kon {

namespace {

constexpr int kAtWarps = 8;
constexpr int kAtThreads = 32 * kAtWarps;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r0), "=r"(r1)
               : "r"(smem_u32(smem_row)));
}
// sigmoid(z) = 0.5 + 0.5 tanh(z/2): ONE special-function op (MUFU.TANH, abs. error ~2.5e-4 on the
// sigmoid, below the bf16 rounding of P) instead of ex2 + rcp -- the kernels below are bound by
// instruction issue and by the 4-lane/clk special-function unit, not by memory.
__device__ __forceinline__ float fast_sigmoid(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}

// B fragments of W[kin][H][8] for m16n8k16 (k = c, n = e), built by every CTA in shared memory:
//   word[((mat*H + h)*KS + ks)*64 + r*32 + lane],  r = 0: (W[16ks+2t][h][g], W[16ks+2t+1][h][g]),  r = 1: rows +8
__device__ __forceinline__ void pack_w_frags(const float* __restrict__ wq, const float* __restrict__ wk,
                                             const float* __restrict__ wr, int KS, int H, uint32_t* out,
                                             int tid, int nthreads) {
  const int total = 3 * H * KS * 64;
  for (int idx = tid; idx < total; idx += nthreads) {
    const int lane = idx & 31, r = (idx >> 5) & 1;
    int t = idx >> 6;
    const int ks = t % KS;
    t /= KS;
    const int h = t % H, mat = t / H;
    const float* W = mat == 0 ? wq : (mat == 1 ? wk : wr);
    const int g = lane >> 2, t4 = lane & 3;
    const int c = 16 * ks + 2 * t4 + 8 * r;
    const float lo = W ? W[((long long)c * H + h) * 8 + g] : 0.f;
    const float hi = W ? W[((long long)(c + 1) * H + h) * 8 + g] : 0.f;
    out[idx] = pack2(lo, hi);
  }
}

template <int KS, bool ALL>
__global__ void __launch_bounds__(kAtThreads)
attn_tc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wq, const float* __restrict__ wk,
                   const float* __restrict__ wr, const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                   const AttnDims p) {
  const bool f_scale = ALL || p.use_scale, f_res = ALL || p.use_res, f_ln = ALL || p.use_ln, f_relu = ALL || p.relu;   // cf. attn_tc_bwd_kernel
  extern __shared__ __align__(16) uint32_t smem_w[];                 // [3*H*KS*64] W fragments
  __shared__ __align__(16) unsigned short s_k[kAtWarps][32 * 8];     // per warp: K (bf16) [j][e]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int H = p.H, F = p.F, kin = 16 * KS;
  pack_w_frags(wq, wk, wr, KS, H, smem_w, tid, kAtThreads);
  __syncthreads();
  const float sc = f_scale ? rsqrtf(8.f) : 1.f;
  float gam[2], bet[2];
  gam[0] = f_ln ? gamma[2 * t] : 1.f;  gam[1] = f_ln ? gamma[2 * t + 1] : 1.f;
  bet[0] = f_ln ? beta[2 * t] : 0.f;   bet[1] = f_ln ? beta[2 * t + 1] : 0.f;
  unsigned short* ksm = s_k[warp];

  // X of the NEXT sample is loaded (raw fp32, 8 float2 per k-step) while the current one is processed: the loads at
  // the top of an iteration were consumed at once (their F2FP pack carried 31 % of the stall samples)
  float2 nx[2][KS][4];
  auto load_x = [&](long long bb) {
    const float* xb = x + min(bb, p.B - 1) * (long long)F * kin;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int q0 = min(16 * mt + g, F - 1), q1 = min(16 * mt + g + 8, F - 1);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int c0 = 16 * ks + 2 * t;
        nx[mt][ks][0] = __ldg(reinterpret_cast<const float2*>(xb + q0 * kin + c0));
        nx[mt][ks][1] = __ldg(reinterpret_cast<const float2*>(xb + q1 * kin + c0));
        nx[mt][ks][2] = __ldg(reinterpret_cast<const float2*>(xb + q0 * kin + c0 + 8));
        nx[mt][ks][3] = __ldg(reinterpret_cast<const float2*>(xb + q1 * kin + c0 + 8));
      }
    }
  };
  load_x((long long)blockIdx.x * kAtWarps + warp);
  for (long long b = (long long)blockIdx.x * kAtWarps + warp; b < p.B; b += (long long)gridDim.x * kAtWarps) {
    // ---- A fragments of X: rows 16mt+g (+8), cols 16ks+2t (+8); padded rows are exact zeros ----
    uint32_t ax[2][KS][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const bool on0 = 16 * mt + g < F, on1 = 16 * mt + g + 8 < F;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        ax[mt][ks][0] = on0 ? pack2(nx[mt][ks][0].x, nx[mt][ks][0].y) : 0u;
        ax[mt][ks][1] = on1 ? pack2(nx[mt][ks][1].x, nx[mt][ks][1].y) : 0u;
        ax[mt][ks][2] = on0 ? pack2(nx[mt][ks][2].x, nx[mt][ks][2].y) : 0u;
        ax[mt][ks][3] = on1 ? pack2(nx[mt][ks][3].x, nx[mt][ks][3].y) : 0u;
      }
    }
    load_x(b + (long long)gridDim.x * kAtWarps);
    for (int h = 0; h < H; ++h) {
      // ---- projections ----------------------------------------------------------------------
      float qc[2][4], kc[2][4], rc[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int q = 0; q < 4; ++q) qc[mt][q] = kc[mt][q] = rc[mt][q] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t* wq_ = smem_w + ((0 * H + h) * KS + ks) * 64;
        const uint32_t* wk_ = smem_w + ((1 * H + h) * KS + ks) * 64;
        const uint32_t* wr_ = smem_w + ((2 * H + h) * KS + ks) * 64;
        const uint32_t bq0 = wq_[lane], bq1 = wq_[32 + lane], bk0 = wk_[lane], bk1 = wk_[32 + lane];
        const uint32_t br0 = wr_[lane], br1 = wr_[32 + lane];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma16816(qc[mt], ax[mt][ks], bq0, bq1);
          mma16816(kc[mt], ax[mt][ks], bk0, bk1);
          if (f_res) mma16816(rc[mt], ax[mt][ks], br0, br1);
        }
      }
      // ---- K (bf16) -> shared [j][e] for the transposed read; Q / K fragments for S = Q K^T ----
      __syncwarp();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        *reinterpret_cast<uint32_t*>(ksm + (16 * mt + g) * 8 + 2 * t) = pack2(kc[mt][0], kc[mt][1]);
        *reinterpret_cast<uint32_t*>(ksm + (16 * mt + g + 8) * 8 + 2 * t) = pack2(kc[mt][2], kc[mt][3]);
      }
      __syncwarp();
      uint32_t aq[2][4], bkf[4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        aq[mt][0] = pack2(qc[mt][0] * sc, qc[mt][1] * sc);
        aq[mt][1] = pack2(qc[mt][2] * sc, qc[mt][3] * sc);
        aq[mt][2] = aq[mt][3] = 0u;                       // k = 8..15: padding of d = 8
        bkf[2 * mt] = pack2(kc[mt][0], kc[mt][1]);        // n-tile 2mt   : rows j = 16mt + g
        bkf[2 * mt + 1] = pack2(kc[mt][2], kc[mt][3]);    // n-tile 2mt+1 : rows j = 16mt + 8 + g
      }
      float o[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int q = 0; q < 4; ++q) s[nt][q] = 0.f;
          mma16816(s[nt], aq[mt], bkf[nt], 0u);
        }
        // P = sigmoid(S); two n-tiles -> one A fragment (k = 16 rows of K).  No mask on the padded
        // columns j >= F: the padded rows of X are exact zeros, so rows j >= F of K are zeros and
        // those columns of P contribute nothing to O = P K.
        uint32_t ap[2][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          ap[nt >> 1][(nt & 1) * 2] = pack2(fast_sigmoid(s[nt][0]), fast_sigmoid(s[nt][1]));        // row g
          ap[nt >> 1][(nt & 1) * 2 + 1] = pack2(fast_sigmoid(s[nt][2]), fast_sigmoid(s[nt][3]));    // row g + 8
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) o[mt][q] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint32_t bt0, bt1;
          ldmatrix_x2_trans(bt0, bt1, ksm + (16 * kk + (lane & 15)) * 8);
          mma16816(o[mt], ap[kk], bt0, bt1);
        }
      }
      // ---- LayerNorm over d = 8 (a row's 8 values live in the 4 lanes of a quad), residual, ReLU -----
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float v0 = o[mt][2 * half], v1 = o[mt][2 * half + 1];
          if (f_ln) {
            float sum = v0 + v1;
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            const float mean = sum * 0.125f;
            float var = (v0 - mean) * (v0 - mean) + (v1 - mean) * (v1 - mean);
            var += __shfl_xor_sync(0xffffffffu, var, 1);
            var += __shfl_xor_sync(0xffffffffu, var, 2);
            const float rstd = rsqrtf(var * 0.125f + p.ln_eps);
            v0 = (v0 - mean) * rstd * gam[0] + bet[0];
            v1 = (v1 - mean) * rstd * gam[1] + bet[1];
          }
          if (f_res) { v0 += rc[mt][2 * half]; v1 += rc[mt][2 * half + 1]; }
          if (f_relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          const int row = 16 * mt + g + 8 * half;
          if (row < F)
            *reinterpret_cast<float2*>(y + h * p.ysh + b * p.ysb + row * p.ysf + 2 * t) = make_float2(v0, v1);
        }
      }
    }
  }
}

// =============================================================================================
// backward
// =============================================================================================
// Everything of the forward is recomputed per head in fragments; then (q' = q / sqrt(d)):
//   gP = gy . relu'            LayerNorm backward -> gO            gR = gP
//   gS = gO K^T,  gZ = gS . P . (1 - P)                           gQ' = gZ K
//   transposed side: P and gZ are parked in shared memory as bf16 [i][j] (the words of their A
//   fragments) and P^T / gZ^T come back as A fragments through ldmatrix.trans:
//   gK = gZ^T q' + P^T gO
//   dX += (gQ'/sqrt(d)) Wq^T + gK Wk^T + gR Wr^T
//   dWq += X^T (gQ'/sqrt(d)),  dWk += X^T gK,  dWr += X^T gR       (accumulated per warp in registers)
// Operands that must be read "the other way round" (K, q', gO, X, gQ, gK, gR) go through a few
// hundred bytes of per-warp shared memory and come back with ldmatrix.trans.
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t* r, const void* smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
// C fragment [16 x 8] -> A fragment whose k = 0..7 are those 8 columns (k = 8..15 zero)
__device__ __forceinline__ void c_to_a_k8(const float* c, uint32_t* a, float scale = 1.f) {
  a[0] = pack2(c[0] * scale, c[1] * scale);
  a[1] = pack2(c[2] * scale, c[3] * scale);
  a[2] = a[3] = 0u;
}
// two C fragments of adjacent n-tiles -> one A fragment (k = 16)
__device__ __forceinline__ void cc_to_a(const float* c0, const float* c1, uint32_t* a) {
  a[0] = pack2(c0[0], c0[1]);
  a[1] = pack2(c0[2], c0[3]);
  a[2] = pack2(c1[0], c1[1]);
  a[3] = pack2(c1[2], c1[3]);
}
// store a C fragment [16 x 8] as bf16 rows [row][8] (row = 16mt + g, +8)
__device__ __forceinline__ void store_c_bf16(unsigned short* sm, int mt, int g, int t, const float* c, float scale = 1.f) {
  *reinterpret_cast<uint32_t*>(sm + (16 * mt + g) * 8 + 2 * t) = pack2(c[0] * scale, c[1] * scale);
  *reinterpret_cast<uint32_t*>(sm + (16 * mt + g + 8) * 8 + 2 * t) = pack2(c[2] * scale, c[3] * scale);
}

#ifndef KON_ATB_MINB
#define KON_ATB_MINB 4
#endif
constexpr int kAtBwdWarps = 4;
constexpr int kPStride = 40;   // bf16 per row of the parked P / gZ tiles: 80 B rows keep both the 32-bit fragment
                               // stores and the ldmatrix row reads free of bank conflicts
constexpr int kAtBwdThreads = 32 * kAtBwdWarps;

// ALL: use_scale, use_res, use_ln and the ReLU are all on (the AutoInt block as the builders make it, MD:159-163):
// the four run-time flags become constants and their selects / branches disappear from the 1800-instruction body.
template <int KS, int HT, bool ALL>
__global__ void __launch_bounds__(kAtBwdThreads, KON_ATB_MINB)
attn_tc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wq, const float* __restrict__ wk,
                   const float* __restrict__ wr, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ gy, float* __restrict__ dx, float* __restrict__ partial,
                   const AttnDims p) {
  constexpr int KIN = 16 * KS;
  extern __shared__ __align__(16) uint32_t smem_dyn[];
  // dynamic smem: [W fragments 3*H*KS*64] [W^T fragments 3*H*(KIN/8)*32]
  const int H = p.H, F = p.F;
  const bool f_scale = ALL || p.use_scale, f_res = ALL || p.use_res, f_ln = ALL || p.use_ln, f_relu = ALL || p.relu;
  uint32_t* w_frag = smem_dyn;
  uint32_t* wt_frag = w_frag + 3 * H * KS * 64;
  __shared__ __align__(16) unsigned short s_buf[kAtBwdWarps][6][32 * 8];     // K, q', gO, gQ, gK, gR  (bf16 [row][8])
  __shared__ __align__(16) unsigned short s_x[kAtBwdWarps][32 * KIN];         // X (bf16) [i][c]
  __shared__ __align__(16) unsigned short s_pz[kAtBwdWarps][2][32 * kPStride]; // P, gZ (bf16) [i][j], rows padded to 80 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  pack_w_frags(wq, wk, wr, KS, H, w_frag, tid, kAtBwdThreads);
  // W^T fragments (k = e, n = c): word[((mat*H + h)*(KIN/8) + ntc)*32 + lane] = (W[8ntc+g][h][2t], W[8ntc+g][h][2t+1])
  for (int idx = tid; idx < 3 * H * (KIN / 8) * 32; idx += kAtBwdThreads) {
    const int l = idx & 31;
    int u = idx >> 5;
    const int ntc = u % (KIN / 8);
    u /= (KIN / 8);
    const int h = u % H, mat = u / H;
    const float* W = mat == 0 ? wq : (mat == 1 ? wk : wr);
    const int c = 8 * ntc + (l >> 2), e = 2 * (l & 3);
    wt_frag[idx] = W ? pack2(W[((long long)c * H + h) * 8 + e], W[((long long)c * H + h) * 8 + e + 1]) : 0u;
  }
  __syncthreads();
  const float sc = f_scale ? rsqrtf(8.f) : 1.f;
  float gam[2], bet[2];
  gam[0] = f_ln ? gamma[2 * t] : 1.f;  gam[1] = f_ln ? gamma[2 * t + 1] : 1.f;
  bet[0] = f_ln ? beta[2 * t] : 0.f;   bet[1] = f_ln ? beta[2 * t + 1] : 0.f;
  unsigned short* ksm = s_buf[warp][0];
  unsigned short* qsm = s_buf[warp][1];
  unsigned short* gosm = s_buf[warp][2];
  unsigned short* gqsm = s_buf[warp][3];
  unsigned short* gksm = s_buf[warp][4];
  unsigned short* grsm = s_buf[warp][5];
  unsigned short* xsm = s_x[warp];
  unsigned short* psm = s_pz[warp][0];
  unsigned short* zsm = s_pz[warp][1];

  // per-warp accumulators: dW fragments (rows c = 16ks + g (+8), cols e = 2t, 2t+1) and dgamma/dbeta (cols 2t, 2t+1)
  float dwq[HT][KS][4], dwk[HT][KS][4], dwr[HT][KS][4];      // H <= HT
#pragma unroll
  for (int h = 0; h < HT; ++h)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) dwq[h][ks][q] = dwk[h][ks][q] = dwr[h][ks][q] = 0.f;
  float dgam[2] = {0.f, 0.f}, dbet[2] = {0.f, 0.f};

  for (long long b = (long long)blockIdx.x * kAtBwdWarps + warp; b < p.B; b += (long long)gridDim.x * kAtBwdWarps) {
    uint32_t ax[2][KS][4];
    const float* xb = x + b * (long long)F * KIN;
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int c0 = 16 * ks + 2 * t;
        float2 v00 = make_float2(0.f, 0.f), v10 = v00, v01 = v00, v11 = v00;
        if (r0 < F) {
          v00 = __ldg(reinterpret_cast<const float2*>(xb + r0 * KIN + c0));
          v01 = __ldg(reinterpret_cast<const float2*>(xb + r0 * KIN + c0 + 8));
        }
        if (r1 < F) {
          v10 = __ldg(reinterpret_cast<const float2*>(xb + r1 * KIN + c0));
          v11 = __ldg(reinterpret_cast<const float2*>(xb + r1 * KIN + c0 + 8));
        }
        ax[mt][ks][0] = pack2(v00.x, v00.y);
        ax[mt][ks][1] = pack2(v10.x, v10.y);
        ax[mt][ks][2] = pack2(v01.x, v01.y);
        ax[mt][ks][3] = pack2(v11.x, v11.y);
        *reinterpret_cast<uint32_t*>(xsm + r0 * KIN + c0) = ax[mt][ks][0];
        *reinterpret_cast<uint32_t*>(xsm + r1 * KIN + c0) = ax[mt][ks][1];
        *reinterpret_cast<uint32_t*>(xsm + r0 * KIN + c0 + 8) = ax[mt][ks][2];
        *reinterpret_cast<uint32_t*>(xsm + r1 * KIN + c0 + 8) = ax[mt][ks][3];
      }
    }
    float dxc[2][KIN / 8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < KIN / 8; ++n)
#pragma unroll
        for (int q = 0; q < 4; ++q) dxc[mt][n][q] = 0.f;

#pragma unroll
    for (int h = 0; h < HT; ++h) {
      if (h < H) {
        // the output gradient of this (sample, head): loaded here, consumed ~200 instructions below (the select
        // right behind these loads carried 35 % of the kernel's stall samples when they sat at their point of use)
        float2 gyv[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int row = min(16 * mt + g + 8 * half, F - 1);
            gyv[mt][half] = __ldg(reinterpret_cast<const float2*>(gy + h * p.gsh + b * p.gsb + row * p.gsf + 2 * t));
          }
        // ---- forward recompute ----------------------------------------------------------------
        float qc[2][4], kc[2][4], rc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int q = 0; q < 4; ++q) qc[mt][q] = kc[mt][q] = rc[mt][q] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t* wq_ = w_frag + ((0 * H + h) * KS + ks) * 64;
          const uint32_t* wk_ = w_frag + ((1 * H + h) * KS + ks) * 64;
          const uint32_t* wr_ = w_frag + ((2 * H + h) * KS + ks) * 64;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma16816(qc[mt], ax[mt][ks], wq_[lane], wq_[32 + lane]);
            mma16816(kc[mt], ax[mt][ks], wk_[lane], wk_[32 + lane]);
            if (f_res) mma16816(rc[mt], ax[mt][ks], wr_[lane], wr_[32 + lane]);
          }
        }
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          store_c_bf16(ksm, mt, g, t, kc[mt]);
          store_c_bf16(qsm, mt, g, t, qc[mt], sc);
        }
        __syncwarp();
        uint32_t aq[2][4], bkf[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          c_to_a_k8(qc[mt], aq[mt], sc);
          bkf[2 * mt] = pack2(kc[mt][0], kc[mt][1]);
          bkf[2 * mt + 1] = pack2(kc[mt][2], kc[mt][3]);
        }
        uint32_t kt[2][2], qt[2][2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          ldmatrix_x2_trans(kt[kk][0], kt[kk][1], ksm + (16 * kk + (lane & 15)) * 8);
          ldmatrix_x2_trans(qt[kk][0], qt[kk][1], qsm + (16 * kk + (lane & 15)) * 8);
        }
        // Row block mt (16 rows i) at a time: S -> P -> O -> LayerNorm/ReLU backward -> gO -> gS -> gZ -> gQ'.
        // No masks on the padded rows/columns (>= F): padded rows of X are exact zeros, so K, q', gy, gO
        // vanish there and every product that could see a padded P or gZ entry multiplies a zero.
        // P and gZ (bf16, the very words of their A fragments) are parked in shared memory [i][j] so
        // that the transposed products below read P^T / gZ^T back with ldmatrix.trans.
        float gq[2][4], gO[2][4], gR[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          float P[4][4];
          uint32_t ap[2][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(s4, aq[mt], bkf[nt], 0u);
#pragma unroll
            for (int q = 0; q < 4; ++q) P[nt][q] = fast_sigmoid(s4[q]);
            ap[nt >> 1][(nt & 1) * 2] = pack2(P[nt][0], P[nt][1]);
            ap[nt >> 1][(nt & 1) * 2 + 1] = pack2(P[nt][2], P[nt][3]);
            *reinterpret_cast<uint32_t*>(psm + (16 * mt + g) * kPStride + 8 * nt + 2 * t) = ap[nt >> 1][(nt & 1) * 2];
            *reinterpret_cast<uint32_t*>(psm + (16 * mt + g + 8) * kPStride + 8 * nt + 2 * t) = ap[nt >> 1][(nt & 1) * 2 + 1];
          }
          float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) mma16816(o, ap[kk], kt[kk][0], kt[kk][1]);
          // ---- output gradient through ReLU / residual / LayerNorm -> gO, gR ----------------------
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int row = 16 * mt + g + 8 * half;
            float v0 = o[2 * half], v1 = o[2 * half + 1];
            float mean = 0.f, rstd = 1.f;
            if (f_ln) {
              float sum = v0 + v1;
              sum += __shfl_xor_sync(0xffffffffu, sum, 1);
              sum += __shfl_xor_sync(0xffffffffu, sum, 2);
              mean = sum * 0.125f;
              float var = (v0 - mean) * (v0 - mean) + (v1 - mean) * (v1 - mean);
              var += __shfl_xor_sync(0xffffffffu, var, 1);
              var += __shfl_xor_sync(0xffffffffu, var, 2);
              rstd = rsqrtf(var * 0.125f + p.ln_eps);
            }
            const float xh0 = (v0 - mean) * rstd, xh1 = (v1 - mean) * rstd;
            float pre0 = f_ln ? xh0 * gam[0] + bet[0] : v0;
            float pre1 = f_ln ? xh1 * gam[1] + bet[1] : v1;
            if (f_res) { pre0 += rc[mt][2 * half]; pre1 += rc[mt][2 * half + 1]; }
            float g0 = 0.f, g1 = 0.f;
            if (row < F) {
              const float2 gv = gyv[mt][half];
              g0 = (f_relu && !(pre0 > 0.f)) ? 0.f : gv.x;
              g1 = (f_relu && !(pre1 > 0.f)) ? 0.f : gv.y;
            }
            gR[mt][2 * half] = f_res ? g0 : 0.f;
            gR[mt][2 * half + 1] = f_res ? g1 : 0.f;
            if (f_ln) {
              dgam[0] = fmaf(g0, xh0, dgam[0]); dgam[1] = fmaf(g1, xh1, dgam[1]);
              dbet[0] += g0; dbet[1] += g1;
              const float gx0 = g0 * gam[0], gx1 = g1 * gam[1];
              float m1 = gx0 + gx1, m2 = gx0 * xh0 + gx1 * xh1;
              m1 += __shfl_xor_sync(0xffffffffu, m1, 1);
              m1 += __shfl_xor_sync(0xffffffffu, m1, 2);
              m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
              m2 += __shfl_xor_sync(0xffffffffu, m2, 2);
              m1 *= 0.125f;
              m2 *= 0.125f;
              gO[mt][2 * half] = rstd * (gx0 - m1 - xh0 * m2);
              gO[mt][2 * half + 1] = rstd * (gx1 - m1 - xh1 * m2);
            } else {
              gO[mt][2 * half] = g0;
              gO[mt][2 * half + 1] = g1;
            }
          }
          uint32_t ago[4];
          c_to_a_k8(gO[mt], ago);
          *reinterpret_cast<uint32_t*>(gosm + (16 * mt + g) * 8 + 2 * t) = ago[0];
          *reinterpret_cast<uint32_t*>(gosm + (16 * mt + g + 8) * 8 + 2 * t) = ago[1];
          store_c_bf16(grsm, mt, g, t, gR[mt]);
          // ---- gZ = (gO K^T) . P . (1 - P);  gQ' = gZ K ------------------------------------------
          uint32_t az[2][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(s4, ago, bkf[nt], 0u);
#pragma unroll
            for (int q = 0; q < 4; ++q) s4[q] *= P[nt][q] * (1.f - P[nt][q]);
            az[nt >> 1][(nt & 1) * 2] = pack2(s4[0], s4[1]);
            az[nt >> 1][(nt & 1) * 2 + 1] = pack2(s4[2], s4[3]);
            *reinterpret_cast<uint32_t*>(zsm + (16 * mt + g) * kPStride + 8 * nt + 2 * t) = az[nt >> 1][(nt & 1) * 2];
            *reinterpret_cast<uint32_t*>(zsm + (16 * mt + g + 8) * kPStride + 8 * nt + 2 * t) = az[nt >> 1][(nt & 1) * 2 + 1];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) gq[mt][q] = 0.f;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) mma16816(gq[mt], az[kk], kt[kk][0], kt[kk][1]);
        }
        __syncwarp();
        // ---- transposed side: gK = gZ^T q' + P^T gO  (rows = j); A fragments of the transposes come
        // back from shared memory with ldmatrix.trans ------------------------------------------------
        uint32_t got[2][2];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) ldmatrix_x2_trans(got[kk][0], got[kk][1], gosm + (16 * kk + (lane & 15)) * 8);
        float gk[2][4];
#pragma unroll
        for (int mtj = 0; mtj < 2; ++mtj) {
#pragma unroll
          for (int q = 0; q < 4; ++q) gk[mtj][q] = 0.f;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            uint32_t a1[4], a2[4];
            const int m = lane >> 3;
            const int off = (16 * kk + 8 * (m >> 1) + (lane & 7)) * kPStride + 16 * mtj + 8 * (m & 1);
            ldmatrix_x4_trans(a1, zsm + off);
            ldmatrix_x4_trans(a2, psm + off);
            mma16816(gk[mtj], a1, qt[kk][0], qt[kk][1]);
            mma16816(gk[mtj], a2, got[kk][0], got[kk][1]);
          }
        }
        // ---- dX += [gQ | gK] [Wq^T ; Wk^T] + gR Wr^T ;  gQ = gQ' / sqrt(d) ---------------------------
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t a12[4], a3[4];
          a12[0] = pack2(gq[mt][0] * sc, gq[mt][1] * sc);
          a12[1] = pack2(gq[mt][2] * sc, gq[mt][3] * sc);
          a12[2] = pack2(gk[mt][0], gk[mt][1]);
          a12[3] = pack2(gk[mt][2], gk[mt][3]);
          c_to_a_k8(gR[mt], a3);
#pragma unroll
          for (int n = 0; n < KIN / 8; ++n) {
            mma16816(dxc[mt][n], a12, wt_frag[((0 * H + h) * (KIN / 8) + n) * 32 + lane],
                     wt_frag[((1 * H + h) * (KIN / 8) + n) * 32 + lane]);
            if (f_res) mma16816(dxc[mt][n], a3, wt_frag[((2 * H + h) * (KIN / 8) + n) * 32 + lane], 0u);
          }
          *reinterpret_cast<uint32_t*>(gqsm + (16 * mt + g) * 8 + 2 * t) = a12[0];
          *reinterpret_cast<uint32_t*>(gqsm + (16 * mt + g + 8) * 8 + 2 * t) = a12[1];
          *reinterpret_cast<uint32_t*>(gksm + (16 * mt + g) * 8 + 2 * t) = a12[2];
          *reinterpret_cast<uint32_t*>(gksm + (16 * mt + g + 8) * 8 + 2 * t) = a12[3];
        }
        __syncwarp();
        // ---- dW += X^T g*  (M = c, K = i, N = e) ------------------------------------------------------
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          uint32_t bq[2], bk[2], br[2];
          ldmatrix_x2_trans(bq[0], bq[1], gqsm + (16 * kk + (lane & 15)) * 8);
          ldmatrix_x2_trans(bk[0], bk[1], gksm + (16 * kk + (lane & 15)) * 8);
          ldmatrix_x2_trans(br[0], br[1], grsm + (16 * kk + (lane & 15)) * 8);
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t axt[4];     // X^T fragment: rows c = 16ks + .., k = i = 16kk + ..
            const int m = lane >> 3;
            ldmatrix_x4_trans(axt, xsm + (16 * kk + 8 * (m >> 1) + (lane & 7)) * KIN + 16 * ks + 8 * (m & 1));
            mma16816(dwq[h][ks], axt, bq[0], bq[1]);
            mma16816(dwk[h][ks], axt, bk[0], bk[1]);
            if (f_res) mma16816(dwr[h][ks], axt, br[0], br[1]);
          }
        }
      }
    }
    // ---- dX out -------------------------------------------------------------------------------------
    float* dxb = dx + b * (long long)F * KIN;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int n = 0; n < KIN / 8; ++n) {
        const int r0 = 16 * mt + g, c0 = 8 * n + 2 * t;
        if (r0 < F) *reinterpret_cast<float2*>(dxb + r0 * KIN + c0) = make_float2(dxc[mt][n][0], dxc[mt][n][1]);
        if (r0 + 8 < F) *reinterpret_cast<float2*>(dxb + (r0 + 8) * KIN + c0) = make_float2(dxc[mt][n][2], dxc[mt][n][3]);
      }
  }
  // ---- reduce the per-warp accumulators: warps add their fragments into ONE CTA-wide array in warp
  // order (deterministic), which reuses the shared memory of the P / gZ tiles -> partial[blockIdx] ----
  const int wsz = KIN * H * 8;
  const int pf = 3 * wsz + 16;
  static_assert(sizeof(s_pz) >= (3 * KIN * HT * 8 + 16) * sizeof(float), "CTA accumulator must fit in the P/gZ tiles");
  float* acc = reinterpret_cast<float*>(&s_pz[0][0][0]);
  // dgamma / dbeta: lanes with the same t hold the same columns -> sum over g
  float dg[2], db[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float a = dgam[q], c = dbet[q];
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      c += __shfl_xor_sync(0xffffffffu, c, off);
    }
    dg[q] = a;
    db[q] = c;
  }
  for (int w = 0; w < kAtBwdWarps; ++w) {
    __syncthreads();
    if (warp != w) continue;
    auto put = [&](float* dst, float v) { *dst = (w == 0) ? v : *dst + v; };
#pragma unroll
    for (int h = 0; h < HT; ++h) {
      if (h < H) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const int c0 = 16 * ks + g;
          float* a0 = acc + ((c0 * H + h) * 8 + 2 * t);
          float* a1 = acc + (((c0 + 8) * H + h) * 8 + 2 * t);
          put(a0, dwq[h][ks][0]); put(a0 + 1, dwq[h][ks][1]); put(a1, dwq[h][ks][2]); put(a1 + 1, dwq[h][ks][3]);
          put(a0 + wsz, dwk[h][ks][0]); put(a0 + wsz + 1, dwk[h][ks][1]);
          put(a1 + wsz, dwk[h][ks][2]); put(a1 + wsz + 1, dwk[h][ks][3]);
          put(a0 + 2 * wsz, dwr[h][ks][0]); put(a0 + 2 * wsz + 1, dwr[h][ks][1]);
          put(a1 + 2 * wsz, dwr[h][ks][2]); put(a1 + 2 * wsz + 1, dwr[h][ks][3]);
        }
      }
    }
    if (g == 0) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        put(acc + 3 * wsz + 2 * t + q, dg[q]);
        put(acc + 3 * wsz + 8 + 2 * t + q, db[q]);
      }
    }
  }
  __syncthreads();
  float* out = partial + (long long)blockIdx.x * pf;
  for (int i = tid; i < pf; i += kAtBwdThreads) out[i] = acc[i];
}

}  // namespace

bool attn_tc_supported(const AttnDims& p, int DH) {
  return DH == 8 && p.F <= 32 && p.kin % 16 == 0 && p.kin <= 64 && p.H <= 8;
}

int attn_tc_fwd(const float* x, const float* wq, const float* wk, const float* wr, const float* gamma,
                const float* beta, float* y, const AttnDims& p, int sms, cudaStream_t st) {
  const int KS = p.kin / 16;
  const float* wr_ = p.use_res ? wr : nullptr;
  const size_t smem = (size_t)3 * p.H * KS * 64 * 4;
  const int grid = (int)std::max<long long>(1, std::min<long long>((p.B + kAtWarps - 1) / kAtWarps, (long long)sms * 6));
  ProfileScope ps("attn_tc_fwd_kernel", st);
  const bool all_on = p.use_scale && p.use_res && p.use_ln && p.relu;
#define KON_ATF(KS_)                                                                                       \
  if (all_on) attn_tc_fwd_kernel<KS_, true><<<grid, kAtThreads, smem, st>>>(x, wq, wk, wr_, gamma, beta, y, p); \
  else attn_tc_fwd_kernel<KS_, false><<<grid, kAtThreads, smem, st>>>(x, wq, wk, wr_, gamma, beta, y, p)
  switch (KS) {
    case 1: KON_ATF(1); break;
    case 2: KON_ATF(2); break;
    case 3: KON_ATF(3); break;
    default: KON_ATF(4); break;
  }
#undef KON_ATF
  KON_LAUNCH_CHECK("attn_tc_fwd_kernel");
  return KON_OK;
}

bool attn_tc_bwd_supported(const AttnDims& p, int DH) {
  return attn_tc_supported(p, DH) && p.kin <= 32 && p.H <= 4;
}

// partial: [grid][3*wsz + 2*8] floats (same layout as the fp32 path: dWq | dWk | dWr | dgamma | dbeta)
int attn_tc_bwd(const float* x, const float* wq, const float* wk, const float* wr, const float* gamma,
                const float* beta, const float* gy, float* dx, float* partial, int max_grid,
                const AttnDims& p, int sms, int* grid_used, cudaStream_t st) {
  const int KS = p.kin / 16;
  const float* wr_ = p.use_res ? wr : nullptr;
  const int wsz = p.kin * p.H * 8, pf = 3 * wsz + 16;
  const size_t smem = (size_t)3 * p.H * KS * 64 * 4 + (size_t)3 * p.H * (p.kin / 8) * 32 * 4;
  int grid = (int)std::max<long long>(1, std::min<long long>((p.B + kAtBwdWarps - 1) / kAtBwdWarps,
                                                             (long long)sms * KON_ATB_MINB));   // one resident wave
  grid = std::min(grid, max_grid);
  *grid_used = grid;
  ProfileScope ps("attn_tc_bwd_kernel", st);
  const bool all_on = p.use_scale && p.use_res && p.use_ln && p.relu;
#define KON_ATB(KS_, HT_)                                                                                   \
  do {                                                                                                      \
    if (all_on) {                                                                                           \
      KON_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<KS_, HT_, true>,                                     \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
      attn_tc_bwd_kernel<KS_, HT_, true><<<grid, kAtBwdThreads, smem, st>>>(x, wq, wk, wr_, gamma, beta, gy, \
                                                                           dx, partial, p);                 \
    } else {                                                                                                \
      KON_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<KS_, HT_, false>,                                    \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
      attn_tc_bwd_kernel<KS_, HT_, false><<<grid, kAtBwdThreads, smem, st>>>(x, wq, wk, wr_, gamma, beta,    \
                                                                            gy, dx, partial, p);            \
    }                                                                                                       \
  } while (0)
  if (KS == 1 && p.H <= 2) KON_ATB(1, 2);
  else if (KS == 1) KON_ATB(1, 4);
  else if (p.H <= 2) KON_ATB(2, 2);
  else KON_ATB(2, 4);
#undef KON_ATB
  KON_LAUNCH_CHECK("attn_tc_bwd_kernel");
  return KON_OK;
}

}  // namespace kon
