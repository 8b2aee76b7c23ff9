// a9-a10: AutoInt field self-attention block, forward and backward.
//
// Replaces MultHeadAttentionLayer.call (BL:356-377: 4 tensordot + transposes),
// ProductAttentionLayer.call (BL:292-311: QK^T, sigmoid -- not softmax, BL:286 -- and
// S.V with V = X.key_w, BL:360), LayerNormalization (BL:368-369, Keras eps = 1e-3) and the
// Add + ReLU the wrapping DnnLayer applies (CL:205-216).
//
//   Q = X Wq, K = X Wk, R = X Wr           X [B,F,kin], W [kin,H,d]
//   S = sigmoid(Q K^T / sqrt(d))            per (head, sample): [F,F]
//   O = S K
//   Y = ReLU(O * inv + (beta - mean * inv) + R),  inv = rsqrt(var + eps) * gamma   -> [H,B,F,d]
//
// One warp owns one sample; lane i owns field row i (F <= 32).  The op is HBM-bound on
// paper (25 FLOP/B); this first version does the tiny per-sample matmuls on CUDA cores in
// fp32 and keeps every intermediate on chip: the only HBM traffic is X in, Y out (fwd) and
// X, gY in, dX out (bwd; Q/K/S/O and the LayerNorm statistics are recomputed).
// Weight gradients are accumulated per warp in shared memory, reduced per CTA in warp
// order and finished by a second kernel in CTA order (deterministic).
#include "attn_common.cuh"

namespace kon {

constexpr int kAttnFwdThreads = 256;
constexpr int kAttnBwdThreads = 128;

__device__ __forceinline__ float sigmoidf_(float z) { return __frcp_rn(1.f + __expf(-z)); }

__host__ __device__ inline int al4(int n) { return (n + 3) & ~3; }

// DH contiguous floats from a 16-byte aligned shared-memory row (128-bit loads when DH % 4 == 0)
template <int DH>
__device__ __forceinline__ void load_row(const float* __restrict__ p, float* out) {
  if constexpr (DH % 4 == 0) {
#pragma unroll
    for (int e = 0; e < DH; e += 4) {
      const float4 v = *reinterpret_cast<const float4*>(p + e);
      out[e] = v.x; out[e + 1] = v.y; out[e + 2] = v.z; out[e + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < DH; ++e) out[e] = p[e];
  }
}

// Projection of this lane's row onto head h: q/k/r[e] = sum_c x[c] * W[c][h][e].
template <int DH>
__device__ __forceinline__ void project(const float* __restrict__ xrow, int kin, int H, int h,
                                        const float* __restrict__ Wq, const float* __restrict__ Wk,
                                        const float* __restrict__ Wr, bool use_res, float* q,
                                        float* k, float* r) {
#pragma unroll
  for (int e = 0; e < DH; ++e) q[e] = k[e] = r[e] = 0.f;
  for (int c = 0; c < kin; ++c) {
    const float xv = xrow[c];
    const float* wq = Wq + (c * H + h) * DH;
    const float* wk = Wk + (c * H + h) * DH;
    const float* wr = Wr + (c * H + h) * DH;
#pragma unroll
    for (int e = 0; e < DH; ++e) {
      q[e] = fmaf(xv, wq[e], q[e]);
      k[e] = fmaf(xv, wk[e], k[e]);
    }
    if (use_res) {
#pragma unroll
      for (int e = 0; e < DH; ++e) r[e] = fmaf(xv, wr[e], r[e]);
    }
  }
}

// o[e] = sum_j sigmoid(q . k_j [/ sqrt(d)]) * k_j[e] over the F rows staged in ks.
template <int DH>
__device__ __forceinline__ void attend(const float* q, const float* __restrict__ ks, int F,
                                       bool use_scale, float sqrt_d, float* o) {
#pragma unroll
  for (int e = 0; e < DH; ++e) o[e] = 0.f;
  const float inv_sd = use_scale ? 1.f / sqrt_d : 1.f;
  for (int j = 0; j < F; ++j) {
    float kj[DH];
    load_row<DH>(ks + j * DH, kj);
    float z = 0.f;
#pragma unroll
    for (int e = 0; e < DH; ++e) z = fmaf(q[e], kj[e], z);
    z *= inv_sd;
    const float sg = sigmoidf_(z);
#pragma unroll
    for (int e = 0; e < DH; ++e) o[e] = fmaf(sg, kj[e], o[e]);
  }
}

template <int DH>
__global__ void __launch_bounds__(kAttnFwdThreads)
attn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wq,
                const float* __restrict__ wk, const float* __restrict__ wr,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ y, const AttnDims p) {
  extern __shared__ __align__(16) float smem[];
  const int kin = p.kin, H = p.H, F = p.F;
  const int wsz = kin * H * DH;
  float* Wq = smem;
  float* Wk = Wq + wsz;
  float* Wr = Wk + wsz;
  float* gam = Wr + wsz;
  float* bet = gam + DH;
  const int xstride = kin + 1;
  constexpr int kWarps = kAttnFwdThreads / 32;
  float* warp_base = smem + al4(3 * wsz + 2 * DH);
  const int per_warp = al4(F * xstride) + al4(F * DH);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = warp_base + wid * per_warp;
  float* ks = xs + al4(F * xstride);
  for (int i = threadIdx.x; i < wsz; i += kAttnFwdThreads) {
    Wq[i] = wq[i];
    Wk[i] = wk[i];
    Wr[i] = p.use_res ? wr[i] : 0.f;
  }
  if (threadIdx.x < DH) {
    gam[threadIdx.x] = p.use_ln ? gamma[threadIdx.x] : 1.f;
    bet[threadIdx.x] = p.use_ln ? beta[threadIdx.x] : 0.f;
  }
  __syncthreads();
  const float sqrt_d = sqrtf((float)DH);
  const bool row_on = lane < F;
  for (long long b = (long long)blockIdx.x * kWarps + wid; b < p.B;
       b += (long long)gridDim.x * kWarps) {
    const float* xb = x + b * (long long)F * kin;
    for (int i = lane; i < F * kin; i += 32) xs[(i / kin) * xstride + (i % kin)] = __ldg(xb + i);
    __syncwarp();
    for (int h = 0; h < H; ++h) {
      float q[DH], k[DH], r[DH], o[DH];
      if (row_on) {
        project<DH>(xs + lane * xstride, kin, H, h, Wq, Wk, Wr, p.use_res, q, k, r);
#pragma unroll
        for (int e = 0; e < DH; ++e) ks[lane * DH + e] = k[e];
      }
      __syncwarp();
      if (row_on) {
        attend<DH>(q, ks, F, p.use_scale, sqrt_d, o);
        if (p.use_ln) {
          float mean = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) mean += o[e];
          mean *= (1.f / DH);
          float var = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) var = fmaf(o[e] - mean, o[e] - mean, var);
          var *= (1.f / DH);
          const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll
          for (int e = 0; e < DH; ++e) {
            const float inv = rstd * gam[e];
            o[e] = o[e] * inv + (bet[e] - mean * inv);
          }
        }
        float* yp = y + (((long long)h * p.B + b) * F + lane) * DH;
#pragma unroll
        for (int e = 0; e < DH; ++e) {
          float v = p.use_res ? r[e] + o[e] : o[e];   // Add([res, atten_v]) (CL:207,212)
          if (p.relu) v = fmaxf(v, 0.f);
          yp[e] = v;
        }
      }
      __syncwarp();
    }
  }
}

// Partial layout per CTA: dWq | dWk | dWr (each kin*H*DH) | dgamma[DH] | dbeta[DH]
__host__ __device__ inline int attn_partial_floats(int kin, int H, int DH) {
  return 3 * kin * H * DH + 2 * DH;
}

template <int DH>
__global__ void __launch_bounds__(kAttnBwdThreads)
attn_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wq,
                const float* __restrict__ wk, const float* __restrict__ wr,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ gy, float* __restrict__ dx, float* __restrict__ partial,
                const AttnDims p) {
  extern __shared__ __align__(16) float smem[];
  const int kin = p.kin, H = p.H, F = p.F;
  const int wsz = kin * H * DH;
  constexpr int kWarps = kAttnBwdThreads / 32;
  float* Wq = smem;
  float* Wk = Wq + wsz;
  float* Wr = Wk + wsz;
  float* gam = Wr + wsz;
  float* bet = gam + DH;
  float* warp_base = smem + al4(3 * wsz + 2 * DH);
  const int xstride = kin + 1;
  // per warp (every array 16-byte aligned): xs, dxs [F][xstride]; qs, ks, gos, gqs, gks, grs [F][DH];
  // dW [3][wsz] directly followed by dgb [2*DH]
  const int axs = al4(F * xstride), afd = al4(F * DH);
  const int per_warp = 2 * axs + 6 * afd + al4(3 * wsz + 2 * DH);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* xs = warp_base + wid * per_warp;
  float* dxs = xs + axs;
  float* qs = dxs + axs;
  float* ks = qs + afd;
  float* gos = ks + afd;
  float* gqs = gos + afd;
  float* gks = gqs + afd;
  float* grs = gks + afd;
  float* dWs = grs + afd;
  float* dgb = dWs + 3 * wsz;
  for (int i = threadIdx.x; i < wsz; i += kAttnBwdThreads) {
    Wq[i] = wq[i];
    Wk[i] = wk[i];
    Wr[i] = p.use_res ? wr[i] : 0.f;
  }
  if (threadIdx.x < DH) {
    gam[threadIdx.x] = p.use_ln ? gamma[threadIdx.x] : 1.f;
    bet[threadIdx.x] = p.use_ln ? beta[threadIdx.x] : 0.f;
  }
  for (int i = lane; i < 3 * wsz; i += 32) dWs[i] = 0.f;
  float dgam[DH], dbet[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) dgam[e] = dbet[e] = 0.f;
  __syncthreads();
  const float sqrt_d = sqrtf((float)DH);
  const bool row_on = lane < F;

  for (long long b = (long long)blockIdx.x * kWarps + wid; b < p.B;
       b += (long long)gridDim.x * kWarps) {
    const float* xb = x + b * (long long)F * kin;
    for (int i = lane; i < F * kin; i += 32) {
      xs[(i / kin) * xstride + (i % kin)] = __ldg(xb + i);
      dxs[(i / kin) * xstride + (i % kin)] = 0.f;
    }
    __syncwarp();
    for (int h = 0; h < H; ++h) {
      float q[DH], k[DH], r[DH], o[DH], gO[DH], gP[DH];
#pragma unroll
      for (int e = 0; e < DH; ++e) q[e] = k[e] = r[e] = o[e] = gO[e] = gP[e] = 0.f;
      if (row_on) {
        project<DH>(xs + lane * xstride, kin, H, h, Wq, Wk, Wr, p.use_res, q, k, r);
#pragma unroll
        for (int e = 0; e < DH; ++e) {
          ks[lane * DH + e] = k[e];
          qs[lane * DH + e] = q[e];
        }
      }
      __syncwarp();
      if (row_on) {
        attend<DH>(q, ks, F, p.use_scale, sqrt_d, o);
        // forward tail again: LayerNorm statistics and the ReLU mask
        float mean = 0.f, rstd = 1.f, xhat[DH];
        if (p.use_ln) {
#pragma unroll
          for (int e = 0; e < DH; ++e) mean += o[e];
          mean *= (1.f / DH);
          float var = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) var = fmaf(o[e] - mean, o[e] - mean, var);
          var *= (1.f / DH);
          rstd = rsqrtf(var + p.ln_eps);
        }
        const float* gyp = gy + (((long long)h * p.B + b) * F + lane) * DH;
#pragma unroll
        for (int e = 0; e < DH; ++e) {
          xhat[e] = (o[e] - mean) * rstd;
          float pre;
          if (p.use_ln) {
            const float inv = rstd * gam[e];
            pre = o[e] * inv + (bet[e] - mean * inv);
          } else {
            pre = o[e];
          }
          if (p.use_res) pre = r[e] + pre;
          const float gv = __ldg(gyp + e);
          gP[e] = (p.relu && !(pre > 0.f)) ? 0.f : gv;
        }
        if (p.use_ln) {
          float m1 = 0.f, m2 = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) {
            dgam[e] = fmaf(gP[e], xhat[e], dgam[e]);
            dbet[e] += gP[e];
            const float gx = gP[e] * gam[e];
            m1 += gx;
            m2 = fmaf(gx, xhat[e], m2);
          }
          m1 *= (1.f / DH);
          m2 *= (1.f / DH);
#pragma unroll
          for (int e = 0; e < DH; ++e) gO[e] = rstd * (gP[e] * gam[e] - m1 - xhat[e] * m2);
        } else {
#pragma unroll
          for (int e = 0; e < DH; ++e) gO[e] = gP[e];
        }
#pragma unroll
        for (int e = 0; e < DH; ++e) {
          gos[lane * DH + e] = gO[e];
          grs[lane * DH + e] = p.use_res ? gP[e] : 0.f;
        }
      }
      __syncwarp();
      if (row_on) {
        // row pass (lane = i): gq_i = sum_j gZ[i][j] k_j
        // column pass (lane = j): gk_j = sum_i gZ[i][j] q_i + S[i][j] gO_i
        float gq[DH], gk[DH];
#pragma unroll
        for (int e = 0; e < DH; ++e) gq[e] = gk[e] = 0.f;
        const float inv_scale = p.use_scale ? 1.f / sqrt_d : 1.f;
        for (int j = 0; j < F; ++j) {
          float kj[DH];
          load_row<DH>(ks + j * DH, kj);
          float z = 0.f, gs = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) {
            z = fmaf(q[e], kj[e], z);
            gs = fmaf(gO[e], kj[e], gs);
          }
          z *= inv_scale;
          const float sg = sigmoidf_(z);
          const float gz = gs * sg * (1.f - sg) * inv_scale;
#pragma unroll
          for (int e = 0; e < DH; ++e) gq[e] = fmaf(gz, kj[e], gq[e]);
        }
        for (int i = 0; i < F; ++i) {
          float qi[DH], goi[DH];
          load_row<DH>(qs + i * DH, qi);
          load_row<DH>(gos + i * DH, goi);
          float z = 0.f, gs = 0.f;
#pragma unroll
          for (int e = 0; e < DH; ++e) {
            z = fmaf(qi[e], k[e], z);
            gs = fmaf(goi[e], k[e], gs);
          }
          z *= inv_scale;
          const float sg = sigmoidf_(z);
          const float gz = gs * sg * (1.f - sg) * inv_scale;
#pragma unroll
          for (int e = 0; e < DH; ++e) gk[e] = fmaf(gz, qi[e], fmaf(sg, goi[e], gk[e]));
        }
#pragma unroll
        for (int e = 0; e < DH; ++e) {
          gqs[lane * DH + e] = gq[e];
          gks[lane * DH + e] = gk[e];
        }
        // dX row: sum over heads, accumulated in shared memory
        float* dxr = dxs + lane * xstride;
        for (int c = 0; c < kin; ++c) {
          const float* a = Wq + (c * H + h) * DH;
          const float* bb = Wk + (c * H + h) * DH;
          const float* cc = Wr + (c * H + h) * DH;
          float acc = dxr[c];
#pragma unroll
          for (int e = 0; e < DH; ++e) {
            acc = fmaf(gq[e], a[e], acc);
            acc = fmaf(gk[e], bb[e], acc);
            acc = fmaf(gP[e], cc[e], acc);   // Wr staged as zeros when !use_res
          }
          dxr[c] = acc;
        }
      }
      __syncwarp();
      // weight gradients: lane owns (c,e) pairs pr = lane + 32 t of head h
      for (int pr = lane; pr < kin * DH; pr += 32) {
        const int c = pr / DH, e = pr % DH;
        float aq = 0.f, ak = 0.f, ar = 0.f;
        for (int i = 0; i < F; ++i) {
          const float xv = xs[i * xstride + c];
          aq = fmaf(xv, gqs[i * DH + e], aq);
          ak = fmaf(xv, gks[i * DH + e], ak);
          ar = fmaf(xv, grs[i * DH + e], ar);
        }
        const int o_ = (c * H + h) * DH + e;
        dWs[o_] += aq;
        dWs[wsz + o_] += ak;
        dWs[2 * wsz + o_] += ar;
      }
      __syncwarp();
    }
    float* dxb = dx + b * (long long)F * kin;
    for (int i = lane; i < F * kin; i += 32) dxb[i] = dxs[(i / kin) * xstride + (i % kin)];
    __syncwarp();
  }
  // dgamma/dbeta: lanes -> warp
#pragma unroll
  for (int e = 0; e < DH; ++e) {
    const float a = warp_sum(dgam[e]);
    const float c = warp_sum(dbet[e]);
    if (lane == 0) {
      dgb[e] = a;
      dgb[DH + e] = c;
    }
  }
  __syncthreads();
  // warps -> CTA partial (warp order)
  const int pf = attn_partial_floats(kin, H, DH);
  float* out = partial + (long long)blockIdx.x * pf;
  for (int i = threadIdx.x; i < pf; i += kAttnBwdThreads) {
    float acc = 0.f;
    for (int w = 0; w < kWarps; ++w) {
      const float* base = warp_base + w * per_warp + 2 * axs + 6 * afd;
      acc += base[i];   // dWs (3*wsz) is directly followed by dgb (2*DH)
    }
    out[i] = acc;
  }
}

// Fixed-order sum of the per-CTA partials: a CTA takes 32 outputs x 8 slices of the partial list (each
// thread adds its slice front to back, the 8 slice sums are added in slice order) -> deterministic, and
// pf/32 CTAs instead of the former 4 CTAs each walking all ~600 partials serially (27 us -> a few us).
__global__ void __launch_bounds__(256)
attn_bwd_finalize_kernel(const float* __restrict__ partial, int n_part, int pf, int wsz, int DH,
                         float* __restrict__ dwq, float* __restrict__ dwk, float* __restrict__ dwr,
                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float sm[8][32];
  const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  const int per = (n_part + 7) / 8;
  const int lo = slice * per, hi = min(n_part, lo + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (i < pf) {
    int p = lo;
    for (; p + 3 < hi; p += 4) {
      a0 += partial[(long long)p * pf + i];
      a1 += partial[(long long)(p + 1) * pf + i];
      a2 += partial[(long long)(p + 2) * pf + i];
      a3 += partial[(long long)(p + 3) * pf + i];
    }
    for (; p < hi; ++p) a0 += partial[(long long)p * pf + i];
  }
  sm[slice][col] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (slice == 0 && i < pf) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += sm[k][col];
    if (i < wsz) dwq[i] = acc;
    else if (i < 2 * wsz) dwk[i - wsz] = acc;
    else if (i < 3 * wsz) { if (dwr) dwr[i - 2 * wsz] = acc; }
    else if (i < 3 * wsz + DH) { if (dgamma) dgamma[i - 3 * wsz] = acc; }
    else { if (dbeta) dbeta[i - 3 * wsz - DH] = acc; }
  }
}

static int attn_bwd_grid(long long B, int sms) {
  constexpr int kWarps = kAttnBwdThreads / 32;
  // also the size of the partial-sum workspace: the tensor-core backward (4 warps per CTA) runs up to
  // 6 resident CTAs per SM
  return (int)std::max<long long>(1, std::min<long long>((B + kWarps - 1) / kWarps, (long long)sms * 6));
}

static size_t attn_fwd_smem(int F, int kin, int H, int DH) {
  return sizeof(float) * ((size_t)al4(3 * kin * H * DH + 2 * DH) +
                          (size_t)(kAttnFwdThreads / 32) * (al4(F * (kin + 1)) + al4(F * DH)));
}
static size_t attn_bwd_smem(int F, int kin, int H, int DH) {
  const size_t per_warp = 2 * (size_t)al4(F * (kin + 1)) + 6 * (size_t)al4(F * DH) + al4(3 * kin * H * DH + 2 * DH);
  return sizeof(float) * ((size_t)al4(3 * kin * H * DH + 2 * DH) + (kAttnBwdThreads / 32) * per_warp);
}

}  // namespace kon

using namespace kon;

static int attn_check(const DLTensor* x, const DLTensor* wq, const DLTensor* wk, const DLTensor* wr,
                      const DLTensor* gamma, const DLTensor* beta, int32_t flags, AttnDims* p,
                      int* DH) {
  KON_TRY(check_cuda_tensor(x, "x"));
  const int dev = x->device.device_id;
  KON_TRY(check_cuda_tensor(wq, "wq", dev));
  KON_TRY(check_cuda_tensor(wk, "wk", dev));
  KON_REQUIRE(is_f32(x) && x->ndim == 3 && is_compact(x), KON_EINVAL,
              "x must be compact float32 [B,F,kin]");
  p->ysh = p->ysb = p->ysf = p->gsh = p->gsb = p->gsf = 0;
  p->B = x->shape[0];
  p->F = (int)x->shape[1];
  p->kin = (int)x->shape[2];
  KON_REQUIRE(is_f32(wq) && wq->ndim == 3 && wq->shape[0] == p->kin && is_compact(wq), KON_EINVAL,
              "wq must be compact float32 [kin,H,d]");
  p->H = (int)wq->shape[1];
  *DH = (int)wq->shape[2];
  KON_REQUIRE(is_f32(wk) && numel(wk) == numel(wq) && is_compact(wk), KON_EINVAL,
              "wk must match wq");
  p->use_scale = (flags & KON_ATTN_USE_SCALE) ? 1 : 0;
  p->use_ln = (flags & KON_ATTN_USE_LN) ? 1 : 0;
  p->use_res = (flags & KON_ATTN_USE_RES) ? 1 : 0;
  p->relu = (flags & KON_ATTN_RELU) ? 1 : 0;
  if (p->use_res) {
    KON_TRY(check_cuda_tensor(wr, "wr", dev));
    KON_REQUIRE(is_f32(wr) && numel(wr) == numel(wq) && is_compact(wr), KON_EINVAL,
                "wr must match wq");
  }
  if (p->use_ln) {
    KON_TRY(check_cuda_tensor(gamma, "gamma", dev));
    KON_TRY(check_cuda_tensor(beta, "beta", dev));
    KON_REQUIRE(is_f32(gamma) && numel(gamma) == *DH && is_f32(beta) && numel(beta) == *DH,
                KON_EINVAL, "gamma and beta must be float32 [d]");
  }
  KON_REQUIRE(p->F >= 1 && p->F <= 32, KON_EUNSUPPORTED, "fields F=%d outside [1,32]", p->F);
  KON_REQUIRE(p->kin >= 1 && p->kin <= 128, KON_EUNSUPPORTED, "kin=%d outside [1,128]", p->kin);
  KON_REQUIRE(p->H >= 1 && p->H <= 16, KON_EUNSUPPORTED, "heads H=%d outside [1,16]", p->H);
  KON_REQUIRE(*DH == 4 || *DH == 8 || *DH == 16 || *DH == 32, KON_EUNSUPPORTED,
              "attention dim d=%d not in {4,8,16,32}", *DH);
  return KON_OK;
}

#define KON_ATTN_DISPATCH(DH, CALL) \
  switch (DH) {                     \
    case 4: CALL(4); break;         \
    case 8: CALL(8); break;         \
    case 16: CALL(16); break;       \
    default: CALL(32); break;       \
  }

extern "C" int kon_attn_fwd(const DLTensor* x, const DLTensor* wq, const DLTensor* wk,
                            const DLTensor* wr, const DLTensor* gamma, const DLTensor* beta,
                            DLTensor* y, float ln_eps, int32_t flags, void* stream) {
  AttnDims p;
  int DH;
  KON_TRY(attn_check(x, wq, wk, wr, gamma, beta, flags, &p, &DH));
  p.ln_eps = ln_eps;
  const int dev = x->device.device_id;
  KON_TRY(check_cuda_tensor(y, "y", dev));
  KON_REQUIRE(is_f32(y) && y->ndim == 4 && y->shape[0] == p.H && y->shape[1] == p.B &&
                  y->shape[2] == p.F && y->shape[3] == DH && (DH == 1 || stride_of(y, 3) == 1),
              KON_EINVAL, "y must be float32 [H,B,F,d] with a compact last dim");
  KON_REQUIRE(is_compact(y) || ((flags & KON_ATTN_BF16) && stride_of(y, 0) % 2 == 0 && stride_of(y, 1) % 2 == 0 &&
                                stride_of(y, 2) % 2 == 0 && ((uintptr_t)data_ptr<float>(y) & 7u) == 0),
              KON_EINVAL, "a strided y (permuted window) is only taken by the KON_ATTN_BF16 path, 8-B aligned rows");
  p.ysh = stride_of(y, 0);
  p.ysb = stride_of(y, 1);
  p.ysf = stride_of(y, 2);
  if (p.B == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (flags & KON_ATTN_BF16) {
    KON_REQUIRE(attn_tc_supported(p, DH), KON_EUNSUPPORTED,
                "KON_ATTN_BF16 needs F <= 32, kin in {16,32,48,64}, d == 8, H <= 8 (got F=%d kin=%d d=%d H=%d)",
                p.F, p.kin, DH, p.H);
    return attn_tc_fwd(data_ptr<float>(x), data_ptr<float>(wq), data_ptr<float>(wk),
                       p.use_res ? data_ptr<float>(wr) : nullptr, p.use_ln ? data_ptr<float>(gamma) : nullptr,
                       p.use_ln ? data_ptr<float>(beta) : nullptr, data_ptr<float>(y), p, sm_count_of(dev), st);
  }
  const size_t smem = attn_fwd_smem(p.F, p.kin, p.H, DH);
  KON_REQUIRE(smem <= 227 * 1024, KON_EUNSUPPORTED, "attention shape needs %zu B of shared memory",
              smem);
  constexpr int kWarps = kAttnFwdThreads / 32;
  const int grid = (int)std::min<long long>((p.B + kWarps - 1) / kWarps,
                                            (long long)sm_count_of(dev) * 4);
#define CALL(N)                                                                                  \
  KON_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem));                                                     \
  attn_fwd_kernel<N><<<grid, kAttnFwdThreads, smem, st>>>(                                       \
      data_ptr<float>(x), data_ptr<float>(wq), data_ptr<float>(wk),                              \
      p.use_res ? data_ptr<float>(wr) : nullptr, p.use_ln ? data_ptr<float>(gamma) : nullptr,    \
      p.use_ln ? data_ptr<float>(beta) : nullptr, data_ptr<float>(y), p)
  KON_ATTN_DISPATCH(DH, CALL)
#undef CALL
  KON_LAUNCH_CHECK("attn_fwd_kernel");
  return KON_OK;
}

extern "C" size_t kon_attn_bwd_workspace_bytes(int64_t batch, int32_t fields, int32_t kin,
                                               int32_t heads, int32_t d, int device_id) {
  (void)fields;
  return (size_t)attn_bwd_grid(batch, sm_count_of(device_id)) *
         attn_partial_floats(kin, heads, d) * sizeof(float);
}

extern "C" int kon_attn_bwd(const DLTensor* x, const DLTensor* wq, const DLTensor* wk,
                            const DLTensor* wr, const DLTensor* gamma, const DLTensor* beta,
                            const DLTensor* gy, DLTensor* dx, DLTensor* dwq, DLTensor* dwk,
                            DLTensor* dwr, DLTensor* dgamma, DLTensor* dbeta, DLTensor* workspace,
                            float ln_eps, int32_t flags, void* stream) {
  AttnDims p;
  int DH;
  KON_TRY(attn_check(x, wq, wk, wr, gamma, beta, flags, &p, &DH));
  p.ln_eps = ln_eps;
  const int dev = x->device.device_id;
  KON_TRY(check_cuda_tensor(gy, "gy", dev));
  KON_TRY(check_cuda_tensor(dx, "dx", dev));
  KON_TRY(check_cuda_tensor(dwq, "dwq", dev));
  KON_TRY(check_cuda_tensor(dwk, "dwk", dev));
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  KON_REQUIRE(is_f32(gy) && gy->ndim == 4 && gy->shape[0] == p.H && gy->shape[1] == p.B &&
                  gy->shape[2] == p.F && gy->shape[3] == DH && (DH == 1 || stride_of(gy, 3) == 1),
              KON_EINVAL, "gy must be float32 [H,B,F,d] with a compact last dim");
  KON_REQUIRE(is_compact(gy) || ((flags & KON_ATTN_BF16) && attn_tc_bwd_supported(p, DH) &&
                                 stride_of(gy, 0) % 2 == 0 && stride_of(gy, 1) % 2 == 0 && stride_of(gy, 2) % 2 == 0 &&
                                 ((uintptr_t)data_ptr<float>(gy) & 7u) == 0),
              KON_EINVAL, "a strided gy (permuted window) is only taken by the KON_ATTN_BF16 backward, 8-B aligned rows");
  p.gsh = stride_of(gy, 0);
  p.gsb = stride_of(gy, 1);
  p.gsf = stride_of(gy, 2);
  KON_REQUIRE(is_f32(dx) && numel(dx) == numel(x) && is_compact(dx), KON_EINVAL,
              "dx must be compact float32 like x");
  KON_REQUIRE(is_f32(dwq) && numel(dwq) == numel(wq) && is_compact(dwq) && is_f32(dwk) &&
                  numel(dwk) == numel(wq) && is_compact(dwk),
              KON_EINVAL, "dwq/dwk must be compact float32 like wq");
  float *dwr_p = nullptr, *dg_p = nullptr, *db_p = nullptr;
  if (p.use_res) {
    KON_TRY(check_cuda_tensor(dwr, "dwr", dev));
    KON_REQUIRE(is_f32(dwr) && numel(dwr) == numel(wq) && is_compact(dwr), KON_EINVAL,
                "dwr must be compact float32 like wq");
    dwr_p = data_ptr<float>(dwr);
  }
  if (p.use_ln) {
    KON_TRY(check_cuda_tensor(dgamma, "dgamma", dev));
    KON_TRY(check_cuda_tensor(dbeta, "dbeta", dev));
    KON_REQUIRE(is_f32(dgamma) && numel(dgamma) == DH && is_f32(dbeta) && numel(dbeta) == DH,
                KON_EINVAL, "dgamma/dbeta must be float32 [d]");
    dg_p = data_ptr<float>(dgamma);
    db_p = data_ptr<float>(dbeta);
  }
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = attn_bwd_grid(p.B, sm_count_of(dev));
  const int pf = attn_partial_floats(p.kin, p.H, DH);
  const size_t need = (size_t)grid * pf * sizeof(float);
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= need, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), need);
  const size_t smem = attn_bwd_smem(p.F, p.kin, p.H, DH);
  KON_REQUIRE(smem <= 227 * 1024, KON_EUNSUPPORTED, "attention shape needs %zu B of shared memory",
              smem);
  float* partial = data_ptr<float>(workspace);
  const int grid32 = std::min(grid, sm_count_of(dev) * 4);      // fp32 CUDA-core path: 4 CTAs per SM
  if ((flags & KON_ATTN_BF16) && attn_tc_bwd_supported(p, DH)) {
    int used = grid;
    KON_TRY(attn_tc_bwd(data_ptr<float>(x), data_ptr<float>(wq), data_ptr<float>(wk),
                        p.use_res ? data_ptr<float>(wr) : nullptr, p.use_ln ? data_ptr<float>(gamma) : nullptr,
                        p.use_ln ? data_ptr<float>(beta) : nullptr, data_ptr<float>(gy), data_ptr<float>(dx),
                        partial, grid, p, sm_count_of(dev), &used, st));
    attn_bwd_finalize_kernel<<<(pf + 31) / 32, 256, 0, st>>>(partial, used, pf, p.kin * p.H * DH, DH,
                                                              data_ptr<float>(dwq), data_ptr<float>(dwk), dwr_p,
                                                              dg_p, db_p);
    KON_LAUNCH_CHECK("attn_bwd_finalize_kernel");
    return KON_OK;
  }
#define CALL(N)                                                                                  \
  KON_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem));                                                     \
  attn_bwd_kernel<N><<<grid32, kAttnBwdThreads, smem, st>>>(                                     \
      data_ptr<float>(x), data_ptr<float>(wq), data_ptr<float>(wk),                              \
      p.use_res ? data_ptr<float>(wr) : nullptr, p.use_ln ? data_ptr<float>(gamma) : nullptr,    \
      p.use_ln ? data_ptr<float>(beta) : nullptr, data_ptr<float>(gy), data_ptr<float>(dx),      \
      partial, p)
  KON_ATTN_DISPATCH(DH, CALL)
#undef CALL
  KON_LAUNCH_CHECK("attn_bwd_kernel");
  attn_bwd_finalize_kernel<<<(pf + 31) / 32, 256, 0, st>>>(
      partial, grid32, pf, p.kin * p.H * DH, DH, data_ptr<float>(dwq), data_ptr<float>(dwk), dwr_p,
      dg_p, db_p);
  KON_LAUNCH_CHECK("attn_bwd_finalize_kernel");
  return KON_OK;
}
