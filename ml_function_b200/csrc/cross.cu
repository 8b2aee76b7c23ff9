// a7: DCN-v1 cross network, forward and backward.
//
// Replaces CrossLayer.call (IL:275-282): per layer Transpose + MatMul ([B,1,D]x[D,1]) +
// BatchMatMul ([B,D,1]x[B,1,1]) + 2 AddV2, i.e. 5*L passes over [B,D].
//
// HBM-bound fp32 work; everything rests on one identity: x_l is never needed as a vector,
//     x_l = c_l x0 + beta_l,   c_0 = 1, c_{l+1} = c_l + s_l,   beta_l = sum_{j<l} b_j  (batch-independent)
// so with the L INDEPENDENT dot products p_l = x0 . w_l of a sample and the per-layer constants
// q_l = beta_l . w_l the whole recurrence is scalar:  s_l = x_l . w_l = c_l p_l + q_l.
//
//   fwd : one pass: read x0, L dots against w (shared memory, 128-bit reads), scalar recurrence,
//         write x_L = c_L x0 + beta_L and the L scalars s_l.         2*D*4 + 4L bytes / sample
//   bwd : with G_l = dL/dx_{l+1} = g + sum_{j>l} t_j w_j and t_l = x0 . G_l:
//             t_l = a + sum_{j>l} t_j p_j,  a = x0 . g        (scalar recurrence again)
//             dx0 = c_L g + sum_l u_l w_l,  u_l = c_l t_l
//             dw_l[d] = sum_b x0[b,d] u_l[b] + beta_l[d] T_l,   T_l = sum_b t_l[b]
//             db_l[d] = sum_b g[b,d] + sum_{j>l} w_j[d] T_j
//         kernel 1 (per sample): x0, g -> dx0, u; column sums of g and T_l in registers.  3*D*4 B / sample
//         kernel 2: the skinny GEMM  M = X0^T U  ([D x B] x [B x L]) per CTA slab, in registers
//         kernel 3/4: fixed-order reduction of the per-CTA partials and the dw / db formulas
//         (deterministic: every sum has a fixed association order).
// A lane owns 4 consecutive columns per 128-column group, so rows move as 128-bit accesses when the
// row stride allows it (the model's concat buffer: W % 4 == 0); any other layout takes scalar accesses.
#include "common.cuh"

namespace kon {

constexpr int kCrossThreads = 256;
constexpr int kCrossWarps = kCrossThreads / 32;
constexpr int kCrossMaxL = 8;
constexpr int kCrossSlab = 32;      // samples per staging step of the dw kernel

template <int NJ>
__device__ __forceinline__ void cross_load_row(const float* __restrict__ p, bool vec, int D, int lane,
                                               float4 (&x)[NJ]) {
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int d0 = 128 * j + 4 * lane;
    if (vec && d0 + 3 < D) {
      x[j] = __ldg(reinterpret_cast<const float4*>(p + d0));
    } else {
      x[j].x = d0 < D ? __ldg(p + d0) : 0.f;
      x[j].y = d0 + 1 < D ? __ldg(p + d0 + 1) : 0.f;
      x[j].z = d0 + 2 < D ? __ldg(p + d0 + 2) : 0.f;
      x[j].w = d0 + 3 < D ? __ldg(p + d0 + 3) : 0.f;
    }
  }
}

__device__ __forceinline__ void cross_store4(float* __restrict__ p, bool vec, int D, int d0, float4 v) {
  if (vec && d0 + 3 < D) {
    *reinterpret_cast<float4*>(p + d0) = v;
  } else {
    if (d0 < D) p[d0] = v.x;
    if (d0 + 1 < D) p[d0 + 1] = v.y;
    if (d0 + 2 < D) p[d0 + 2] = v.z;
    if (d0 + 3 < D) p[d0 + 3] = v.w;
  }
}

// 4 columns of a row this kernel also writes (no read-only path)
__device__ __forceinline__ float4 cross_load4_rw(const float* p, bool vec, int D, int d0) {
  if (vec && d0 + 3 < D) return *reinterpret_cast<const float4*>(p + d0);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d0 < D) v.x = p[d0];
  if (d0 + 1 < D) v.y = p[d0 + 1];
  if (d0 + 2 < D) v.z = p[d0 + 2];
  if (d0 + 3 < D) v.w = p[d0 + 3];
  return v;
}

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// w -> shared [L][Dp] (zero padded), beta_L -> bl_s[Dp], q_l = beta_l . w_l -> q_s[kCrossMaxL].
// Fixed association order (column-strided per thread, shuffle tree, warps in order): every CTA gets
// bit-identical constants.
__device__ void cross_stage_constants(const float* __restrict__ w, const float* __restrict__ b, int D, int Dp,
                                      int L, float* w_s, float* bl_s, float* q_s, float* red_s) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < L * Dp; i += kCrossThreads) {
    const int l = i / Dp, d = i - l * Dp;
    w_s[i] = d < D ? w[l * D + d] : 0.f;
  }
  float qp[kCrossMaxL];
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) qp[l] = 0.f;
  for (int d = tid; d < Dp; d += kCrossThreads) {
    float beta = 0.f;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      if (l < L && d < D) {
        qp[l] = fmaf(beta, w[l * D + d], qp[l]);
        if (b) beta += b[l * D + d];
      }
    }
    if (bl_s) bl_s[d] = beta;
  }
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) {
    const float v = warp_sum(qp[l]);
    if (lane == 0) red_s[wid * kCrossMaxL + l] = v;
  }
  __syncthreads();
  if (tid < kCrossMaxL) {
    float v = 0.f;
    for (int ws = 0; ws < kCrossWarps; ++ws) v += red_s[ws * kCrossMaxL + tid];
    q_s[tid] = v;
  }
  __syncthreads();
}

template <int NJ>
__global__ void __launch_bounds__(kCrossThreads, 2)
cross_fwd_kernel(const float* __restrict__ x0, long long xs, const float* __restrict__ w,
                 const float* __restrict__ b, float* __restrict__ out, float* __restrict__ s,
                 long long B, int D, int L, int vec_in, int vec_out) {
  constexpr int Dp = NJ * 128;
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                    // [L][Dp]
  float* bl_s = w_s + L * Dp;           // [Dp]  beta_L
  float* q_s = bl_s + Dp;               // [kCrossMaxL]
  float* red_s = q_s + kCrossMaxL;      // [kCrossWarps][kCrossMaxL]
  cross_stage_constants(w, b, D, Dp, L, w_s, bl_s, q_s, red_s);

  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kCrossWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kCrossWarps;
  // two samples per warp iteration: every 128-bit read of w serves both
  for (long long r = 2 * warp0; r < B; r += 2 * nwarps) {
    const bool two = r + 1 < B;
    float4 xa[NJ], xb[NJ];
    cross_load_row<NJ>(x0 + r * xs, vec_in != 0, D, lane, xa);
    cross_load_row<NJ>(x0 + (two ? r + 1 : r) * xs, vec_in != 0, D, lane, xb);
    float pa[kCrossMaxL], pb[kCrossMaxL];
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      pa[l] = pb[l] = 0.f;
      if (l < L) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float4 wv = *reinterpret_cast<const float4*>(w_s + l * Dp + 128 * j + 4 * lane);
          pa[l] = dot4(xa[j], wv, pa[l]);
          pb[l] = dot4(xb[j], wv, pb[l]);
        }
        pa[l] = warp_sum(pa[l]);
        pb[l] = warp_sum(pb[l]);
      }
    }
    float ca = 1.f, cb = 1.f;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      if (l < L) {
        const float sa = fmaf(ca, pa[l], q_s[l]), sb = fmaf(cb, pb[l], q_s[l]);
        if (lane == 0) {
          s[r * L + l] = sa;
          if (two) s[(r + 1) * L + l] = sb;
        }
        ca += sa;
        cb += sb;
      }
    }
    float* oa = out + r * (long long)D;
    float* ob = oa + D;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d0 = 128 * j + 4 * lane;
      const float4 bl = *reinterpret_cast<const float4*>(bl_s + d0);
      cross_store4(oa, vec_out != 0, D, d0,
                   make_float4(fmaf(ca, xa[j].x, bl.x), fmaf(ca, xa[j].y, bl.y), fmaf(ca, xa[j].z, bl.z),
                               fmaf(ca, xa[j].w, bl.w)));
      if (two)
        cross_store4(ob, vec_out != 0, D, d0,
                     make_float4(fmaf(cb, xb[j].x, bl.x), fmaf(cb, xb[j].y, bl.y), fmaf(cb, xb[j].z, bl.z),
                                 fmaf(cb, xb[j].w, bl.w)));
    }
  }
}

// ---- backward, kernel 1: per sample -------------------------------------------------------------
// partial_cg [grid][D], partial_T [grid][kCrossMaxL], u [B][kCrossMaxL]
template <int NJ>
__global__ void __launch_bounds__(kCrossThreads, 2)
cross_bwd_kernel(const float* __restrict__ x0, long long xs, const float* __restrict__ w,
                 const float* __restrict__ s, const float* __restrict__ g, long long gs,
                 float* __restrict__ dx0, float* __restrict__ u, float* __restrict__ partial_cg,
                 float* __restrict__ partial_T, long long B, int D, int L, int vec_x, int vec_g, int vec_dx,
                 int acc_dx) {
  constexpr int Dp = NJ * 128;
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                    // [L][Dp]
  float* q_s = w_s + L * Dp;            // unused constants slot (keeps the helper shared)
  float* red_s = q_s + kCrossMaxL;      // [kCrossWarps][kCrossMaxL]
  float* cg_s = red_s + kCrossWarps * kCrossMaxL;   // [kCrossWarps][Dp] (epilogue)
  cross_stage_constants(w, nullptr, D, Dp, L, w_s, nullptr, q_s, red_s);

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float4 cg[NJ];
  float Tacc[kCrossMaxL];
#pragma unroll
  for (int j = 0; j < NJ; ++j) cg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) Tacc[l] = 0.f;

  for (long long r = (long long)blockIdx.x * kCrossWarps + wid; r < B; r += (long long)gridDim.x * kCrossWarps) {
    float4 x[NJ], G[NJ];
    cross_load_row<NJ>(x0 + r * xs, vec_x != 0, D, lane, x);
    cross_load_row<NJ>(g + r * gs, vec_g != 0, D, lane, G);
    float a = 0.f, pl[kCrossMaxL];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      a = dot4(x[j], G[j], a);
      cg[j].x += G[j].x; cg[j].y += G[j].y; cg[j].z += G[j].z; cg[j].w += G[j].w;
    }
    a = warp_sum(a);
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      pl[l] = 0.f;
      if (l < L) {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          pl[l] = dot4(x[j], *reinterpret_cast<const float4*>(w_s + l * Dp + 128 * j + 4 * lane), pl[l]);
        pl[l] = warp_sum(pl[l]);
      }
    }
    // scalars: c_l from the saved s_l; t_l backwards; u_l = c_l t_l
    float cl[kCrossMaxL + 1], tl[kCrossMaxL], ul[kCrossMaxL];
    cl[0] = 1.f;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) cl[l + 1] = cl[l] + (l < L ? __ldg(s + r * L + l) : 0.f);
#pragma unroll
    for (int l = kCrossMaxL - 1; l >= 0; --l) {
      float t = 0.f;
      if (l < L) {
        t = a;
#pragma unroll
        for (int j = kCrossMaxL - 1; j > l; --j)
          if (j < L) t = fmaf(tl[j], pl[j], t);
      }
      tl[l] = t;
      ul[l] = cl[l] * t;
      Tacc[l] += t;
    }
    if (lane < kCrossMaxL) {
      float v = 0.f;
#pragma unroll
      for (int l = 0; l < kCrossMaxL; ++l) v = lane == l ? ul[l] : v;
      u[r * kCrossMaxL + lane] = v;
    }
    // dx0 = c_L g + sum_l u_l w_l
    const float cL = cl[kCrossMaxL];
    float* dp = dx0 + r * (long long)D;
    if (acc_dx) {        // dx0 already holds another consumer's gradient of x0 (kon_cross_bwd_acc): all of the row's
#pragma unroll           // loads in flight at once, in the registers x0 no longer needs
      for (int j = 0; j < NJ; ++j) x[j] = cross_load4_rw(dp, vec_dx != 0, D, 128 * j + 4 * lane);
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d0 = 128 * j + 4 * lane;
      float4 o = make_float4(cL * G[j].x, cL * G[j].y, cL * G[j].z, cL * G[j].w);
      if (acc_dx) {
        o.x += x[j].x; o.y += x[j].y; o.z += x[j].z; o.w += x[j].w;
      }
#pragma unroll
      for (int l = 0; l < kCrossMaxL; ++l) {
        if (l < L) {
          const float4 wv = *reinterpret_cast<const float4*>(w_s + l * Dp + d0);
          o.x = fmaf(ul[l], wv.x, o.x); o.y = fmaf(ul[l], wv.y, o.y);
          o.z = fmaf(ul[l], wv.z, o.z); o.w = fmaf(ul[l], wv.w, o.w);
        }
      }
      cross_store4(dp, vec_dx != 0, D, d0, o);
    }
  }
  // epilogue: per-warp column sums / T -> CTA partial, warps in order
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    *reinterpret_cast<float4*>(cg_s + wid * Dp + 128 * j + 4 * lane) = cg[j];
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) red_s[wid * kCrossMaxL + l] = Tacc[l];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kCrossThreads) {
    float v = 0.f;
#pragma unroll
    for (int ws = 0; ws < kCrossWarps; ++ws) v += cg_s[ws * Dp + d];
    partial_cg[(long long)blockIdx.x * D + d] = v;
  }
  if (threadIdx.x < kCrossMaxL) {
    float v = 0.f;
#pragma unroll
    for (int ws = 0; ws < kCrossWarps; ++ws) v += red_s[ws * kCrossMaxL + threadIdx.x];
    partial_T[blockIdx.x * kCrossMaxL + threadIdx.x] = v;
  }
}

// ---- backward, kernel 2: M = X0^T U per CTA; thread = 4 columns, slabs of kCrossSlab samples ----
__global__ void __launch_bounds__(kCrossThreads)
cross_dw_kernel(const float* __restrict__ x0, long long xs, const float* __restrict__ u,
                float* __restrict__ partial_M, long long B, int D, int L, int vec_x) {
  __shared__ __align__(16) float u_s[2][kCrossSlab * kCrossMaxL];
  const int tid = threadIdx.x;
  const int d0 = 4 * tid;
  float4 M[kCrossMaxL];
#pragma unroll
  for (int l = 0; l < kCrossMaxL; ++l) M[l] = make_float4(0.f, 0.f, 0.f, 0.f);
  // every CTA owns one contiguous, evenly sized range of samples (balanced to within one sample);
  // the u rows of slab i+1 are staged while slab i is consumed (one barrier per slab)
  const long long lo = B * blockIdx.x / gridDim.x, hi = B * (blockIdx.x + 1) / gridDim.x;
  const int n_slabs = (int)((hi - lo + kCrossSlab - 1) / kCrossSlab);
  auto stage = [&](int sl, int buf) {
    const long long b0 = lo + (long long)sl * kCrossSlab;
    const int nb = (int)min((long long)kCrossSlab, hi - b0);
    u_s[buf][tid] = (tid < nb * kCrossMaxL) ? __ldg(u + b0 * kCrossMaxL + tid) : 0.f;
  };
  if (n_slabs > 0) stage(0, 0);
  for (int sl = 0; sl < n_slabs; ++sl) {
    const int buf = sl & 1;
    __syncthreads();                     // u_s[buf] staged; everyone is done reading u_s[buf ^ 1]
    if (sl + 1 < n_slabs) stage(sl + 1, buf ^ 1);
    const long long b0 = lo + (long long)sl * kCrossSlab;
    const int nb = (int)min((long long)kCrossSlab, hi - b0);
    if (d0 < D) {
      for (int i0 = 0; i0 < kCrossSlab; i0 += 8) {
        float4 xv[8];     // 8 independent 128-bit loads in flight per thread before any of them is consumed
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = i0 + q;
          xv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < nb) {
            const float* xp = x0 + (b0 + i) * xs;
            if (vec_x && d0 + 3 < D) {
              xv[q] = __ldg(reinterpret_cast<const float4*>(xp + d0));
            } else {
              xv[q].x = __ldg(xp + d0);
              if (d0 + 1 < D) xv[q].y = __ldg(xp + d0 + 1);
              if (d0 + 2 < D) xv[q].z = __ldg(xp + d0 + 2);
              if (d0 + 3 < D) xv[q].w = __ldg(xp + d0 + 3);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int i = i0 + q;
          const float4 ua = *reinterpret_cast<const float4*>(&u_s[buf][i * kCrossMaxL]);
          const float4 ub = *reinterpret_cast<const float4*>(&u_s[buf][i * kCrossMaxL + 4]);
          const float uu[kCrossMaxL] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
          for (int l = 0; l < kCrossMaxL; ++l) {
            M[l].x = fmaf(xv[q].x, uu[l], M[l].x); M[l].y = fmaf(xv[q].y, uu[l], M[l].y);
            M[l].z = fmaf(xv[q].z, uu[l], M[l].z); M[l].w = fmaf(xv[q].w, uu[l], M[l].w);
          }
        }
      }
    }
  }
  if (d0 < D) {
    float* p = partial_M + (long long)blockIdx.x * L * D;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) {
      if (l < L) {
        p[l * D + d0] = M[l].x;
        if (d0 + 1 < D) p[l * D + d0 + 1] = M[l].y;
        if (d0 + 2 < D) p[l * D + d0 + 2] = M[l].z;
        if (d0 + 3 < D) p[l * D + d0 + 3] = M[l].w;
      }
    }
  }
}

// ---- backward, kernel 3: fixed-order sums of the per-CTA partials --------------------------------
// grid (ceil(D/32), L + 2): rows y < L: M_y (n2 partials), y == L: colsum(g) (n1), y == L+1: T (n1, 8 wide)
// red layout: [L][D] M | [D] cg | [kCrossMaxL] T
__global__ void __launch_bounds__(256)
cross_reduce_kernel(const float* __restrict__ partial_M, int n2, const float* __restrict__ partial_cg,
                    const float* __restrict__ partial_T, int n1, float* __restrict__ red, int D, int L) {
  __shared__ float sm[8][32];
  const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int y = blockIdx.y;
  const int d = blockIdx.x * 32 + col;
  const float* base;
  long long stride;
  int n, width;
  if (y < L) { base = partial_M + (long long)y * D; stride = (long long)L * D; n = n2; width = D; }
  else if (y == L) { base = partial_cg; stride = D; n = n1; width = D; }
  else { base = partial_T; stride = kCrossMaxL; n = n1; width = kCrossMaxL; }
  const int per = (n + 7) / 8;
  const int lo = slice * per, hi = min(n, lo + per);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (d < width) {
    int p = lo;
    for (; p + 3 < hi; p += 4) {
      a0 += base[(long long)p * stride + d];
      a1 += base[(long long)(p + 1) * stride + d];
      a2 += base[(long long)(p + 2) * stride + d];
      a3 += base[(long long)(p + 3) * stride + d];
    }
    for (; p < hi; ++p) a0 += base[(long long)p * stride + d];
  }
  sm[slice][col] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (slice == 0 && d < width) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sm[k][col];
    red[(y < L ? (long long)y * D : (y == L ? (long long)L * D : (long long)L * D + D)) + d] = v;
  }
}

// ---- backward, kernel 4: dw / db from the reduced sums ---------------------------------------------
__global__ void __launch_bounds__(256)
cross_bwd_finalize_kernel(const float* __restrict__ red, const float* __restrict__ w,
                          const float* __restrict__ b, float* __restrict__ dw, float* __restrict__ db,
                          int D, int L) {
  const float* T = red + (long long)L * D + D;
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
    const float cg = red[(long long)L * D + d];
    float beta = 0.f;   // beta_l[d] = sum_{j<l} b_j[d]
    for (int l = 0; l < L; ++l) {
      dw[l * D + d] = red[(long long)l * D + d] + beta * T[l];
      beta += b[l * D + d];
    }
    float tail = 0.f;   // sum_{j>l} w_j[d] T_j
    for (int l = L - 1; l >= 0; --l) {
      db[l * D + d] = cg + tail;
      tail = fmaf(w[l * D + d], T[l], tail);
    }
  }
}

static int cross_grid1(long long B, int sms) {
  return (int)std::max<long long>(1, std::min<long long>((B + kCrossWarps - 1) / kCrossWarps, (long long)sms * 2));   // = resident CTAs
}
static int cross_grid2(long long B, int sms) {
  return (int)std::max<long long>(1, std::min<long long>((B + kCrossSlab - 1) / kCrossSlab, (long long)sms * 4));
}
static size_t cross_align(size_t x) { return (x + 255) / 256 * 256; }

struct CrossWs {
  size_t u, pm, pcg, pt, red, total;
};
static CrossWs cross_ws_layout(long long B, int D, int L, int sms) {
  CrossWs l;
  const int g1 = cross_grid1(B, sms), g2 = cross_grid2(B, sms);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = cross_align(o + bytes); return r; };
  l.u = take((size_t)std::max<long long>(B, 1) * kCrossMaxL * 4);
  l.pm = take((size_t)g2 * L * D * 4);
  l.pcg = take((size_t)g1 * D * 4);
  l.pt = take((size_t)g1 * kCrossMaxL * 4);
  l.red = take(((size_t)L * D + D + kCrossMaxL) * 4);
  l.total = o;
  return l;
}

}  // namespace kon

using namespace kon;

static int cross_check(const DLTensor* x0, const DLTensor* w, const DLTensor* b, int64_t* B,
                       int64_t* D, int64_t* L) {
  KON_TRY(check_cuda_tensor(x0, "x0"));
  const int dev = x0->device.device_id;
  KON_TRY(check_cuda_tensor(w, "w", dev));
  KON_TRY(check_cuda_tensor(b, "b", dev));
  KON_REQUIRE(is_f32(x0) && x0->ndim == 2 && stride_of(x0, 1) == 1, KON_EINVAL,
              "x0 must be float32 [B,D] with a compact last dim");
  *B = x0->shape[0];
  *D = x0->shape[1];
  KON_REQUIRE(is_f32(w) && w->ndim == 2 && w->shape[1] == *D && is_compact(w), KON_EINVAL,
              "w must be compact float32 [L,D]");
  *L = w->shape[0];
  KON_REQUIRE(is_f32(b) && b->ndim == 2 && b->shape[0] == *L && b->shape[1] == *D && is_compact(b),
              KON_EINVAL, "b must be compact float32 [L,D]");
  KON_REQUIRE(*L >= 1 && *L <= kCrossMaxL, KON_EUNSUPPORTED, "cross layers %lld outside [1,%d]",
              (long long)*L, kCrossMaxL);
  KON_REQUIRE(*D >= 1 && *D <= 1024, KON_EUNSUPPORTED, "cross width D=%lld outside [1,1024]",
              (long long)*D);
  return KON_OK;
}

#define KON_CROSS_DISPATCH(D, CALL)            \
  do {                                         \
    const int nj = (int)(((D) + 127) / 128);   \
    if (nj <= 1) { CALL(1); }                  \
    else if (nj <= 2) { CALL(2); }             \
    else if (nj <= 4) { CALL(4); }             \
    else if (nj <= 7) { CALL(7); }             \
    else { CALL(8); }                          \
  } while (0)

static int cross_nj(int64_t D) {
  const int nj = (int)((D + 127) / 128);
  return nj <= 1 ? 1 : nj <= 2 ? 2 : nj <= 4 ? 4 : nj <= 7 ? 7 : 8;
}
// rows can move as 128-bit accesses when every row start is 16-B aligned
static int cross_vec(const float* p, int64_t row_stride) {
  return (aligned16(p) && row_stride % 4 == 0) ? 1 : 0;
}

extern "C" int kon_cross_fwd(const DLTensor* x0, const DLTensor* w, const DLTensor* b,
                             DLTensor* out, DLTensor* s, void* stream) {
  int64_t B, D, L;
  KON_TRY(cross_check(x0, w, b, &B, &D, &L));
  const int dev = x0->device.device_id;
  KON_TRY(check_cuda_tensor(out, "out", dev));
  KON_TRY(check_cuda_tensor(s, "s", dev));
  KON_REQUIRE(is_f32(out) && out->ndim == 2 && out->shape[0] == B && out->shape[1] == D &&
                  is_compact(out),
              KON_EINVAL, "out must be compact float32 [B,D]");
  KON_REQUIRE(is_f32(s) && s->ndim == 2 && s->shape[0] == B && s->shape[1] == L && is_compact(s),
              KON_EINVAL, "s must be compact float32 [B,L]");
  if (B == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Dp = cross_nj(D) * 128;
  const size_t smem = ((size_t)L * Dp + Dp + kCrossMaxL + kCrossWarps * kCrossMaxL) * sizeof(float);
  const int grid = (int)std::min<long long>((B + 2 * kCrossWarps - 1) / (2 * kCrossWarps),
                                            (long long)sm_count_of(dev) * 2);   // = resident CTAs
  const int vin = cross_vec(data_ptr<float>(x0), stride_of(x0, 0));
  const int vout = cross_vec(data_ptr<float>(out), D);
  ProfileScope ps("cross_fwd_kernel", st);
#define CALL(N)                                                                                  \
  KON_CUDA(cudaFuncSetAttribute(cross_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem));                                                     \
  cross_fwd_kernel<N><<<grid, kCrossThreads, smem, st>>>(                                        \
      data_ptr<float>(x0), stride_of(x0, 0), data_ptr<float>(w), data_ptr<float>(b),             \
      data_ptr<float>(out), data_ptr<float>(s), B, (int)D, (int)L, vin, vout)
  KON_CROSS_DISPATCH(D, CALL);
#undef CALL
  KON_LAUNCH_CHECK("cross_fwd_kernel");
  return KON_OK;
}

extern "C" size_t kon_cross_bwd_workspace_bytes(int64_t batch, int32_t dim, int32_t layers,
                                                int device_id) {
  return cross_ws_layout(batch, dim, layers, sm_count_of(device_id)).total;
}

static int cross_bwd_impl(const DLTensor* x0, const DLTensor* w, const DLTensor* b,
                          const DLTensor* s, const DLTensor* g, DLTensor* dx0, DLTensor* dw,
                          DLTensor* db, DLTensor* workspace, int acc_dx, void* stream) {
  int64_t B, D, L;
  KON_TRY(cross_check(x0, w, b, &B, &D, &L));
  const int dev = x0->device.device_id;
  KON_TRY(check_cuda_tensor(s, "s", dev));
  KON_TRY(check_cuda_tensor(g, "g", dev));
  KON_TRY(check_cuda_tensor(dx0, "dx0", dev));
  KON_TRY(check_cuda_tensor(dw, "dw", dev));
  KON_TRY(check_cuda_tensor(db, "db", dev));
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  KON_REQUIRE(is_f32(s) && s->ndim == 2 && s->shape[0] == B && s->shape[1] == L && is_compact(s),
              KON_EINVAL, "s must be compact float32 [B,L]");
  KON_REQUIRE(is_f32(g) && g->ndim == 2 && g->shape[0] == B && g->shape[1] == D &&
                  stride_of(g, 1) == 1,
              KON_EINVAL, "g must be float32 [B,D] with a compact last dim");
  KON_REQUIRE(is_f32(dx0) && dx0->ndim == 2 && dx0->shape[0] == B && dx0->shape[1] == D &&
                  is_compact(dx0),
              KON_EINVAL, "dx0 must be compact float32 [B,D]");
  KON_REQUIRE(is_f32(dw) && numel(dw) == L * D && is_compact(dw) && is_f32(db) &&
                  numel(db) == L * D && is_compact(db),
              KON_EINVAL, "dw and db must be compact float32 [L,D]");
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = sm_count_of(dev);
  const CrossWs l = cross_ws_layout(B, (int)D, (int)L, sms);
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= l.total, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), l.total);
  char* ws = data_ptr<char>(workspace);
  KON_REQUIRE(((uintptr_t)ws & 15u) == 0, KON_EINVAL, "workspace must be 16-B aligned");
  float* u = (float*)(ws + l.u);
  float* pm = (float*)(ws + l.pm);
  float* pcg = (float*)(ws + l.pcg);
  float* pt = (float*)(ws + l.pt);
  float* red = (float*)(ws + l.red);
  const int g1 = cross_grid1(B, sms), g2 = cross_grid2(B, sms);
  const int Dp = cross_nj(D) * 128;
  const size_t smem = ((size_t)L * Dp + kCrossMaxL + kCrossWarps * kCrossMaxL + (size_t)kCrossWarps * Dp) *
                      sizeof(float);
  const int vx = cross_vec(data_ptr<float>(x0), stride_of(x0, 0));
  const int vg = cross_vec(data_ptr<float>(g), stride_of(g, 0));
  const int vdx = cross_vec(data_ptr<float>(dx0), D);
  ProfileScope ps("cross_bwd_kernels", st);     // per-sample kernel + dw kernel + partial reduction + finalize
#define CALL(N)                                                                                  \
  KON_CUDA(cudaFuncSetAttribute(cross_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem));                                                     \
  cross_bwd_kernel<N><<<g1, kCrossThreads, smem, st>>>(                                          \
      data_ptr<float>(x0), stride_of(x0, 0), data_ptr<float>(w), data_ptr<float>(s),             \
      data_ptr<float>(g), stride_of(g, 0), data_ptr<float>(dx0), u, pcg, pt, B, (int)D, (int)L,  \
      vx, vg, vdx, acc_dx)
  KON_CROSS_DISPATCH(D, CALL);
#undef CALL
  KON_LAUNCH_CHECK("cross_bwd_kernel");
  cross_dw_kernel<<<g2, kCrossThreads, 0, st>>>(data_ptr<float>(x0), stride_of(x0, 0), u, pm, B, (int)D,
                                                (int)L, vx);
  KON_LAUNCH_CHECK("cross_dw_kernel");
  cross_reduce_kernel<<<dim3((unsigned)((D + 31) / 32), (unsigned)(L + 2)), 256, 0, st>>>(pm, g2, pcg, pt, g1,
                                                                                       red, (int)D, (int)L);
  KON_LAUNCH_CHECK("cross_reduce_kernel");
  cross_bwd_finalize_kernel<<<(int)((D + 255) / 256), 256, 0, st>>>(red, data_ptr<float>(w), data_ptr<float>(b),
                                                                    data_ptr<float>(dw), data_ptr<float>(db),
                                                                    (int)D, (int)L);
  KON_LAUNCH_CHECK("cross_bwd_finalize_kernel");
  return KON_OK;
}

extern "C" int kon_cross_bwd(const DLTensor* x0, const DLTensor* w, const DLTensor* b,
                             const DLTensor* s, const DLTensor* g, DLTensor* dx0, DLTensor* dw,
                             DLTensor* db, DLTensor* workspace, void* stream) {
  return cross_bwd_impl(x0, w, b, s, g, dx0, dw, db, workspace, 0, stream);
}

extern "C" int kon_cross_bwd_acc(const DLTensor* x0, const DLTensor* w, const DLTensor* b,
                                 const DLTensor* s, const DLTensor* g, DLTensor* dx0, DLTensor* dw,
                                 DLTensor* db, DLTensor* workspace, void* stream) {
  return cross_bwd_impl(x0, w, b, s, g, dx0, dw, db, workspace, 1, stream);
}
