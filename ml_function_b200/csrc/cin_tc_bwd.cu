// a8 backward on tcgen05 (KON_CIN_BF16).  Per layer l (last to first), with dZ_l[r,o] the
// gradient of z_l (pooled-grad broadcast over o, IL:322, plus dpre_{l+1}):
//
//   dW kernel  (stream skeleton of cin_tc_common.cuh):
//       dW_l[c,o] = sum_r A[r,c] dZ_l[r,o],   A[r,c] = pre[r,h] x0[r,i]  regenerated on the fly.
//       GEMM with M-side = c (TMEM lane = c, 2 x 128 lanes per CTA), N = o, K = rows.  One k-step
//       (16 rows) is one sample's 16 embedding coordinates, so lane c multiplies two contiguous
//       32-byte vectors pre[b,h,:] * x0[b,i,:] (bf16x2 HMUL2) and stores them to the TMEM ring.
//       B operand = dZ_l in a blocked 8x8 layout (MN-major core matrices) streamed with bulk copies.
//       An all-ones lane c == C yields dbias.  Work = (c-tile pair) x (row slice); the slices'
//       partial sums are reduced by a second kernel in a fixed order (deterministic).
//   dA kernel:
//       dA[r,c] = sum_o dZ_l[r,o] W_l[c,o]  (never stored), contracted in the epilogue:
//       dpre_l[r,h] = sum_i dA[r,h*m+i] x0[r,i],   dx0[r,i] += sum_h dA[r,h*m+i] pre[r,h].
//       A operand = the dZ tile, written once per 128-row tile into TMEM (13 k-steps);
//       B operand = chunks of 4*m = 104 weight rows (K-major), 13 MMAs per chunk and sub-tile;
//       the two sub-tiles ping-pong: while the row warps of one contract their 104 accumulator
//       columns (packed fp32x2 FMAs), the tensor core fills the other's.
//       dZ_{l-1} = dpre_l + pooled-grad broadcast is written in the blocked layout the next
//       iteration's two kernels consume.
#include "cin_tc_common.cuh"

namespace kon {

using namespace tcs;

namespace {

static_assert(kSub == 2, "the backward kernels are written for two 128-lane sub-tiles per CTA");

// ---------------------------------------------------------------------------------------------
// small helper kernels
// ---------------------------------------------------------------------------------------------
// x0 fp32 [B,m,D] (batch stride sb) -> x0b bf16 [B,m,D] compact and, for D == 16, x0h bf16 [B][2][m][8]
// (the two 16-byte halves of every field's 16 coordinates in separate planes: 32 lanes reading 32
// different fields hit 32 different banks).  One thread per 8 consecutive coordinates.
__global__ void __launch_bounds__(256)
cin_x0_convert_kernel(const float* __restrict__ x0, long long sb, unsigned short* __restrict__ x0b,
                      unsigned short* __restrict__ x0h, long long B, int m, int D) {
  const int d8 = D / 8;
  const long long total = B * m * d8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(idx % d8);
    const long long t = idx / d8;
    const int i = (int)(t % m);
    const long long b = t / m;
    const float4* src = reinterpret_cast<const float4*>(x0 + b * sb + (long long)i * D + q * 8);
    const float4 lo = src[0], hi = src[1];
    const uint4 pk = make_uint4(tc::pack_bf16(lo.x, lo.y), tc::pack_bf16(lo.z, lo.w), tc::pack_bf16(hi.x, hi.y),
                                tc::pack_bf16(hi.z, hi.w));
    *reinterpret_cast<uint4*>(x0b + (b * m + i) * (long long)D + q * 8) = pk;
    if (x0h) *reinterpret_cast<uint4*>(x0h + b * (m * 16) + q * (m * 8) + i * 8) = pk;
  }
}

// blocked dZ: [rows/8][N8][8 rows][8 cols] bf16; element (r,o) at ((r/8)*N8 + o/8)*128 + (r%8)*16 + (o%8)*2
__global__ void __launch_bounds__(256)
cin_dz_init_kernel(const float* __restrict__ gpool, int gstride, int gcol, long long rows, int D, int N,
                   int N8, unsigned char* __restrict__ dz) {
  const long long total = rows * N8;      // one 16-byte unit per (row, column group)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // consecutive idx -> consecutive 16-byte units of the blocked buffer
    const int r7 = (int)(idx & 7);
    const long long t = idx >> 3;
    const int o8 = (int)(t % N8);
    const long long rg = t / N8;
    const long long r = rg * 8 + r7;
    const long long b = r / D;
    const float g = gpool[b * gstride + gcol + (int)(r - b * D)];
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int o = o8 * 8 + 2 * q;
      w[q] = tc::pack_bf16(o < N ? g : 0.f, o + 1 < N ? g : 0.f);
    }
    *reinterpret_cast<uint4*>(dz + idx * 16) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// dA B operand, K-major B[n = c_local][k = o], chunk j = weight rows [CH*j, CH*j+CH):
//   offset = j*chunk + s*(2*CH8*128) + kg*CH8*128 + (c_local/8)*128 + (c%8)*16 + (o%8)*2,  o = 16s + 8kg + o%8
__global__ void __launch_bounds__(256)
cin_pack_w_bwd_kernel(const float* __restrict__ W, int C, int N, int CH8, int nkA, int n_chunks,
                      __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)n_chunks * nkA * 2 * CH8 * 64;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int o7 = (int)(idx & 7);
    const int c7 = (int)((idx >> 3) & 7);
    long long t = idx >> 6;
    const int n8 = (int)(t % CH8);
    t /= CH8;
    const int kg = (int)(t & 1);
    t >>= 1;
    const int s = (int)(t % nkA);
    const long long j = t / nkA;
    const long long c = j * (CH8 * 8) + n8 * 8 + c7;
    const int o = 16 * s + 8 * kg + o7;
    float v = 0.f;
    if (c < C && o < N) v = W[c * N + o];
    out[idx] = __float2bfloat16_rn(v);
  }
}

// dW[c,o] = sum_slices part[s][c][o] (c < C);  dbias[o] = sum_slices part[s][C][o]
__global__ void __launch_bounds__(256)
cin_dw_reduce_kernel(const float* __restrict__ part, int n_slices, long long slice_stride, int Npad, int C,
                     int N, float* __restrict__ dW, float* __restrict__ dbias) {
  const long long total = (long long)(C + 1) * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / N;
    const int o = (int)(idx - c * N);
    const float* p = part + c * Npad + o;
    float acc = 0.f;
    for (int s = 0; s < n_slices; ++s) acc += p[s * slice_stride];
    if (c < C) dW[idx] = acc;
    else dbias[o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Last layer: the reference pools over the feature maps (IL:322), so dZ_L[r,o] = g[r] for every o and
// the two backward GEMMs of the LAST layer collapse:
//     dA[r,c] = g[r] * wsum[c],  wsum[c] = sum_o W[c,o]
//     dpre[r,h] = g[r] * sum_i wsum[h*m+i] x0[r,i]          dx0[r,i] += g[r] * sum_h wsum[h*m+i] pre[r,h]
//     dW[c,o]  = v[c] for every o,  v = A^T g   (computed by the dW kernel with a 16-column B operand)
// ---------------------------------------------------------------------------------------------
constexpr int kTRow = 28;     // wsum rows padded to 28 floats (16-byte aligned rows, 2 zero columns)

// T[h*28 + i] = sum_o W[(h*m+i)*N + o]; one warp per weight row
__global__ void __launch_bounds__(256)
cin_wsum_kernel(const float* __restrict__ W, int C, int N, int m, int Hp8, float* __restrict__ T) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)Hp8 * kTRow;
  for (long long t = w0; t < total; t += nw) {
    const int h = (int)(t / kTRow), i = (int)(t - (long long)h * kTRow);
    const long long c = (long long)h * m + i;
    float acc = 0.f;
    if (i < m && c < C)
      for (int o = lane; o < N; o += 32) acc += W[c * N + o];
    acc = warp_sum(acc);
    if (lane == 0) T[t] = acc;
  }
}

struct LastDaArgs {
  const unsigned short* x0b;   // [B,m,D] bf16
  const unsigned short* pre;   // [B,Hp,D] bf16
  const float* T;              // [Hp8, 28]
  const float* gpool;          // d_pooled [B, gstride]
  int gstride, gcol, gcol_prev;
  unsigned char* dz_prev;      // blocked, N8p column groups
  float* dx0;                  // [B,m,D] accumulated; batch stride dx_sb
  long long dx_sb;
  long long rows;
  int D, Hp, Hp8, N8p;
};

__device__ __forceinline__ void ffma2_acc(float2& acc, const float2 a, const float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(rc));
}

template <int MF>
__global__ void __launch_bounds__(256) cin_last_da_kernel(const LastDaArgs a) {
  static_assert(MF <= kTRow && MF % 2 == 0, "field count");
  extern __shared__ __align__(16) float sT[];                  // [Hp8][28]
  for (int i = threadIdx.x; i < a.Hp8 * kTRow; i += blockDim.x) sT[i] = a.T[i];
  __syncthreads();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < a.rows;
       r += (long long)gridDim.x * blockDim.x) {
    const long long b = r / a.D;
    const int d = (int)(r - b * a.D);
    const unsigned short* xrow = a.x0b + b * (long long)MF * a.D + d;
    const unsigned short* prow = a.pre + b * (long long)a.Hp * a.D + d;
    float2 x[kTRow / 2], dx[kTRow / 2];
#pragma unroll
    for (int i = 0; i < kTRow / 2; ++i) {
      x[i].x = 2 * i < MF ? bf16_to_f32(__ldg(xrow + (long long)(2 * i) * a.D)) : 0.f;
      x[i].y = 2 * i + 1 < MF ? bf16_to_f32(__ldg(xrow + (long long)(2 * i + 1) * a.D)) : 0.f;
      dx[i] = make_float2(0.f, 0.f);
    }
    const float g = a.gpool[b * a.gstride + a.gcol + d];
    const float gprev = a.gpool[b * a.gstride + a.gcol_prev + d];
    unsigned short praw[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) praw[q] = __ldg(prow + (long long)min(q, a.Hp - 1) * a.D);
    for (int h0 = 0; h0 < a.Hp8; h0 += 8) {
      float pv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) pv[q] = (h0 + q < a.Hp) ? bf16_to_f32(praw[q]) * g : 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) praw[q] = __ldg(prow + (long long)min(h0 + 8 + q, a.Hp - 1) * a.D);
      uint32_t outw[4];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4* trow = reinterpret_cast<const float4*>(sT + (h0 + q) * kTRow);
        float2 s2 = make_float2(0.f, 0.f);
        const float2 gp2 = make_float2(pv[q], pv[q]);
#pragma unroll
        for (int z = 0; z < kTRow / 4; ++z) {
          const float4 t = trow[z];
          ffma2_acc(s2, make_float2(t.x, t.y), x[2 * z]);
          ffma2_acc(s2, make_float2(t.z, t.w), x[2 * z + 1]);
          ffma2_acc(dx[2 * z], make_float2(t.x, t.y), gp2);
          ffma2_acc(dx[2 * z + 1], make_float2(t.z, t.w), gp2);
        }
        const float v = (h0 + q < a.Hp) ? g * (s2.x + s2.y) + gprev : 0.f;
        if (q & 1) outw[q >> 1] = tc::pack_bf16(__uint_as_float(outw[q >> 1]), v);
        else outw[q >> 1] = __float_as_uint(v);
      }
      if ((h0 >> 3) < a.N8p)
        *reinterpret_cast<uint4*>(a.dz_prev + ((r >> 3) * a.N8p + (h0 >> 3)) * 128 + (r & 7) * 16) =
            make_uint4(outw[0], outw[1], outw[2], outw[3]);
    }
    float* dxr = a.dx0 + b * a.dx_sb + d;
    float old[MF];
#pragma unroll
    for (int i = 0; i < MF; ++i) old[i] = dxr[(long long)i * a.D];
#pragma unroll
    for (int i = 0; i < MF / 2; ++i) {
      dxr[(long long)(2 * i) * a.D] = old[2 * i] + dx[i].x;
      dxr[(long long)(2 * i + 1) * a.D] = old[2 * i + 1] + dx[i].y;
    }
  }
}

// v[h,i] = sum_b sum_d pre[b,h,d] * (x0[b,i,d] g[b,d])   (D == 16): per sample one m16n8k16 k-step.
// Warp w owns the 16 feature maps h = 16w..16w+15 and all 4 n-tiles (i < 32); A fragments come
// straight from the bf16 z^T rows in global memory, B = x0*g is built per sample in shared memory.
// (Legacy mma.sync on purpose: 5.4 GFMA of side work, HBM-bound on the 419 MB of `pre`.)
constexpr int kLdwBatch = 4;            // samples per __syncthreads
constexpr int kLdwRow = 24;             // bf16 per padded smem row (12 words: conflict-free fragment reads)

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct LastDwArgs {
  const unsigned short* x0b;   // [B,m,16] bf16
  const unsigned short* pre;   // [B,Hp,16] bf16
  const float* gpool;
  int gstride, gcol;
  float* part;                 // [grid][Hp16][32] fp32 partial v  (+ [grid] partial sum of g after it)
  float* gsum;                 // [grid]
  long long B;
  int Hp, m;
};

__global__ void __launch_bounds__(416) cin_last_dw_kernel(const LastDwArgs a) {
  __shared__ __align__(16) unsigned short sX[2][kLdwBatch][32 * kLdwRow];
  __shared__ float s_g[13];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g8 = lane >> 2, t4 = lane & 3;
  const int n_mt = (a.Hp + 15) / 16;
  const long long per = (a.B + gridDim.x - 1) / gridDim.x;
  const long long b0 = blockIdx.x * per, b1 = min(a.B, b0 + per);
  // zero the padded rows (i >= m) once
  for (int i = tid; i < 2 * kLdwBatch * 32 * kLdwRow; i += blockDim.x) (&sX[0][0][0])[i] = 0;
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[n][q] = 0.f;
  float gacc = 0.f;
  const int h_lo = min(warp * 16 + g8, a.Hp - 1), h_hi = min(warp * 16 + g8 + 8, a.Hp - 1);
  const bool lo_on = warp * 16 + g8 < a.Hp, hi_on = warp * 16 + g8 + 8 < a.Hp;
  const int xi = tid / 16, xd = tid % 16;        // this thread's element of x0*g (416 = 26*16 threads)
  // x0 * g of a batch: the global loads (fetch) are issued one iteration before the values are rounded and
  // stored (stash), so their latency overlaps the MMAs of the batch in between
  float fg[kLdwBatch];
  unsigned short fx[kLdwBatch];
  auto fetch = [&](long long bb) {
#pragma unroll
    for (int u = 0; u < kLdwBatch; ++u) {
      const long long b = min(bb + u, b1 - 1);
      const int xc = min(xi, a.m - 1);
      fg[u] = __ldg(a.gpool + b * a.gstride + a.gcol + xd);
      fx[u] = __ldg(a.x0b + (b * a.m + xc) * 16 + xd);
    }
  };
  auto stash = [&](long long bb, int buf) {
#pragma unroll
    for (int u = 0; u < kLdwBatch; ++u) {
      const long long b = bb + u;
      float v = 0.f;
      if (b < b1 && xi < a.m) {
        v = bf16_to_f32(fx[u]) * fg[u];
        if (xi == 0) gacc += fg[u];
      }
      const __nv_bfloat16 hv = __float2bfloat16_rn(v);
      sX[buf][u][xi * kLdwRow + xd] = *reinterpret_cast<const unsigned short*>(&hv);
    }
  };
  auto load_a = [&](long long bb, uint32_t (*af)[4]) {
#pragma unroll
    for (int u = 0; u < kLdwBatch; ++u) {
      const long long b = min(bb + u, a.B - 1);
      const uint32_t* plo = reinterpret_cast<const uint32_t*>(a.pre + (b * a.Hp + h_lo) * 16);
      const uint32_t* phi = reinterpret_cast<const uint32_t*>(a.pre + (b * a.Hp + h_hi) * 16);
      af[u][0] = __ldg(plo + t4);
      af[u][1] = __ldg(phi + t4);
      af[u][2] = __ldg(plo + t4 + 4);
      af[u][3] = __ldg(phi + t4 + 4);
    }
  };
  uint32_t an[kLdwBatch][4];
  if (b0 < b1) {
    fetch(b0);
    if (warp < n_mt) load_a(b0, an);
    stash(b0, 0);
    if (b0 + kLdwBatch < b1) fetch(b0 + kLdwBatch);
  }
  __syncthreads();
  int buf = 0;
  for (long long bb = b0; bb < b1; bb += kLdwBatch, buf ^= 1) {
    uint32_t ac[kLdwBatch][4];
#pragma unroll
    for (int u = 0; u < kLdwBatch; ++u)
#pragma unroll
      for (int q = 0; q < 4; ++q) ac[u][q] = an[u][q];
    const bool more = bb + kLdwBatch < b1;
    if (more && warp < n_mt) load_a(bb + kLdwBatch, an);
    if (warp < n_mt) {
#pragma unroll
      for (int u = 0; u < kLdwBatch; ++u) {
        if (bb + u < b1) {
          uint32_t af[4] = {lo_on ? ac[u][0] : 0u, hi_on ? ac[u][1] : 0u, lo_on ? ac[u][2] : 0u, hi_on ? ac[u][3] : 0u};
          const uint32_t* xs = reinterpret_cast<const uint32_t*>(&sX[buf][u][0]);
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const uint32_t bq0 = xs[(n * 8 + g8) * (kLdwRow / 2) + t4];
            const uint32_t bq1 = xs[(n * 8 + g8) * (kLdwRow / 2) + t4 + 4];
            mma_bf16_16816(acc[n], af, bq0, bq1);
          }
        }
      }
    }
    if (more) {
      stash(bb + kLdwBatch, buf ^ 1);                                   // fetched one iteration ago
      if (bb + 2 * kLdwBatch < b1) fetch(bb + 2 * kLdwBatch);
    }
    __syncthreads();
  }
  // partial results: part[cta][h][i], h < Hp16 = 16*n_mt, i < 32
  if (warp < n_mt) {
    float* p = a.part + ((long long)blockIdx.x * n_mt * 16 + warp * 16) * 32;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      p[(g8) * 32 + n * 8 + 2 * t4] = acc[n][0];
      p[(g8) * 32 + n * 8 + 2 * t4 + 1] = acc[n][1];
      p[(g8 + 8) * 32 + n * 8 + 2 * t4] = acc[n][2];
      p[(g8 + 8) * 32 + n * 8 + 2 * t4 + 1] = acc[n][3];
    }
  }
  // sum of g over this CTA's samples (threads with xi == 0 hold one coordinate each)
  gacc = warp_sum(gacc);
  if (lane == 0) s_g[warp] = gacc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 13; ++w) t += s_g[w];
    a.gsum[blockIdx.x] = t;
  }
}

// dW[c,o] = sum_cta part[cta][h][i] for every o;  dbias[o] = sum_cta gsum[cta].  One warp per weight
// row c: the lanes split the CTA partials in a fixed pattern, warp-reduce, then write the row.
__global__ void __launch_bounds__(256)
cin_last_dw_reduce_kernel(const float* __restrict__ part, const float* __restrict__ gsum, int n_cta, int Hp16,
                          int m, int C, int N, float* __restrict__ dW, float* __restrict__ dbias) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long c = w0; c <= C; c += nw) {
    float acc = 0.f;
    if (c < C) {
      const int h = (int)(c / m), i = (int)(c % m);
      const float* p = part + (long long)h * 32 + i;
      for (int s = lane; s < n_cta; s += 32) acc += p[(long long)s * Hp16 * 32];
    } else {
      for (int s = lane; s < n_cta; s += 32) acc += gsum[s];
    }
    acc = warp_sum(acc);
    float* dst = c < C ? dW + c * N : dbias;
    for (int o = lane; o < N; o += 32) dst[o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Last layer on tcgen05: both mat-vecs of cin_last_da_kernel as two tiny GEMMs per 128-row tile,
//   D1[r,h] = sum_i x0[r,i] T[h,i]   (A = x0 row in TMEM, K = 32;  B = T   [Nh x 32]  in smem, constant)
//   D2[r,i] = sum_h pre[r,h] T[h,i]  (A = pre row in TMEM, K = Hp16; B = T^T [32 x Hp16] in smem, constant)
// then dZ_{L-1}[r,h] = g[r] D1 + g_prev[r]  and  dx0[r,i] += g[r] D2.  15 MMAs per tile: the kernel is
// bound by moving the rows (419 MB of pre in, 419 MB of dZ out), not by math.
// ---------------------------------------------------------------------------------------------
// T[h,i] = sum_o W[(h*m+i),o] packed twice as bf16 core matrices:
//   Th (K-major B[n=h][k=i]):  off = s*(2*Nh8*128) + kg*Nh8*128 + (h/8)*128 + (h%8)*16 + (i%8)*2,  i = 16s+8kg+i%8
//   Ti (K-major B[n=i][k=h]):  off = s*1024        + kg*512      + (i/8)*128 + (i%8)*16 + (h%8)*2,  h = 16s+8kg+h%8
__global__ void __launch_bounds__(256)
cin_tpack_kernel(const float* __restrict__ T, int Hp, int m, int Nh8, int nkh, unsigned short* __restrict__ th,
                 unsigned short* __restrict__ ti) {
  const int n_th = 2 * 2 * Nh8 * 64, n_ti = nkh * 2 * 4 * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_th + n_ti; idx += gridDim.x * blockDim.x) {
    int h, i;
    if (idx < n_th) {
      const int i7 = idx & 7, h7 = (idx >> 3) & 7;
      int t = idx >> 6;
      const int h8 = t % Nh8;
      t /= Nh8;
      const int kg = t & 1, s = t >> 1;
      h = h8 * 8 + h7;
      i = 16 * s + 8 * kg + i7;
    } else {
      const int j = idx - n_th;
      const int h7 = j & 7, i7 = (j >> 3) & 7;
      int t = j >> 6;
      const int i8 = t & 3;
      t >>= 2;
      const int kg = t & 1, s = t >> 1;
      i = i8 * 8 + i7;
      h = 16 * s + 8 * kg + h7;
    }
    const float v = (h < Hp && i < m) ? T[h * kTRow + i] : 0.f;
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    const unsigned short u = *reinterpret_cast<const unsigned short*>(&b);
    if (idx < n_th) th[idx] = u;
    else ti[idx - n_th] = u;
  }
}

struct LastDaTcArgs {
  const unsigned short* x0b;   // [B,m,D] bf16
  const unsigned short* pre;   // [B,Hp,D] bf16
  const unsigned char* tpack;  // Th then Ti
  const float* gpool;
  int gstride, gcol, gcol_prev;
  unsigned char* dz_prev;      // blocked, N8p column groups
  float* dx0;
  long long dx_sb, rows, n_tiles;
  int D, Hp, N8p, Nh8, nkh;
};

constexpr uint32_t kLtD1 = 0, kLtD2 = 208, kLtAx = 240, kLtAp = 256;

// Warps 0-3: rows (one TMEM lane each); warp 4: MMA issue; warp 5: bulk-copy loader that streams each
// tile's `pre` rows (128*Hp*2 contiguous bytes) and x0 rows one tile ahead into a 2-stage smem ring, so
// the row warps never wait on HBM.
template <int MF>
__global__ void __launch_bounds__(192, 1) cin_last_da_tc_kernel(const LastDaTcArgs a) {
  static_assert(MF <= 32 && MF % 2 == 0, "field count");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t a_full, d_full, s_full[2], s_empty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t th_bytes = 2u * 2u * a.Nh8 * 128u, ti_bytes = (uint32_t)a.nkh * 1024u;
  const uint32_t tab_bytes = (th_bytes + ti_bytes + 127u) & ~127u;
  const uint32_t pre_tile = 128u * a.Hp * 2u, x_tile = 128u * MF * 2u;          // bytes per full tile
  const uint32_t stage_bytes = (pre_tile + x_tile + 127u) & ~127u;
  for (uint32_t i = tid; i < (th_bytes + ti_bytes) / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(smem)[i] = __ldg(reinterpret_cast<const uint4*>(a.tpack) + i);
  if (tid == 0) {
    mbar_init(&a_full, 4);
    mbar_init(&d_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 4) tc::tmem_alloc(&s_tmem, 512);
  fence_proxy_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = s_tmem;
  uint32_t it = 0;
  if (warp < 4) {
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const int rl = warp * 32 + lane;                 // row within the tile
    const int bl = rl / a.D, d = rl - bl * a.D;      // sample within the tile, coordinate
    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const long long r = tile * 128 + rl;
      const bool valid = r < a.rows;
      const long long b = valid ? r / a.D : 0;
      const uint32_t stg = it & 1;
      const unsigned char* sp = smem + tab_bytes + stg * stage_bytes;
      const unsigned short* prow = reinterpret_cast<const unsigned short*>(sp) + (long long)bl * a.Hp * a.D + d;
      const unsigned short* xrow = reinterpret_cast<const unsigned short*>(sp + pre_tile) + (long long)bl * MF * a.D + d;
      mbar_wait(&s_full[stg], (it >> 1) & 1);
      // ---- A operands: x0 row (K = 32) and pre row (K = 16*nkh), from shared memory ------------
      {
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const uint32_t lo = (valid && 2 * q < MF) ? xrow[(2 * q) * a.D] : 0u;
          const uint32_t hi = (valid && 2 * q + 1 < MF) ? xrow[(2 * q + 1) * a.D] : 0u;
          w[q] = lo | (hi << 16);
        }
        tc::st8(tmem + lane_base + kLtAx, w);
        tc::st8(tmem + lane_base + kLtAx + 8, w + 8);
      }
      for (int s0 = 0; s0 < a.nkh; ++s0) {
        uint32_t w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int h = 16 * s0 + 2 * q;
          const uint32_t lo = (valid && h < a.Hp) ? prow[h * a.D] : 0u;
          const uint32_t hi = (valid && h + 1 < a.Hp) ? prow[(h + 1) * a.D] : 0u;
          w[q] = lo | (hi << 16);
        }
        tc::st8(tmem + lane_base + kLtAp + 8 * s0, w);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[stg]);                  // our reads of this stage are done
      tc::wait_st();
      tc::fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full);
      const float g = valid ? a.gpool[b * a.gstride + a.gcol + d] : 0.f;
      const float gprev = valid ? a.gpool[b * a.gstride + a.gcol_prev + d] : 0.f;
      float* dxr = a.dx0 + b * a.dx_sb + d;
      float old[MF];
#pragma unroll
      for (int i = 0; i < MF; ++i) old[i] = valid ? dxr[(long long)i * a.D] : 0.f;
      // ---- epilogue ------------------------------------------------------------------------
      mbar_wait(&d_full, it & 1);
      tc::fence_after();
      for (int o0 = 0; o0 < a.N8p * 8; o0 += 16) {
        uint32_t v[16];
        tc::ld16(tmem + lane_base + kLtD1 + o0, v);
        tc::wait_ld();
        uint32_t pk[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float v0 = (o0 + 2 * q < a.Hp) ? fmaf(g, __uint_as_float(v[2 * q]), gprev) : 0.f;
          const float v1 = (o0 + 2 * q + 1 < a.Hp) ? fmaf(g, __uint_as_float(v[2 * q + 1]), gprev) : 0.f;
          pk[q] = tc::pack_bf16(v0, v1);
        }
        if (valid) {
          unsigned char* dst = a.dz_prev + ((r >> 3) * a.N8p + (o0 >> 3)) * 128 + (r & 7) * 16;
          *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          if ((o0 >> 3) + 1 < a.N8p) *reinterpret_cast<uint4*>(dst + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      {
        uint32_t v[32];
        tc::ld32(tmem + lane_base + kLtD2, v);
        tc::wait_ld();
        if (valid) {
#pragma unroll
          for (int i = 0; i < MF; ++i) dxr[(long long)i * a.D] = fmaf(g, __uint_as_float(v[i]), old[i]);
        }
      }
      tc::fence_before();
    }
  } else if (warp == 4) {
    const bool leader = tc::elect_one();
    const uint32_t idesc1 = tc::idesc_bf16(128, a.Nh8 * 8, 0, 0), idesc2 = tc::idesc_bf16(128, 32, 0, 0);
    const uint64_t bd1 = tc::smem_desc(smem_u32(smem), (uint32_t)a.Nh8 * 128, 128);
    const uint64_t bd2 = tc::smem_desc(smem_u32(smem + th_bytes), 512, 128);
    const uint32_t adv1 = (2u * a.Nh8 * 128u) >> 4;
    for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      mbar_wait(&a_full, it & 1);
      tc::fence_after();
      if (leader) {
        tc::mma_ts(tmem + kLtD1, tmem + kLtAx, bd1, idesc1, 0u);
        tc::mma_ts(tmem + kLtD1, tmem + kLtAx + 8, bd1 + (uint64_t)adv1, idesc1, 1u);
#pragma unroll
        for (int s = 0; s < 13; ++s)
          if (s < a.nkh) tc::mma_ts(tmem + kLtD2, tmem + kLtAp + 8 * s, bd2 + (uint64_t)(s * 64), idesc2, s > 0 ? 1u : 0u);
        tc::commit(&d_full);
      }
      __syncwarp();
    }
  } else {
    if (lane == 0) {
      const long long spt = 128 / a.D;                       // samples per tile
      const long long n_samples = a.rows / a.D;
      for (long long tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t stg = it & 1;
        const long long b0 = tile * spt;
        const long long nb = min(spt, n_samples - b0);
        unsigned char* sp = smem + tab_bytes + stg * stage_bytes;
        mbar_wait(&s_empty[stg], ((it >> 1) & 1) ^ 1);
        const uint32_t pb = (uint32_t)(nb * a.Hp * a.D * 2), xb = (uint32_t)(nb * MF * a.D * 2);
        mbar_expect_tx(&s_full[stg], pb + xb);
        bulk_g2s(sp, a.pre + b0 * a.Hp * a.D, pb, &s_full[stg]);
        bulk_g2s(sp + pre_tile, a.x0b + b0 * MF * a.D, xb, &s_full[stg]);
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 4) tc::tmem_dealloc(tmem, 512);
}

// dW[c,o] = sum_slices part[s][c][0] for every o (c < C);  dbias[o] = sum_slices part[s][C][0]
__global__ void __launch_bounds__(256)
cin_dw_reduce_bcast_kernel(const float* __restrict__ part, int n_slices, long long slice_stride, int Npad,
                           int C, int N, float* __restrict__ dW, float* __restrict__ dbias) {
  const long long total = (long long)(C + 1) * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / N;
    const int o = (int)(idx - c * N);
    const float* p = part + c * Npad;
    float acc = 0.f;
    for (int s = 0; s < n_slices; ++s) acc += p[s * slice_stride];
    if (c < C) dW[idx] = acc;
    else dbias[o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// dW: stream kernel, lanes = c
// ---------------------------------------------------------------------------------------------
struct DwArgs {
  const unsigned short* pre;   // [B,Hp,D] bf16
  const unsigned short* x0b;   // [B,m,D] bf16
  const unsigned short* x0h;   // [B,2,m,8] bf16 (D == 16 only): staged path
  const unsigned char* dz;     // blocked, kblk bytes per k-step (16 rows)
  float* part;                 // [n_slices][n_cp*256][Npad]
  int D, Hp, N8, C, n_cp, n_slices;
  uint32_t kblk;
  long long nks_total, ks_per_slice;
};

// kStaged (D == 16): the loader thread also bulk-copies, per k-step (= one sample b), the slab
// pre[b, h_lo..h_hi, :] (<= 11 rows x 32 B) and x0h[b] (832 B) into the B stage; the row warps read
// their two 32-byte vectors from shared memory (29-cycle latency) instead of chasing L2 with
// per-lane global loads, and release the stage together with the MMA (b_empty count 9).
constexpr int kDwNH = 11;
template <int MF>
constexpr uint32_t dw_stage_bytes(uint32_t kblk) { return kG * (kblk + kDwNH * 32u + MF * 32u); }

template <int MF, bool kStaged>
__global__ void __launch_bounds__(kTcThreads, 1) cin_dw_tc_kernel(const DwArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Barriers bars;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stream_init(bars, tid, warp, kStaged ? 1 + kProdWarps : 1);
  const uint32_t tmem = bars.tmem_base;
  const uint32_t stage_stride = kStaged ? dw_stage_bytes<MF>(a.kblk) : kG * a.kblk;
  const uint32_t offP = kG * a.kblk, offX = offP + kG * kDwNH * 32u;

  const int cp = blockIdx.x % a.n_cp;
  const int slice = blockIdx.x / a.n_cp;
  const long long ks0 = slice * a.ks_per_slice;
  const long long nk_ll = min(a.ks_per_slice, a.nks_total - ks0);
  const int nk = nk_ll > 0 ? (int)nk_ll : 0;
  const int Npad = a.N8 * 8;

  if (warp < kProdWarps) {
    const int sub = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t colD = sub ? kColD1 : kColD0;
    const uint32_t colA = sub ? kColA1 : kColA0;
    const int c = cp * (128 * kSub) + sub * 128 + (warp & 3) * 32 + lane;
    const int mode = c < a.C ? 2 : (c == a.C ? 1 : 0);          // product / ones (dbias) / zero
    const int h = mode == 2 ? c / MF : 0, i = mode == 2 ? c % MF : 0;
    const int ksD = a.D / 16;                                   // k-steps per sample
    if (nk > 0 && kStaged) {
      SlotWriter sw;
      const int n_groups = (nk + kG - 1) / kG;
      const int h_lo = (cp * (128 * kSub)) / MF;
      const uint32_t myP = offP + (uint32_t)(h - h_lo) * 32u, myX = offX + (uint32_t)i * 16u;
      uint32_t bs = 0, bph = 0;
      for (int g = 0; g < n_groups; ++g) {
        const int gk = min(kG, nk - g * kG);
        mbar_wait(&bars.b_full[bs], bph);
        const unsigned char* st = smem + bs * stage_stride;
        uint32_t w[kG][8];
#pragma unroll
        for (int u = 0; u < kG; ++u) {
          if (mode == 2) {
            const uint4 p0 = *reinterpret_cast<const uint4*>(st + myP + u * (kDwNH * 32));
            const uint4 p1 = *reinterpret_cast<const uint4*>(st + myP + u * (kDwNH * 32) + 16);
            const uint4 x0 = *reinterpret_cast<const uint4*>(st + myX + u * (MF * 32));
            const uint4 x1 = *reinterpret_cast<const uint4*>(st + myX + u * (MF * 32) + MF * 16);
            w[u][0] = hmul2_bf16(p0.x, x0.x); w[u][1] = hmul2_bf16(p0.y, x0.y);
            w[u][2] = hmul2_bf16(p0.z, x0.z); w[u][3] = hmul2_bf16(p0.w, x0.w);
            w[u][4] = hmul2_bf16(p1.x, x1.x); w[u][5] = hmul2_bf16(p1.y, x1.y);
            w[u][6] = hmul2_bf16(p1.z, x1.z); w[u][7] = hmul2_bf16(p1.w, x1.w);
          } else {
            const uint32_t v = mode == 1 ? 0x3F803F80u : 0u;
#pragma unroll
            for (int z = 0; z < 8; ++z) w[u][z] = v;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.b_empty[bs]);               // our reads of this stage are done
        if (++bs == kS) { bs = 0; bph ^= 1; }
#pragma unroll
        for (int u = 0; u < kG; ++u)
          if (u < gk) sw.put(bars, sub, tmem + lane_base + colA, w[u], g * kG + u == nk - 1, lane);
      }
    } else if (nk > 0) {
      SlotWriter sw;
      const int n_groups = (nk + kG - 1) / kG;
      // running element offsets of the NEXT k-step to fetch: sample-major, 16 coordinates per k-step
      long long bq = ks0 / ksD;
      int dq = (int)(ks0 - bq * ksD);
      const unsigned short* pp = a.pre + (bq * a.Hp + h) * (long long)a.D + dq * 16;
      const unsigned short* xp = a.x0b + (bq * MF + i) * (long long)a.D + dq * 16;
      const long long p_wrap = (long long)a.Hp * a.D - a.D + 16;   // last k-step of a sample -> first of the next
      const long long x_wrap = (long long)MF * a.D - a.D + 16;
      int fetched = 0;                                             // k-steps fetched so far
      constexpr int kPF = 2;                                       // groups in flight ahead of the compute
      uint4 nb[kPF][kG][4];
      auto load_group = [&](int slot) {
#pragma unroll
        for (int u = 0; u < kG; ++u) {
          if (fetched < nk) {
            const uint4* p4 = reinterpret_cast<const uint4*>(pp);
            const uint4* x4 = reinterpret_cast<const uint4*>(xp);
            nb[slot][u][0] = __ldg(p4); nb[slot][u][1] = __ldg(p4 + 1);
            nb[slot][u][2] = __ldg(x4); nb[slot][u][3] = __ldg(x4 + 1);
            ++fetched;
            if (++dq == ksD) { dq = 0; pp += p_wrap; xp += x_wrap; }
            else { pp += 16; xp += 16; }
          }
        }
      };
#pragma unroll
      for (int q = 0; q < kPF; ++q) load_group(q);
      for (int g0 = 0; g0 < n_groups; g0 += kPF) {
#pragma unroll
        for (int q = 0; q < kPF; ++q) {
          const int g = g0 + q;
          if (g < n_groups) {
            uint32_t w[kG][8];
#pragma unroll
            for (int u = 0; u < kG; ++u) {
              if (mode == 2) {
                const uint4 p0 = nb[q][u][0], p1 = nb[q][u][1], x0 = nb[q][u][2], x1 = nb[q][u][3];
                w[u][0] = hmul2_bf16(p0.x, x0.x); w[u][1] = hmul2_bf16(p0.y, x0.y);
                w[u][2] = hmul2_bf16(p0.z, x0.z); w[u][3] = hmul2_bf16(p0.w, x0.w);
                w[u][4] = hmul2_bf16(p1.x, x1.x); w[u][5] = hmul2_bf16(p1.y, x1.y);
                w[u][6] = hmul2_bf16(p1.z, x1.z); w[u][7] = hmul2_bf16(p1.w, x1.w);
              } else {
                const uint32_t v = mode == 1 ? 0x3F803F80u : 0u;
#pragma unroll
                for (int z = 0; z < 8; ++z) w[u][z] = v;
              }
            }
            load_group(q);                                          // refill this slot: group g + kPF
#pragma unroll
            for (int u = 0; u < kG; ++u) {
              const int kl = g * kG + u;
              if (kl < nk) sw.put(bars, sub, tmem + lane_base + colA, w[u], kl == nk - 1, lane);
            }
          }
        }
      }
    }
    if (nk > 0) {
      // ---- epilogue: this lane's accumulator row -> partial buffer --------------------------
      mbar_wait(&bars.d_full, 0);
      tc::fence_after();
      float* prow = a.part + ((long long)slice * a.n_cp * (128 * kSub) + c) * Npad;
      for (int o0 = 0; o0 < Npad; o0 += 8) {
        uint32_t v[8];
        tc::ld8(tmem + lane_base + colD + o0, v);
        tc::wait_ld();
        *reinterpret_cast<uint4*>(prow + o0) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(prow + o0 + 4) = make_uint4(v[4], v[5], v[6], v[7]);
      }
    } else {
      // empty slice: contribute zeros so that the reduction can read every slice
      float* prow = a.part + ((long long)slice * a.n_cp * (128 * kSub) + c) * Npad;
      for (int o0 = 0; o0 < Npad; o0 += 4) *reinterpret_cast<uint4*>(prow + o0) = make_uint4(0, 0, 0, 0);
    }
  } else if (warp == kProdWarps) {
    if (nk > 0)
      stream_mma_role(bars, smem, tmem, 1, nk, a.kblk, tc::idesc_bf16(128, Npad, 0, 1),
                      (uint32_t)a.N8 * 128, 128, stage_stride);
  } else {
    if (lane == 0 && nk > 0) {
      if (!kStaged) {
        stream_loader_role(bars, smem, 1, nk, a.kblk, [&](long long) { return a.dz + (size_t)ks0 * a.kblk; });
      } else {
        const int n_groups = (nk + kG - 1) / kG;
        const int h_lo = (cp * (128 * kSub)) / MF;
        const int h_hi = min(a.Hp - 1, (cp * (128 * kSub) + 128 * kSub - 1) / MF);
        const uint32_t pbytes = h_hi >= h_lo ? (uint32_t)(h_hi - h_lo + 1) * 32u : 0u;   // 0: only the bias lane
        uint32_t bs = 0, bph = 0;
        for (int g = 0; g < n_groups; ++g) {
          const int gk = min(kG, nk - g * kG);
          const long long k0 = ks0 + (long long)g * kG;            // == first sample of the group (D == 16)
          mbar_wait(&bars.b_empty[bs], bph ^ 1);
          unsigned char* st = smem + bs * stage_stride;
          mbar_expect_tx(&bars.b_full[bs], gk * (a.kblk + pbytes + MF * 32u));
          bulk_g2s(st, a.dz + (size_t)k0 * a.kblk, gk * a.kblk, &bars.b_full[bs]);
          bulk_g2s(st + offX, a.x0h + (size_t)k0 * (MF * 16), gk * MF * 32u, &bars.b_full[bs]);
          for (int u = 0; u < gk && pbytes; ++u)
            bulk_g2s(st + offP + u * (kDwNH * 32), a.pre + ((size_t)(k0 + u) * a.Hp + h_lo) * 16, pbytes,
                     &bars.b_full[bs]);
          if (++bs == kS) { bs = 0; bph ^= 1; }
        }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == kProdWarps) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// dW, D == 16, second generation: 16 row warps in two sets.  Set s fills the A slot s of its
// sub-tile for the groups g = s, s+2, ... (with kNS == 2 a set always owns the same slot), so the
// chain  stage wait -> LDS -> HMUL2 -> slot wait -> tcgen05.st -> wait::st -> arrive  of one group
// overlaps with the other set's chain instead of serialising every group of a lane quarter.
// ---------------------------------------------------------------------------------------------
constexpr int kDw2Prod = 16;
constexpr int kDw2Threads = 32 * (kDw2Prod + 2);

template <int MF>
__global__ void __launch_bounds__(kDw2Threads, 1) cin_dw2_tc_kernel(const DwArgs a) {
  static_assert(kNS == 2 && kSub == 2, "two producer sets <-> two A slots");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Barriers bars;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kS; ++i) { mbar_init(&bars.b_full[i], 1); mbar_init(&bars.b_empty[i], 1 + 8); }
    for (int s = 0; s < kSub; ++s)
      for (int i = 0; i < kNS; ++i) { mbar_init(&bars.a_full[s][i], 4); mbar_init(&bars.a_empty[s][i], 1); }
    mbar_init(&bars.d_full, 1);
    fence_mbar_init();
  }
  if (warp == kDw2Prod) tc::tmem_alloc(&bars.tmem_base, 512);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = bars.tmem_base;
  const uint32_t stage_stride = dw_stage_bytes<MF>(a.kblk);
  const uint32_t offP = kG * a.kblk, offX = offP + kG * kDwNH * 32u;

  const int cp = blockIdx.x % a.n_cp;
  const int slice = blockIdx.x / a.n_cp;
  const long long ks0 = slice * a.ks_per_slice;
  const long long nk_ll = min(a.ks_per_slice, a.nks_total - ks0);
  const int nk = nk_ll > 0 ? (int)nk_ll : 0;
  const int Npad = a.N8 * 8;
  const int n_groups = (nk + kG - 1) / kG;

  if (warp < kDw2Prod) {
    const int set = warp >> 3, w8 = warp & 7, sub = w8 >> 2, quarter = w8 & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t colD = sub ? kColD1 : kColD0;
    const uint32_t colA = (sub ? kColA1 : kColA0) + set * (8 * kG);
    const int c = cp * 256 + sub * 128 + quarter * 32 + lane;
    const int mode = c < a.C ? 2 : (c == a.C ? 1 : 0);          // product / ones (dbias) / zero
    const int h = mode == 2 ? c / MF : 0, i = mode == 2 ? c % MF : 0;
    const int h_lo = (cp * 256) / MF;
    const uint32_t myP = offP + (uint32_t)(h - h_lo) * 32u, myX = offX + (uint32_t)i * 16u;
    float* prow = a.part + ((long long)slice * a.n_cp * 256 + c) * Npad;
    if (nk > 0) {
      for (int g = set; g < n_groups; g += 2) {
        const int gk = min(kG, nk - g * kG);
        const uint32_t bs = g % kS, bph = (g / kS) & 1, use = g >> 1;
        mbar_wait(&bars.b_full[bs], bph);
        const unsigned char* st = smem + bs * stage_stride;
        uint32_t w[kG][8];
#pragma unroll
        for (int u = 0; u < kG; ++u) {
          if (mode == 2) {
            const uint4 p0 = *reinterpret_cast<const uint4*>(st + myP + u * (kDwNH * 32));
            const uint4 p1 = *reinterpret_cast<const uint4*>(st + myP + u * (kDwNH * 32) + 16);
            const uint4 x0 = *reinterpret_cast<const uint4*>(st + myX + u * (MF * 32));
            const uint4 x1 = *reinterpret_cast<const uint4*>(st + myX + u * (MF * 32) + MF * 16);
            w[u][0] = hmul2_bf16(p0.x, x0.x); w[u][1] = hmul2_bf16(p0.y, x0.y);
            w[u][2] = hmul2_bf16(p0.z, x0.z); w[u][3] = hmul2_bf16(p0.w, x0.w);
            w[u][4] = hmul2_bf16(p1.x, x1.x); w[u][5] = hmul2_bf16(p1.y, x1.y);
            w[u][6] = hmul2_bf16(p1.z, x1.z); w[u][7] = hmul2_bf16(p1.w, x1.w);
          } else {
            const uint32_t v = mode == 1 ? 0x3F803F80u : 0u;
#pragma unroll
            for (int z = 0; z < 8; ++z) w[u][z] = v;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.b_empty[bs]);               // our reads of this stage are done
        mbar_wait(&bars.a_empty[sub][set], (use & 1) ^ 1);
        tc::fence_after();
#pragma unroll
        for (int u = 0; u < kG; ++u)
          if (u < gk) tc::st8(tmem + lane_base + colA + 8 * u, w[u]);
        tc::wait_st();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.a_full[sub][set]);
      }
      // ---- epilogue: the two sets split this lane's accumulator row by 8-column chunks ---------
      mbar_wait(&bars.d_full, 0);
      tc::fence_after();
      for (int o0 = set * 8; o0 < Npad; o0 += 16) {
        uint32_t v[8];
        tc::ld8(tmem + lane_base + colD + o0, v);
        tc::wait_ld();
        *reinterpret_cast<uint4*>(prow + o0) = make_uint4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<uint4*>(prow + o0 + 4) = make_uint4(v[4], v[5], v[6], v[7]);
      }
    } else {
      for (int o0 = set * 8; o0 < Npad; o0 += 16) {
        *reinterpret_cast<uint4*>(prow + o0) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(prow + o0 + 4) = make_uint4(0, 0, 0, 0);
      }
    }
  } else if (warp == kDw2Prod) {
    if (nk > 0)
      stream_mma_role(bars, smem, tmem, 1, nk, a.kblk, tc::idesc_bf16(128, Npad, 0, 1),
                      (uint32_t)a.N8 * 128, 128, stage_stride);
  } else {
    if (lane == 0 && nk > 0) {
      const int h_lo = (cp * 256) / MF;
      const int h_hi = min(a.Hp - 1, (cp * 256 + 255) / MF);
      const uint32_t pbytes = h_hi >= h_lo ? (uint32_t)(h_hi - h_lo + 1) * 32u : 0u;   // 0: only the bias lane
      uint32_t bs = 0, bph = 0;
      for (int g = 0; g < n_groups; ++g) {
        const int gk = min(kG, nk - g * kG);
        const long long k0 = ks0 + (long long)g * kG;              // == first sample of the group (D == 16)
        mbar_wait(&bars.b_empty[bs], bph ^ 1);
        unsigned char* st = smem + bs * stage_stride;
        mbar_expect_tx(&bars.b_full[bs], gk * (a.kblk + pbytes + MF * 32u));
        bulk_g2s(st, a.dz + (size_t)k0 * a.kblk, gk * a.kblk, &bars.b_full[bs]);
        bulk_g2s(st + offX, a.x0h + (size_t)k0 * (MF * 16), gk * MF * 32u, &bars.b_full[bs]);
        for (int u = 0; u < gk && pbytes; ++u)
          bulk_g2s(st + offP + u * (kDwNH * 32), a.pre + ((size_t)(k0 + u) * a.Hp + h_lo) * 16, pbytes,
                   &bars.b_full[bs]);
        if (++bs == kS) { bs = 0; bph ^= 1; }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == kDw2Prod) tc::tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------
// dA: chunked GEMM with contracting epilogue
// ---------------------------------------------------------------------------------------------
constexpr int kDaS = 4;                       // weight-chunk stages
constexpr uint32_t kDaColA0 = 0, kDaColD0 = 104, kDaColA1 = 256, kDaColD1 = 360;

struct DaBarriers {
  uint64_t b_full[kDaS], b_empty[kDaS];
  uint64_t a_full[2], a_empty[2], d_full[2], d_empty[2];
  uint32_t tmem_base;
};

struct DaArgs {
  const unsigned char* dz;       // blocked dZ_l
  const unsigned char* wpack;    // n_chunks chunks
  const unsigned short* x0b;     // [B,m,D] bf16
  const unsigned short* pre;     // [B,Hp,D] bf16 (layer 0: == x0b)
  float* dx0;                    // [B,m,D] fp32, accumulated; batch stride dx_sb
  long long dx_sb;
  float* dpre0;                  // [B,m,D] fp32 scratch: layer 0's dpre (pre == x0), folded into dx0 per tile
  unsigned char* dz_prev;        // blocked dZ_{l-1} (N8p column groups) or nullptr on layer 0
  const float* gpool;            // d_pooled [B, gstride]
  int gstride, gcol_prev;
  long long rows, n_pairs;
  int D, Hp, N8, N8p, n_chunks, nkA;
  uint32_t chunk_bytes;
};

__device__ __forceinline__ void ffma2(float2& acc, const float2 a, const float2 b) {
#if KON_NO_FFMA2
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.y = fmaf(a.y, b.y, acc.y);
#else
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(rc));
#endif
}

template <int MF>
__global__ void __launch_bounds__(320, 1) cin_da_tc_kernel(const DaArgs a) {
  constexpr int CH = kDaChunkH * MF;          // accumulator columns per chunk (104)
  static_assert(CH % 8 == 0 && CH <= 152 && MF % 2 == 0, "chunk shape");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ DaBarriers bars;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kDaS; ++i) { mbar_init(&bars.b_full[i], 1); mbar_init(&bars.b_empty[i], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.a_full[s], 4); mbar_init(&bars.a_empty[s], 1);
      mbar_init(&bars.d_full[s], 1); mbar_init(&bars.d_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 8) tc::tmem_alloc(&bars.tmem_base, 512);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = bars.tmem_base;
  const long long n_items = a.n_pairs > blockIdx.x ? (a.n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t kstep_bytes = 2u * (CH / 8) * 128u;

  if (warp < 8) {
    const int sub = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t colA = sub ? kDaColA1 : kDaColA0;
    const uint32_t colD = sub ? kDaColD1 : kDaColD0;
    uint32_t d_use = 0;
    uint32_t tile_it = 0;
    // dx0 of a tile is folded into global memory (read-modify-write, two dependent DRAM round trips) AFTER the next
    // tile's A operand has been staged and handed to the MMA warp: the tensor core then works on the new tile while
    // the old tile's rows are updated, instead of idling behind them at every tile boundary.
#ifndef KON_DA_DEFER_RMW
#define KON_DA_DEFER_RMW 1
#endif
    float2 dx2_prev[MF / 2];
    float* dxr_prev = nullptr;
    const float* dq_prev = nullptr;
    auto fold_dx0 = [&](float* dxr, const float* dq, const float2* dxv) {
      float old[MF];
#pragma unroll
      for (int i = 0; i < MF; ++i) old[i] = dxr[(long long)i * a.D];          // all loads first
      if (!a.dz_prev) {                                                       // layer 0: + dpre (written by this thread)
        float t[MF];
#pragma unroll
        for (int i = 0; i < MF; ++i) t[i] = dq[(long long)i * a.D];
#pragma unroll
        for (int i = 0; i < MF; ++i) old[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < MF / 2; ++i) {
        dxr[(long long)(2 * i) * a.D] = old[2 * i] + dxv[i].x;
        dxr[(long long)(2 * i + 1) * a.D] = old[2 * i + 1] + dxv[i].y;
      }
    };
    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x, ++tile_it) {
      const long long r = pair * 256 + sub * 128 + (warp & 3) * 32 + lane;
      const bool valid = r < a.rows;
      const long long b = valid ? r / a.D : 0;
      const int d = valid ? (int)(r - b * a.D) : 0;
      // ---- A operand: this row of dZ_l, bf16, into TMEM (once per tile) ----------------------
      mbar_wait(&bars.a_empty[sub], (tile_it & 1) ^ 1);
      tc::fence_after();
      {
        // all 16-byte units of the row are requested before the first one is used (one exposed
        // HBM latency per tile instead of one per k-step)
        const long long rc = valid ? r : 0;
        const uint4* zr = reinterpret_cast<const uint4*>(a.dz + ((rc >> 3) * a.N8) * 128 + (rc & 7) * 16);
        uint4 zq[26];
#pragma unroll
        for (int q = 0; q < 26; ++q)
          zq[q] = (valid && q < a.N8) ? __ldg(zr + q * 8) : make_uint4(0, 0, 0, 0);   // column group q (128 B apart)
#pragma unroll
        for (int s = 0; s < 13; ++s) {
          if (s < a.nkA) {
            const uint32_t w[8] = {zq[2 * s].x, zq[2 * s].y, zq[2 * s].z, zq[2 * s].w,
                                   zq[2 * s + 1].x, zq[2 * s + 1].y, zq[2 * s + 1].z, zq[2 * s + 1].w};
            tc::st8(tmem + lane_base + colA + 8 * s, w);
          }
        }
        tc::wait_st();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.a_full[sub]);
      }
      if (KON_DA_DEFER_RMW && dxr_prev) {          // previous tile's dx0, now behind this tile's first MMAs
        fold_dx0(dxr_prev, dq_prev, dx2_prev);
        dxr_prev = nullptr;
      }
      // ---- per-row constants -------------------------------------------------------------------
      const unsigned short* xrow = a.x0b + b * (long long)MF * a.D + d;      // x0[b,i,d] = xrow[i*D]
      const unsigned short* prow = a.pre + b * (long long)a.Hp * a.D + d;    // pre[b,h,d] = prow[h*D]
      float2 x2[MF / 2], dx2[MF / 2];
#pragma unroll
      for (int i = 0; i < MF / 2; ++i) {
        x2[i].x = bf16_to_f32(__ldg(xrow + (long long)(2 * i) * a.D));
        x2[i].y = bf16_to_f32(__ldg(xrow + (long long)(2 * i + 1) * a.D));
        dx2[i] = make_float2(0.f, 0.f);
      }
      const float gp = (a.dz_prev && valid) ? a.gpool[b * a.gstride + a.gcol_prev + d] : 0.f;
      unsigned short praw[kDaChunkH];
#pragma unroll
      for (int q = 0; q < kDaChunkH; ++q) praw[q] = __ldg(prow + (long long)min(q, a.Hp - 1) * a.D);

      for (int j = 0; j < a.n_chunks; ++j, ++d_use) {
        float2 pb[kDaChunkH], dp[kDaChunkH];
#pragma unroll
        for (int q = 0; q < kDaChunkH; ++q) {
          const int h = j * kDaChunkH + q;
          const float p = (h < a.Hp) ? bf16_to_f32(praw[q]) : 0.f;
          pb[q] = make_float2(p, p);
          dp[q] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < kDaChunkH; ++q)
          praw[q] = __ldg(prow + (long long)min((j + 1) * kDaChunkH + q, a.Hp - 1) * a.D);

        mbar_wait(&bars.d_full[sub], d_use & 1);
        tc::fence_after();
        // 104 accumulator columns: column q -> (hh, i) = (q / MF, q % MF); pairs never straddle an h
#pragma unroll
        for (int q0 = 0; q0 < CH; q0 += 32) {
          constexpr int dummy = 0;
          (void)dummy;
          uint32_t v[32];
          if (q0 + 32 <= CH) {
            tc::ld32(tmem + lane_base + colD + q0, v);
          } else {
            tc::ld8(tmem + lane_base + colD + q0, v);            // CH - q0 == 8 for MF == 26
          }
          tc::wait_ld();
          if (q0 + 32 > CH) {                                     // all loads of this chunk are done
            tc::fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.d_empty[sub]);
          }
#pragma unroll
          for (int q = 0; q < 32; q += 2) {
            if (q0 + q < CH) {
              const int col = q0 + q;
              const int hh = col / MF, ii = (col % MF) / 2;
              const float2 val = make_float2(__uint_as_float(v[q]), __uint_as_float(v[q + 1]));
              ffma2(dp[hh], val, x2[ii]);
              ffma2(dx2[ii], val, pb[hh]);
            }
          }
        }
        // ---- dpre of this chunk's 4 feature maps -------------------------------------------------
        if (valid) {
          if (a.dz_prev) {
            uint32_t w2[2];
#pragma unroll
            for (int q = 0; q < kDaChunkH; q += 2) {
              const int h = j * kDaChunkH + q;
              const float v0 = h < a.Hp ? dp[q].x + dp[q].y + gp : 0.f;
              const float v1 = h + 1 < a.Hp ? dp[q + 1].x + dp[q + 1].y + gp : 0.f;
              w2[q / 2] = tc::pack_bf16(v0, v1);
            }
            const int h0 = j * kDaChunkH;
            if (h0 < a.N8p * 8) {
              unsigned char* dst = a.dz_prev + ((r >> 3) * a.N8p + (h0 >> 3)) * 128 + (r & 7) * 16 + (h0 & 7) * 2;
              *reinterpret_cast<uint2*>(dst) = make_uint2(w2[0], w2[1]);
            }
          } else {
            // layer 0: pre == x0, so dpre lands in dx0[b,h,d]
#pragma unroll
            for (int q = 0; q < kDaChunkH; ++q) {
              const int h = j * kDaChunkH + q;
              if (h < a.Hp)      // plain store now, one batched read-modify-write at the end of the tile
                a.dpre0[(b * MF + h) * (long long)a.D + d] = dp[q].x + dp[q].y;
            }
          }
        }
      }
      if (valid) {
        float* dxr = a.dx0 + b * a.dx_sb + d;
        const float* dq = a.dpre0 + b * (long long)MF * a.D + d;
        if (KON_DA_DEFER_RMW) {
#pragma unroll
          for (int i = 0; i < MF / 2; ++i) dx2_prev[i] = dx2[i];
          dxr_prev = dxr;
          dq_prev = dq;
        } else {
          fold_dx0(dxr, dq, dx2);
        }
      }
    }
    if (KON_DA_DEFER_RMW && dxr_prev) fold_dx0(dxr_prev, dq_prev, dx2_prev);
  } else if (warp == 8) {
    // MMA role: warp-uniform control flow, one elected lane issues (see stream_mma_role)
    {
      const uint32_t idesc = tc::idesc_bf16(128, CH, 0, 0);
      const uint32_t lbo = (CH / 8) * 128, sbo = 128;
      const uint32_t kadv = kstep_bytes >> 4;
      uint32_t bs = 0, bph = 0, d_use = 0;
      const bool leader = tc::elect_one();
      for (long long it = 0; it < n_items; ++it) {
        mbar_wait(&bars.a_full[0], it & 1);
        mbar_wait(&bars.a_full[1], it & 1);
        tc::fence_after();
        for (int j = 0; j < a.n_chunks; ++j, ++d_use) {
          mbar_wait(&bars.b_full[bs], bph);
          tc::fence_after();
          const uint64_t bdesc0 = tc::smem_desc(smem_u32(smem + bs * a.chunk_bytes), lbo, sbo);
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            mbar_wait(&bars.d_empty[sub], (d_use & 1) ^ 1);
            tc::fence_after();
            const uint32_t dcol = tmem + (sub ? kDaColD1 : kDaColD0);
            const uint32_t acol = tmem + (sub ? kDaColA1 : kDaColA0);
            if (leader) {
#pragma unroll
              for (int ks = 0; ks < 13; ++ks)
                if (ks < a.nkA)
                  tc::mma_ts(dcol, acol + 8 * ks, bdesc0 + (uint64_t)(ks * kadv), idesc, ks > 0 ? 1u : 0u);
              tc::commit(&bars.d_full[sub]);
            }
            __syncwarp();
          }
          if (leader) tc::commit(&bars.b_empty[bs]);
          __syncwarp();
          if (++bs == kDaS) { bs = 0; bph ^= 1; }
        }
        if (leader) {
          tc::commit(&bars.a_empty[0]);
          tc::commit(&bars.a_empty[1]);
        }
        __syncwarp();
      }
    }
  } else {
    if (lane == 0) {
      uint32_t bs = 0, bph = 0;
      for (long long it = 0; it < n_items; ++it) {
        for (int j = 0; j < a.n_chunks; ++j) {
          mbar_wait(&bars.b_empty[bs], bph ^ 1);
          mbar_expect_tx(&bars.b_full[bs], a.chunk_bytes);
          bulk_g2s(smem + bs * a.chunk_bytes, a.wpack + (size_t)j * a.chunk_bytes, a.chunk_bytes, &bars.b_full[bs]);
          if (++bs == kDaS) { bs = 0; bph ^= 1; }
        }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 8) tc::tmem_dealloc(tmem, 512);
}

int grid_of(long long total, int sms, int mult = 8) {
  return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * mult));
}

}  // namespace

int cin_tc_bwd(const float* x0, long long x0_sb, const float* const* w, const float* const* bias, int nl,
               const int32_t* hs, int64_t B, int m, int D, const float* d_pooled, const void* saved,
               float* dx0, long long dx0_sb, float* const* dw, float* const* dbias, void* workspace, int sms,
               cudaStream_t st) {
  TcLayout L;
  KON_TRY(tc_check_layout(tc_layout(B, m, D, hs, nl, sms, &L), m, D));
  KON_REQUIRE(((uintptr_t)workspace & 255u) == 0 && ((uintptr_t)saved & 255u) == 0, KON_EINVAL,
              "saved / workspace must be 256-B aligned");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const unsigned char* sv = static_cast<const unsigned char*>(saved);
  const long long rows = B * D;
  const size_t smem_stream = stream_smem_bytes();
  const size_t smem_dw = (size_t)kS * dw_stage_bytes<26>(2u * ((kMaxN + 7) / 8) * 128u);
  const size_t smem_da = (size_t)kDaS * 13 * 2 * (kDaChunkH * 26 / 8) * 128;
  static bool attr_done = false;
  if (!attr_done) {
    KON_CUDA(cudaFuncSetAttribute(cin_dw_tc_kernel<26, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
    KON_CUDA(cudaFuncSetAttribute(cin_dw_tc_kernel<26, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dw));
    KON_CUDA(cudaFuncSetAttribute(cin_dw2_tc_kernel<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dw));
    KON_CUDA(cudaFuncSetAttribute(cin_last_da_tc_kernel<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    KON_CUDA(cudaFuncSetAttribute(cin_da_tc_kernel<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_da));
    attr_done = true;
  }
  unsigned short* x0b = reinterpret_cast<unsigned short*>(ws + L.x0b_off);
  const long long nx = B * m * D;
  unsigned short* x0h = reinterpret_cast<unsigned short*>(ws + L.x0h_off);
  KON_REQUIRE(((uintptr_t)x0 & 15u) == 0 && x0_sb % 4 == 0, KON_EINVAL, "x0 rows must be 16-B aligned");
  cin_x0_convert_kernel<<<grid_of(nx / 8, sms), 256, 0, st>>>(x0, x0_sb, x0b, D == 16 ? x0h : nullptr, B, m, D);
  KON_LAUNCH_CHECK("cin_x0_convert_kernel");
  KON_CUDA(cudaMemset2DAsync(dx0, (size_t)dx0_sb * 4, 0, (size_t)m * D * 4, (size_t)B, st));
  bool ragged = false;
  for (int l = 0; l < nl; ++l) {
    const long long tot = (long long)L.n_chunks[l] * L.nkA[l] * 2 * (kDaChunkH * m / 8) * 64;
    cin_pack_w_bwd_kernel<<<grid_of(tot, sms), 256, 0, st>>>(w[l], L.Hp[l] * m, L.N[l], kDaChunkH * m / 8, L.nkA[l],
                                                            L.n_chunks[l], reinterpret_cast<__nv_bfloat16*>(ws + L.wbwd_off[l]));
    KON_LAUNCH_CHECK("cin_pack_w_bwd_kernel");
    if (L.N[l] % 8) ragged = true;
  }
  if (ragged) KON_CUDA(cudaMemsetAsync(ws + L.dz_off[0], 0, 2 * L.dz_bytes, st));
  int cur = 0;
  const bool shortcut = nl >= 2 && L.N8[nl - 1] >= 2;   // last layer's dZ is constant over o: no GEMM needed (see cin_last_da_kernel)
  {
    const int lN = shortcut ? 16 : L.N[nl - 1], lN8 = shortcut ? 2 : L.N8[nl - 1];
    cin_dz_init_kernel<<<grid_of(rows * lN8, sms), 256, 0, st>>>(d_pooled, nl * D, (nl - 1) * D, rows, D, lN, lN8,
                                                                ws + L.dz_off[cur]);
    KON_LAUNCH_CHECK("cin_dz_init_kernel");
  }
  for (int l = nl - 1; l >= 0; --l) {
    const unsigned short* pre = l == 0 ? x0b : reinterpret_cast<const unsigned short*>(sv + L.zt_off[l - 1]);
    // ---- dW_l, dbias_l ----------------------------------------------------------------------
    const bool sc = shortcut && l == nl - 1;
    const bool sc_mma = sc && D == 16;          // v = A^T g with mma.sync (cin_last_dw_kernel)
    if (sc_mma) {
      LastDwArgs z;
      z.x0b = x0b;
      z.pre = pre;
      z.gpool = d_pooled;
      z.gstride = nl * D;
      z.gcol = l * D;
      z.part = reinterpret_cast<float*>(ws + L.part_off);
      const int n_cta = (int)std::min<long long>(B, sms);
      const int Hp16 = (L.Hp[l] + 15) / 16 * 16;
      z.gsum = z.part + (long long)n_cta * Hp16 * 32;
      z.B = B;
      z.Hp = L.Hp[l];
      z.m = m;
      {
        ProfileScope ps("cin_last_dw_kernel", st);
        cin_last_dw_kernel<<<n_cta, 416, 0, st>>>(z);
      }
      KON_LAUNCH_CHECK("cin_last_dw_kernel");
      cin_last_dw_reduce_kernel<<<grid_of((long long)(L.Hp[l] * m + 1) * 32, sms), 256, 0, st>>>(
          z.part, z.gsum, n_cta, Hp16, m, L.Hp[l] * m, L.N[l], dw[l], dbias[l]);
      KON_LAUNCH_CHECK("cin_last_dw_reduce_kernel");
    }
    DwArgs q;
    q.pre = pre;
    q.x0b = x0b;
    q.x0h = x0h;
    q.dz = ws + L.dz_off[cur];
    q.part = reinterpret_cast<float*>(ws + L.part_off);
    q.D = D;
    q.Hp = L.Hp[l];
    q.N8 = sc ? 2 : L.N8[l];
    q.C = L.Hp[l] * m;
    q.n_cp = L.n_cp[l];
    q.n_slices = L.n_slices[l];
    q.kblk = 2u * q.N8 * 128u;
    q.nks_total = rows / 16;
    q.ks_per_slice = (q.nks_total + q.n_slices - 1) / q.n_slices;
    const int Npad = q.N8 * 8;
    if (!sc_mma) {
      {
        ProfileScope ps("cin_dw_tc_kernel", st);
        if (D == 16) cin_dw2_tc_kernel<26><<<q.n_cp * q.n_slices, kDw2Threads, smem_dw, st>>>(q);
        else cin_dw_tc_kernel<26, false><<<q.n_cp * q.n_slices, kTcThreads, smem_stream, st>>>(q);
      }
      KON_LAUNCH_CHECK("cin_dw_tc_kernel");
      if (sc)
        cin_dw_reduce_bcast_kernel<<<grid_of((long long)(q.C + 1) * L.N[l], sms), 256, 0, st>>>(
            q.part, q.n_slices, (long long)q.n_cp * 256 * Npad, Npad, q.C, L.N[l], dw[l], dbias[l]);
      else
        cin_dw_reduce_kernel<<<grid_of((long long)(q.C + 1) * L.N[l], sms), 256, 0, st>>>(
            q.part, q.n_slices, (long long)q.n_cp * 256 * Npad, Npad, q.C, L.N[l], dw[l], dbias[l]);
      KON_LAUNCH_CHECK("cin_dw_reduce_kernel");
    }
    if (sc) {
      const int Hp8 = (L.Hp[l] + 7) / 8 * 8;
      float* T = reinterpret_cast<float*>(ws + L.tail_off);
      cin_wsum_kernel<<<grid_of((long long)Hp8 * kTRow * 32, sms), 256, 0, st>>>(w[l], q.C, L.N[l], m, Hp8, T);
      KON_LAUNCH_CHECK("cin_wsum_kernel");
      const int Nh8 = (L.Hp[l] + 15) / 16 * 2, nkh = (L.Hp[l] + 15) / 16;
      unsigned short* th = reinterpret_cast<unsigned short*>(reinterpret_cast<unsigned char*>(T) + align256((size_t)Hp8 * kTRow * 4));
      unsigned short* ti = th + 2 * 2 * Nh8 * 64;
      cin_tpack_kernel<<<32, 256, 0, st>>>(T, L.Hp[l], m, Nh8, nkh, th, ti);
      KON_LAUNCH_CHECK("cin_tpack_kernel");
      LastDaTcArgs z;
      z.x0b = x0b;
      z.pre = pre;
      z.tpack = reinterpret_cast<const unsigned char*>(th);
      z.gpool = d_pooled;
      z.gstride = nl * D;
      z.gcol = l * D;
      z.gcol_prev = (l - 1) * D;
      z.dz_prev = ws + L.dz_off[cur ^ 1];
      z.dx0 = dx0;
      z.dx_sb = dx0_sb;
      z.rows = rows;
      z.n_tiles = (rows + 127) / 128;
      z.D = D;
      z.Hp = L.Hp[l];
      z.N8p = L.N8[l - 1];
      z.Nh8 = Nh8;
      z.nkh = nkh;
      const size_t smem_lt = (((size_t)2 * 2 * Nh8 * 128 + (size_t)nkh * 1024 + 127) & ~(size_t)127) +
                             2 * (((size_t)128 * L.Hp[l] * 2 + (size_t)128 * m * 2 + 127) & ~(size_t)127);
      {
        ProfileScope ps("cin_last_da_kernel", st);
        cin_last_da_tc_kernel<26><<<(int)std::min<long long>(z.n_tiles, sms), 192, smem_lt, st>>>(z);
      }
      KON_LAUNCH_CHECK("cin_last_da_tc_kernel");
      cur ^= 1;
      continue;
    }
    // ---- dpre_l -> dZ_{l-1}, dx0 ------------------------------------------------------------------
    DaArgs p;
    p.dz = ws + L.dz_off[cur];
    p.wpack = ws + L.wbwd_off[l];
    p.x0b = x0b;
    p.pre = pre;
    p.dx0 = dx0;
    p.dx_sb = dx0_sb;
    p.dpre0 = reinterpret_cast<float*>(ws + L.dpre0_off);
    p.dz_prev = l == 0 ? nullptr : ws + L.dz_off[cur ^ 1];
    p.gpool = d_pooled;
    p.gstride = nl * D;
    p.gcol_prev = (l - 1) * D;
    p.rows = rows;
    p.n_pairs = (rows + 255) / 256;
    p.D = D;
    p.Hp = L.Hp[l];
    p.N8 = L.N8[l];
    p.N8p = l == 0 ? 0 : L.N8[l - 1];
    p.n_chunks = L.n_chunks[l];
    p.nkA = L.nkA[l];
    p.chunk_bytes = L.chunk_bytes[l];
    {
      ProfileScope ps("cin_da_tc_kernel", st);
      cin_da_tc_kernel<26><<<(int)std::min<long long>(p.n_pairs, sms), 320, smem_da, st>>>(p);
    }
    KON_LAUNCH_CHECK("cin_da_tc_kernel");
    cur ^= 1;
  }
  return KON_OK;
}

}  // namespace kon
