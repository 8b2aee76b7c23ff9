// a7: DCN-v1 cross network, forward and backward.
//
// Replaces CrossLayer.call (IL:275-282): per layer Transpose + MatMul ([B,1,D]x[D,1]) +
// BatchMatMul ([B,D,1]x[B,1,1]) + 2 AddV2, i.e. 5*L passes over [B,D].
//
//   fwd: one warp per sample keeps x0 and x_l in registers for all L layers;
//        x_{l+1} = (x0 * s_l + x_l) + b_l with s_l = x_l . w_l (warp reduction).
//        HBM traffic: read x0, write x_L (+ L scalars s_l) = 2*D*4 + 4L bytes/sample.
//   bwd: with G_l = dL/dx_{l+1}:  t_l = x0 . G_l,  dx0 += G_l * s_l,  G_{l-1} = G_l + t_l w_l,
//        finally dx0 += G_{-1}.  x_l is never stored: x_l = x0 * c_l + beta_l with
//        c_l = 1 + sum_{j<l} s_j and beta_l = sum_{j<l} b_j, which turns the weight
//        gradients into batch reductions of x0 and g only:
//            dw_l[d] = sum_b x0[b,d] * (c_l t_l)[b] + beta_l[d] * T_l,   T_l = sum_b t_l[b]
//            db_l[d] = sum_b g[b,d] + sum_{j>l} w_j[d] * T_j
//        Each CTA accumulates its share in registers in a fixed order, writes one partial,
//        and a finalize kernel adds the partials in CTA order (deterministic).
//        HBM traffic: read x0, g, write dx0 = 3*D*4 bytes/sample.
#include "common.cuh"

namespace kon {

constexpr int kCrossThreads = 256;
constexpr int kCrossWarps = kCrossThreads / 32;
constexpr int kCrossMaxL = 8;

template <int NPER>
__global__ void __launch_bounds__(kCrossThreads)
cross_fwd_kernel(const float* __restrict__ x0, long long xs, const float* __restrict__ w,
                 const float* __restrict__ b, float* __restrict__ out, float* __restrict__ s,
                 long long B, int D, int L) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;            // [L][D]
  float* b_s = smem + L * D;    // [L][D]
  for (int i = threadIdx.x; i < L * D; i += kCrossThreads) {
    w_s[i] = w[i];
    b_s[i] = b[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (long long)blockIdx.x * kCrossWarps + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * kCrossWarps;
  for (long long r = warp0; r < B; r += nwarps) {
    float x0r[NPER], xl[NPER];
    const float* xp = x0 + r * xs;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int d = lane + 32 * j;
      x0r[j] = d < D ? __ldg(xp + d) : 0.f;
      xl[j] = x0r[j];
    }
    for (int l = 0; l < L; ++l) {
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < NPER; ++j) {
        const int d = lane + 32 * j;
        if (d < D) dot = fmaf(xl[j], w_s[l * D + d], dot);
      }
      dot = warp_sum(dot);
      if (lane == 0) s[r * L + l] = dot;
#pragma unroll
      for (int j = 0; j < NPER; ++j) {
        const int d = lane + 32 * j;
        if (d < D) xl[j] = __fadd_rn(__fadd_rn(__fmul_rn(x0r[j], dot), xl[j]), b_s[l * D + d]);
      }
    }
    float* op = out + r * (long long)D;
#pragma unroll
    for (int j = 0; j < NPER; ++j) {
      const int d = lane + 32 * j;
      if (d < D) op[d] = xl[j];
    }
  }
}

// Partial layout per CTA: [D] colsum(g) | [L][D] M | [kCrossMaxL] T
__host__ __device__ inline long long cross_partial_floats(int D, int L) {
  return (long long)D + (long long)L * D + kCrossMaxL;
}

template <int NPER>
__global__ void __launch_bounds__(kCrossThreads)
cross_bwd_kernel(const float* __restrict__ x0, long long xs, const float* __restrict__ w,
                 const float* __restrict__ s, const float* __restrict__ g, long long gs,
                 float* __restrict__ dx0, float* __restrict__ partial, long long B, int D, int L) {
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                               // [L][D]
  float* x_s = w_s + L * D;                        // [kCrossWarps][D]
  float* g_s = x_s + kCrossWarps * D;              // [kCrossWarps][D]
  float* u_s = g_s + kCrossWarps * D;              // [kCrossWarps][kCrossMaxL]  c_l * t_l
  float* t_s = u_s + kCrossWarps * kCrossMaxL;     // [kCrossWarps][kCrossMaxL]
  for (int i = threadIdx.x; i < L * D; i += kCrossThreads) w_s[i] = w[i];

  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int CPT = (NPER * 32 + kCrossThreads - 1) / kCrossThreads;   // columns per thread
  float cg[CPT], M[CPT][kCrossMaxL], T = 0.f;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    cg[c] = 0.f;
#pragma unroll
    for (int l = 0; l < kCrossMaxL; ++l) M[c][l] = 0.f;
  }
  __syncthreads();

  const long long n_batches = (B + kCrossWarps - 1) / kCrossWarps;
  for (long long bt = blockIdx.x; bt < n_batches; bt += gridDim.x) {
    const long long r = bt * kCrossWarps + wid;
    float* xr_s = x_s + wid * D;
    float* gr_s = g_s + wid * D;
    if (r < B) {
      float x0r[NPER], G[NPER], dx[NPER];
      const float* xp = x0 + r * xs;
      const float* gp = g + r * gs;
#pragma unroll
      for (int j = 0; j < NPER; ++j) {
        const int d = lane + 32 * j;
        x0r[j] = d < D ? __ldg(xp + d) : 0.f;
        G[j] = d < D ? __ldg(gp + d) : 0.f;
        dx[j] = 0.f;
        if (d < D) {
          xr_s[d] = x0r[j];
          gr_s[d] = G[j];
        }
      }
      float sl[kCrossMaxL], cl[kCrossMaxL];
      float c = 1.f;
#pragma unroll
      for (int l = 0; l < kCrossMaxL; ++l) {
        sl[l] = l < L ? __ldg(s + r * L + l) : 0.f;
        cl[l] = c;
        c += sl[l];
      }
#pragma unroll
      for (int l = kCrossMaxL - 1; l >= 0; --l) {
        if (l < L) {
          float t = 0.f;
#pragma unroll
          for (int j = 0; j < NPER; ++j) t = fmaf(x0r[j], G[j], t);
          t = warp_sum(t);
#pragma unroll
          for (int j = 0; j < NPER; ++j) {
            const int d = lane + 32 * j;
            dx[j] = fmaf(G[j], sl[l], dx[j]);
            if (d < D) G[j] = fmaf(t, w_s[l * D + d], G[j]);
          }
          if (lane == 0) {
            u_s[wid * kCrossMaxL + l] = cl[l] * t;
            t_s[wid * kCrossMaxL + l] = t;
          }
        }
      }
      float* dp = dx0 + r * (long long)D;
#pragma unroll
      for (int j = 0; j < NPER; ++j) {
        const int d = lane + 32 * j;
        if (d < D) dp[d] = dx[j] + G[j];
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        xr_s[d] = 0.f;
        gr_s[d] = 0.f;
      }
      if (lane < kCrossMaxL) {
        u_s[wid * kCrossMaxL + lane] = 0.f;
        t_s[wid * kCrossMaxL + lane] = 0.f;
      }
    }
    __syncthreads();
    // CTA phase: thread owns columns d = tid + 256*c; samples are added in warp order.
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int d = threadIdx.x + kCrossThreads * c;
      if (d < D) {
#pragma unroll
        for (int ws = 0; ws < kCrossWarps; ++ws) {
          const float xv = x_s[ws * D + d];
          cg[c] += g_s[ws * D + d];
#pragma unroll
          for (int l = 0; l < kCrossMaxL; ++l) M[c][l] = fmaf(xv, u_s[ws * kCrossMaxL + l], M[c][l]);
        }
      }
    }
    if (threadIdx.x < kCrossMaxL) {
#pragma unroll
      for (int ws = 0; ws < kCrossWarps; ++ws) T += t_s[ws * kCrossMaxL + threadIdx.x];
    }
    __syncthreads();
  }
  float* p = partial + blockIdx.x * cross_partial_floats(D, L);
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int d = threadIdx.x + kCrossThreads * c;
    if (d < D) {
      p[d] = cg[c];
#pragma unroll
      for (int l = 0; l < kCrossMaxL; ++l)
        if (l < L) p[D + l * D + d] = M[c][l];
    }
  }
  if (threadIdx.x < kCrossMaxL) p[D + L * D + threadIdx.x] = T;
}

__global__ void __launch_bounds__(256)
cross_bwd_finalize_kernel(const float* __restrict__ partial, int n_part,
                          const float* __restrict__ w, const float* __restrict__ b,
                          float* __restrict__ dw, float* __restrict__ db, int D, int L) {
  __shared__ float T[kCrossMaxL];
  const long long pf = cross_partial_floats(D, L);
  if (threadIdx.x < kCrossMaxL) {
    float t = 0.f;
    for (int p = 0; p < n_part; ++p) t += partial[p * pf + D + L * D + threadIdx.x];
    T[threadIdx.x] = t;
  }
  __syncthreads();
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < D; d += gridDim.x * blockDim.x) {
    float cg = 0.f;
    for (int p = 0; p < n_part; ++p) cg += partial[p * pf + d];
    float beta = 0.f;   // beta_l[d] = sum_{j<l} b_j[d]
    for (int l = 0; l < L; ++l) {
      float m = 0.f;
      for (int p = 0; p < n_part; ++p) m += partial[p * pf + D + l * D + d];
      dw[l * D + d] = m + beta * T[l];
      beta += b[l * D + d];
    }
    float tail = 0.f;   // sum_{j>l} w_j[d] T_j
    for (int l = L - 1; l >= 0; --l) {
      db[l * D + d] = cg + tail;
      tail = fmaf(w[l * D + d], T[l], tail);
    }
  }
}

static int cross_bwd_grid(long long B, int sms) {
  const long long n_batches = (B + kCrossWarps - 1) / kCrossWarps;
  return (int)std::max<long long>(1, std::min<long long>(n_batches, (long long)sms * 2));
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
    const int nper = (int)(((D) + 31) / 32);   \
    if (nper <= 4) { CALL(4); }                \
    else if (nper <= 8) { CALL(8); }           \
    else if (nper <= 14) { CALL(14); }         \
    else if (nper <= 20) { CALL(20); }         \
    else if (nper <= 27) { CALL(27); }         \
    else { CALL(32); }                         \
  } while (0)

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
  const size_t smem = 2 * (size_t)L * D * sizeof(float);
  const int grid = (int)std::min<long long>((B + kCrossWarps - 1) / kCrossWarps,
                                            (long long)sm_count_of(dev) * 8);
#define CALL(N)                                                                                  \
  KON_CUDA(cudaFuncSetAttribute(cross_fwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem));                                                     \
  cross_fwd_kernel<N><<<grid, kCrossThreads, smem, st>>>(                                        \
      data_ptr<float>(x0), stride_of(x0, 0), data_ptr<float>(w), data_ptr<float>(b),             \
      data_ptr<float>(out), data_ptr<float>(s), B, (int)D, (int)L)
  KON_CROSS_DISPATCH(D, CALL);
#undef CALL
  KON_LAUNCH_CHECK("cross_fwd_kernel");
  return KON_OK;
}

extern "C" size_t kon_cross_bwd_workspace_bytes(int64_t batch, int32_t dim, int32_t layers,
                                                int device_id) {
  const int grid = cross_bwd_grid(batch, sm_count_of(device_id));
  return (size_t)grid * cross_partial_floats(dim, layers) * sizeof(float);
}

extern "C" int kon_cross_bwd(const DLTensor* x0, const DLTensor* w, const DLTensor* b,
                             const DLTensor* s, const DLTensor* g, DLTensor* dx0, DLTensor* dw,
                             DLTensor* db, DLTensor* workspace, void* stream) {
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
  const int grid = cross_bwd_grid(B, sm_count_of(dev));
  const size_t need = (size_t)grid * cross_partial_floats((int)D, (int)L) * sizeof(float);
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= need, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), need);
  float* partial = data_ptr<float>(workspace);
  const size_t smem = ((size_t)L * D + 2 * (size_t)kCrossWarps * D + 2 * kCrossWarps * kCrossMaxL) *
                      sizeof(float);
#define CALL(N)                                                                                  \
  KON_CUDA(cudaFuncSetAttribute(cross_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                (int)smem));                                                     \
  cross_bwd_kernel<N><<<grid, kCrossThreads, smem, st>>>(                                        \
      data_ptr<float>(x0), stride_of(x0, 0), data_ptr<float>(w), data_ptr<float>(s),             \
      data_ptr<float>(g), stride_of(g, 0), data_ptr<float>(dx0), partial, B, (int)D, (int)L)
  KON_CROSS_DISPATCH(D, CALL);
#undef CALL
  KON_LAUNCH_CHECK("cross_bwd_kernel");
  cross_bwd_finalize_kernel<<<(int)((D + 255) / 256), 256, 0, st>>>(
      partial, grid, data_ptr<float>(w), data_ptr<float>(b), data_ptr<float>(dw),
      data_ptr<float>(db), (int)D, (int)L);
  KON_LAUNCH_CHECK("cross_bwd_finalize_kernel");
  return KON_OK;
}
