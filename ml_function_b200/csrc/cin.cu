// a8: xDeepFM CIN -- C-ABI entry points and the fp32 CUDA-core path (KON_CIN_FP32).
//
// Replaces CIN.call (IL:310-327): per layer split / batched outer-product matmul (which
// materialises [D,B,m,H] in TF) / transpose / reshape / Conv1D(k=1, bias, linear) /
// reduce_sum over the feature maps.
//
// GEMM view used by both paths: rows r = (b,d) (M = B*D), reduction index c = h*m + i
// (IL:317-318), A[r,c] = pre[r,h] * x0[r,i] is never materialised, z[r,o] = A W + bias,
// pre_{l+1}[r,:] = z_l[r,:], pooled[b, l*D + d] = sum_o z_l[r,o].
//
// The fp32 path is the 1e-5 parity mode: straightforward one-thread-per-output kernels
// whose summation order is fixed (deterministic); it is not the performance path -- that
// is cin_tc.cu (bf16 operands on tcgen05, fp32 accumulate).
#include "common.cuh"

namespace kon {

// implemented in cin_tc.cu
size_t cin_tc_saved_bytes(int64_t B, int m, int D, const int32_t* hs, int nl);
size_t cin_tc_workspace_bytes(int64_t B, int m, int D, const int32_t* hs, int nl, int sms);
int cin_tc_fwd(const float* x0, long long x0_sb, const float* const* w, const float* const* bias, int nl,
               const int32_t* hs, int64_t B, int m, int D, float* pooled, void* saved,
               void* workspace, int sms, cudaStream_t st);
int cin_tc_bwd(const float* x0, long long x0_sb, const float* const* w, const float* const* bias, int nl,
               const int32_t* hs, int64_t B, int m, int D, const float* d_pooled,
               const void* saved, float* dx0, long long dx0_sb, float* const* dw, float* const* dbias,
               void* workspace, int sms, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// fp32 CUDA-core kernels
// ---------------------------------------------------------------------------------------
// z[r,o] = sum_{h,i} (pre[r,h] * x0[b,i,d]) * W[h*m+i, o] + bias[o]
__global__ void __launch_bounds__(256)
cin_fwd_simt_kernel(const float* __restrict__ x0, const float* __restrict__ pre, int Hp,
                    const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ z, long long rows, int m, int D, int N) {
  const long long total = rows * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / N;
    const int o = (int)(idx - r * N);
    const long long b = r / D;
    const int d = (int)(r - b * D);
    const float* xb = x0 + b * (long long)m * D + d;   // x0[b,i,d] = xb[i*D]
    float acc = 0.f;
    for (int h = 0; h < Hp; ++h) {
      const float ph = pre ? __ldg(pre + r * Hp + h) : __ldg(xb + (long long)h * D);
      const float* wrow = W + (long long)h * m * N + o;
      for (int i = 0; i < m; ++i)
        acc = fmaf(__fmul_rn(__ldg(xb + (long long)i * D), ph), __ldg(wrow + (long long)i * N), acc);
    }
    z[idx] = acc + bias[o];
  }
}

// pooled[b, col0 + d] = sum_o z[r,o]   (one warp per row)
__global__ void __launch_bounds__(256)
cin_pool_kernel(const float* __restrict__ z, float* __restrict__ pooled, long long rows, int D,
                int N, int col0, int pooled_stride) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = w0; r < rows; r += nw) {
    float acc = 0.f;
    for (int o = lane; o < N; o += 32) acc += z[r * N + o];
    acc = warp_sum(acc);
    if (lane == 0) {
      const long long b = r / D;
      pooled[b * pooled_stride + col0 + (int)(r - b * D)] = acc;
    }
  }
}

// dZ[r,o] = d_pooled[b, col0 + d] (+ dnext[r,o])
__global__ void __launch_bounds__(256)
cin_dz_kernel(const float* __restrict__ d_pooled, const float* dnext,   // dnext may alias dz
              float* dz, long long rows, int D, int N, int col0, int pooled_stride) {
  const long long total = rows * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / N;
    const long long b = r / D;
    float v = d_pooled[b * pooled_stride + col0 + (int)(r - b * D)];
    if (dnext) v += dnext[idx];
    dz[idx] = v;
  }
}

// dW[c,o] = sum_r pre[r,h] x0[r,i] dZ[r,o];  c == K: dbias[o] = sum_r dZ[r,o]
__global__ void __launch_bounds__(256)
cin_dw_simt_kernel(const float* __restrict__ x0, const float* __restrict__ pre, int Hp,
                   const float* __restrict__ dz, float* __restrict__ dW, float* __restrict__ dbias,
                   long long rows, int m, int D, int N) {
  const long long K = (long long)Hp * m;
  const long long total = (K + 1) * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long c = idx / N;
    const int o = (int)(idx - c * N);
    float acc = 0.f;
    if (c == K) {
      for (long long r = 0; r < rows; ++r) acc += dz[r * N + o];
      dbias[o] = acc;
    } else {
      const int h = (int)(c / m), i = (int)(c - (long long)h * m);
      for (long long r = 0; r < rows; ++r) {
        const long long b = r / D;
        const int d = (int)(r - b * D);
        const float* xb = x0 + b * (long long)m * D + d;
        const float ph = pre ? pre[r * Hp + h] : xb[(long long)h * D];
        acc = fmaf(__fmul_rn(xb[(long long)i * D], ph), dz[r * N + o], acc);
      }
      dW[idx] = acc;
    }
  }
}

// dpre[r,h] = sum_i x0[r,i] sum_o W[h*m+i,o] dZ[r,o]
__global__ void __launch_bounds__(256)
cin_dpre_simt_kernel(const float* __restrict__ x0, const float* __restrict__ W,
                     const float* __restrict__ dz, float* __restrict__ dpre, long long rows,
                     int Hp, int m, int D, int N) {
  const long long total = rows * Hp;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / Hp;
    const int h = (int)(idx - r * Hp);
    const long long b = r / D;
    const float* xb = x0 + b * (long long)m * D + (int)(r - b * D);
    const float* dzr = dz + r * N;
    float acc = 0.f;
    for (int i = 0; i < m; ++i) {
      const float* wrow = W + ((long long)h * m + i) * N;
      float e = 0.f;
      for (int o = 0; o < N; ++o) e = fmaf(__ldg(wrow + o), dzr[o], e);
      acc = fmaf(xb[(long long)i * D], e, acc);
    }
    dpre[idx] = acc;
  }
}

// dx0[b,i,d] (+)= sum_h pre[r,h] sum_o W[h*m+i,o] dZ[r,o]  (+ dpre0[r,i] on layer 0, where
// pre == x0 so the "pre" role of x0 contributes too)
__global__ void __launch_bounds__(256)
cin_dx0_simt_kernel(const float* __restrict__ x0, const float* __restrict__ pre, int Hp,
                    const float* __restrict__ W, const float* __restrict__ dz,
                    const float* __restrict__ dpre0, float* __restrict__ dx0, long long rows, int m,
                    int D, int N, int accumulate) {
  const long long total = rows * m;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long r = idx / m;
    const int i = (int)(idx - r * m);
    const long long b = r / D;
    const int d = (int)(r - b * D);
    const float* xb = x0 + b * (long long)m * D + d;
    const float* dzr = dz + r * N;
    float acc = 0.f;
    for (int h = 0; h < Hp; ++h) {
      const float* wrow = W + ((long long)h * m + i) * N;
      float e = 0.f;
      for (int o = 0; o < N; ++o) e = fmaf(__ldg(wrow + o), dzr[o], e);
      const float ph = pre ? pre[r * Hp + h] : xb[(long long)h * D];
      acc = fmaf(ph, e, acc);
    }
    if (dpre0) acc += dpre0[r * m + i];
    float* dst = dx0 + b * (long long)m * D + (long long)i * D + d;
    *dst = accumulate ? *dst + acc : acc;
  }
}

static int grid_for(long long total, int sms) {
  return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * 32));
}

static int cin_simt_fwd(const float* x0, const float* const* w, const float* const* bias, int nl,
                        const int32_t* hs, int64_t B, int m, int D, float* pooled, float* saved,
                        int sms, cudaStream_t st) {
  const long long rows = B * D;
  const float* pre = nullptr;
  int Hp = m;
  float* z = saved;
  for (int l = 0; l < nl; ++l) {
    const int N = hs[l];
    cin_fwd_simt_kernel<<<grid_for(rows * N, sms), 256, 0, st>>>(x0, pre, Hp, w[l], bias[l], z,
                                                                 rows, m, D, N);
    KON_LAUNCH_CHECK("cin_fwd_simt_kernel");
    cin_pool_kernel<<<grid_for(rows * 32, sms), 256, 0, st>>>(z, pooled, rows, D, N, l * D, nl * D);
    KON_LAUNCH_CHECK("cin_pool_kernel");
    pre = z;
    Hp = N;
    z += rows * N;
  }
  return KON_OK;
}

static int cin_simt_bwd(const float* x0, const float* const* w, int nl, const int32_t* hs,
                        int64_t B, int m, int D, const float* d_pooled, const float* saved,
                        float* dx0, float* const* dw, float* const* dbias, float* workspace,
                        int sms, cudaStream_t st) {
  const long long rows = B * D;
  int hmax = m;
  for (int l = 0; l < nl; ++l) hmax = std::max(hmax, (int)hs[l]);
  // Two buffers: `dz` holds dZ_l (built in place on top of dpre_{l+1}, same shape), `dpre`
  // receives dpre_l; then they swap roles.
  float* dz = workspace;
  float* dpre = workspace + rows * hmax;
  long long off[KON_CIN_MAX_LAYERS + 1];
  off[0] = 0;
  for (int l = 0; l < nl; ++l) off[l + 1] = off[l] + rows * hs[l];
  bool have_next = false;
  for (int l = nl - 1; l >= 0; --l) {
    const int N = hs[l];
    const int Hp = l == 0 ? m : hs[l - 1];
    const float* pre = l == 0 ? nullptr : saved + off[l - 1];
    cin_dz_kernel<<<grid_for(rows * N, sms), 256, 0, st>>>(d_pooled, have_next ? dz : nullptr, dz,
                                                           rows, D, N, l * D, nl * D);
    KON_LAUNCH_CHECK("cin_dz_kernel");
    cin_dw_simt_kernel<<<grid_for(((long long)Hp * m + 1) * N, sms), 256, 0, st>>>(
        x0, pre, Hp, dz, dw[l], dbias[l], rows, m, D, N);
    KON_LAUNCH_CHECK("cin_dw_simt_kernel");
    cin_dpre_simt_kernel<<<grid_for(rows * Hp, sms), 256, 0, st>>>(x0, w[l], dz, dpre, rows, Hp, m,
                                                                   D, N);
    KON_LAUNCH_CHECK("cin_dpre_simt_kernel");
    cin_dx0_simt_kernel<<<grid_for(rows * m, sms), 256, 0, st>>>(
        x0, pre, Hp, w[l], dz, l == 0 ? dpre : nullptr, dx0, rows, m, D, N, l == nl - 1 ? 0 : 1);
    KON_LAUNCH_CHECK("cin_dx0_simt_kernel");
    std::swap(dz, dpre);   // dpre_l becomes the seed of dZ_{l-1}
    have_next = true;
  }
  return KON_OK;
}

}  // namespace kon

using namespace kon;

namespace {
struct CinShape {
  int64_t B;
  int m, D, nl;
  int32_t hs[KON_CIN_MAX_LAYERS];
};

// [B,m,D] float32 whose last two dims are compact; the batch stride is free in the bf16 path
// (x0 / dx0 may be windows of the [B, W] concat buffer and of its gradient)
bool rows_compact(const DLTensor* t) {
  return t->ndim == 3 && (t->shape[2] == 1 || stride_of(t, 2) == 1) &&
         (t->shape[1] == 1 || stride_of(t, 1) == t->shape[2]);
}

int cin_check(const DLTensor* x0, const DLTensor* const* w, const DLTensor* const* bias,
              int32_t n_layers, int32_t precision, CinShape* s) {
  KON_TRY(check_cuda_tensor(x0, "x0"));
  const int dev = x0->device.device_id;
  KON_REQUIRE(is_f32(x0) && x0->ndim == 3 && (precision == KON_CIN_BF16 ? rows_compact(x0) : is_compact(x0)),
              KON_EINVAL, "x0 must be float32 [B,m,D] (compact; KON_CIN_BF16: free batch stride)");
  KON_REQUIRE(n_layers >= 1 && n_layers <= KON_CIN_MAX_LAYERS, KON_EUNSUPPORTED,
              "n_layers=%d outside [1,%d]", n_layers, KON_CIN_MAX_LAYERS);
  KON_REQUIRE(w != nullptr && bias != nullptr, KON_EINVAL, "w / bias arrays are NULL");
  s->B = x0->shape[0];
  s->m = (int)x0->shape[1];
  s->D = (int)x0->shape[2];
  s->nl = n_layers;
  int64_t hp = s->m;
  for (int l = 0; l < n_layers; ++l) {
    KON_TRY(check_cuda_tensor(w[l], "w[l]", dev));
    KON_TRY(check_cuda_tensor(bias[l], "bias[l]", dev));
    KON_REQUIRE(is_f32(w[l]) && w[l]->ndim == 2 && w[l]->shape[0] == hp * s->m && is_compact(w[l]),
                KON_EINVAL, "w[%d] must be compact float32 [%lld, H_l]", l, (long long)(hp * s->m));
    const int64_t n = w[l]->shape[1];
    KON_REQUIRE(n >= 1, KON_EINVAL, "w[%d] has no output maps", l);
    KON_REQUIRE(is_f32(bias[l]) && numel(bias[l]) == n && is_compact(bias[l]), KON_EINVAL,
                "bias[%d] must be float32 [%lld]", l, (long long)n);
    s->hs[l] = (int32_t)n;
    hp = n;
  }
  return KON_OK;
}
}  // namespace

extern "C" size_t kon_cin_saved_bytes(int64_t batch, int32_t m, int32_t D,
                                      const int32_t* layer_sizes, int32_t n_layers,
                                      int32_t precision) {
  if (precision == KON_CIN_BF16) return cin_tc_saved_bytes(batch, m, D, layer_sizes, n_layers);
  size_t tot = 0;
  for (int l = 0; l < n_layers; ++l) tot += (size_t)batch * D * layer_sizes[l] * sizeof(float);
  return tot ? tot : 4;
}

extern "C" size_t kon_cin_workspace_bytes(int64_t batch, int32_t m, int32_t D,
                                          const int32_t* layer_sizes, int32_t n_layers,
                                          int32_t precision, int device_id) {
  if (precision == KON_CIN_BF16)
    return cin_tc_workspace_bytes(batch, m, D, layer_sizes, n_layers, sm_count_of(device_id));
  int hmax = m;
  for (int l = 0; l < n_layers; ++l) hmax = std::max(hmax, (int)layer_sizes[l]);
  const size_t tot = 2 * (size_t)batch * D * hmax * sizeof(float);
  return tot ? tot : 4;
}

extern "C" int kon_cin_fwd(const DLTensor* x0, const DLTensor* const* w,
                           const DLTensor* const* bias, int32_t n_layers, DLTensor* pooled,
                           DLTensor* saved, DLTensor* workspace, int32_t precision, void* stream) {
  CinShape s;
  KON_TRY(cin_check(x0, w, bias, n_layers, precision, &s));
  const int dev = x0->device.device_id;
  KON_TRY(check_cuda_tensor(pooled, "pooled", dev));
  KON_TRY(check_cuda_tensor(saved, "saved", dev));
  KON_REQUIRE(is_f32(pooled) && pooled->ndim == 2 && pooled->shape[0] == s.B &&
                  pooled->shape[1] == (int64_t)s.nl * s.D && is_compact(pooled),
              KON_EINVAL, "pooled must be compact float32 [B, n_layers*D]");
  KON_REQUIRE(precision == KON_CIN_FP32 || precision == KON_CIN_BF16, KON_EINVAL,
              "unknown precision %d", precision);
  const size_t need = kon_cin_saved_bytes(s.B, s.m, s.D, s.hs, s.nl, precision);
  KON_REQUIRE(is_u8(saved) && (size_t)numel(saved) >= need, KON_EWORKSPACE,
              "saved has %lld bytes, need %zu", (long long)numel(saved), need);
  if (s.B == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* wp[KON_CIN_MAX_LAYERS];
  const float* bp[KON_CIN_MAX_LAYERS];
  for (int l = 0; l < s.nl; ++l) {
    wp[l] = data_ptr<float>(w[l]);
    bp[l] = data_ptr<float>(bias[l]);
  }
  const int sms = sm_count_of(dev);
  if (precision == KON_CIN_FP32)
    return cin_simt_fwd(data_ptr<float>(x0), wp, bp, s.nl, s.hs, s.B, s.m, s.D,
                        data_ptr<float>(pooled), data_ptr<float>(saved), sms, st);
  const size_t wneed = cin_tc_workspace_bytes(s.B, s.m, s.D, s.hs, s.nl, sms);
  KON_REQUIRE(workspace != nullptr, KON_EINVAL, "workspace is NULL");
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= wneed, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), wneed);
  return cin_tc_fwd(data_ptr<float>(x0), s.B > 1 ? stride_of(x0, 0) : (long long)s.m * s.D, wp, bp, s.nl, s.hs, s.B, s.m, s.D,
                    data_ptr<float>(pooled), data_ptr<char>(saved), data_ptr<char>(workspace), sms,
                    st);
}

extern "C" int kon_cin_bwd(const DLTensor* x0, const DLTensor* const* w,
                           const DLTensor* const* bias, int32_t n_layers,
                           const DLTensor* d_pooled, const DLTensor* saved, DLTensor* dx0,
                           DLTensor* const* dw, DLTensor* const* dbias, DLTensor* workspace,
                           int32_t precision, void* stream) {
  CinShape s;
  KON_TRY(cin_check(x0, w, bias, n_layers, precision, &s));
  const int dev = x0->device.device_id;
  KON_TRY(check_cuda_tensor(d_pooled, "d_pooled", dev));
  KON_TRY(check_cuda_tensor(saved, "saved", dev));
  KON_TRY(check_cuda_tensor(dx0, "dx0", dev));
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  KON_REQUIRE(dw != nullptr && dbias != nullptr, KON_EINVAL, "dw / dbias arrays are NULL");
  KON_REQUIRE(is_f32(d_pooled) && d_pooled->ndim == 2 && d_pooled->shape[0] == s.B &&
                  d_pooled->shape[1] == (int64_t)s.nl * s.D && is_compact(d_pooled),
              KON_EINVAL, "d_pooled must be compact float32 [B, n_layers*D]");
  KON_REQUIRE(is_f32(dx0) && numel(dx0) == numel(x0) && dx0->ndim == 3 &&
                  (precision == KON_CIN_BF16 ? rows_compact(dx0) : is_compact(dx0)),
              KON_EINVAL, "dx0 must be float32 [B,m,D] like x0 (compact; KON_CIN_BF16: free batch stride)");
  KON_REQUIRE(precision == KON_CIN_FP32 || precision == KON_CIN_BF16, KON_EINVAL,
              "unknown precision %d", precision);
  float* dwp[KON_CIN_MAX_LAYERS];
  float* dbp[KON_CIN_MAX_LAYERS];
  const float* wp[KON_CIN_MAX_LAYERS];
  const float* bp[KON_CIN_MAX_LAYERS];
  for (int l = 0; l < s.nl; ++l) {
    KON_TRY(check_cuda_tensor(dw[l], "dw[l]", dev));
    KON_TRY(check_cuda_tensor(dbias[l], "dbias[l]", dev));
    KON_REQUIRE(is_f32(dw[l]) && numel(dw[l]) == numel(w[l]) && is_compact(dw[l]), KON_EINVAL,
                "dw[%d] must be compact float32 like w[%d]", l, l);
    KON_REQUIRE(is_f32(dbias[l]) && numel(dbias[l]) == s.hs[l] && is_compact(dbias[l]), KON_EINVAL,
                "dbias[%d] must be float32 [H_l]", l);
    dwp[l] = data_ptr<float>(dw[l]);
    dbp[l] = data_ptr<float>(dbias[l]);
    wp[l] = data_ptr<float>(w[l]);
    bp[l] = data_ptr<float>(bias[l]);
  }
  const size_t sneed = kon_cin_saved_bytes(s.B, s.m, s.D, s.hs, s.nl, precision);
  KON_REQUIRE(is_u8(saved) && (size_t)numel(saved) >= sneed, KON_EWORKSPACE,
              "saved has %lld bytes, need %zu", (long long)numel(saved), sneed);
  const size_t wneed = kon_cin_workspace_bytes(s.B, s.m, s.D, s.hs, s.nl, precision, dev);
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= wneed, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), wneed);
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = sm_count_of(dev);
  if (s.B == 0) {
    for (int l = 0; l < s.nl; ++l) {
      KON_CUDA(cudaMemsetAsync(dwp[l], 0, numel(dw[l]) * 4, st));
      KON_CUDA(cudaMemsetAsync(dbp[l], 0, (size_t)s.hs[l] * 4, st));
    }
    return KON_OK;
  }
  if (precision == KON_CIN_FP32)
    return cin_simt_bwd(data_ptr<float>(x0), wp, s.nl, s.hs, s.B, s.m, s.D,
                        data_ptr<float>(d_pooled), data_ptr<float>(saved), data_ptr<float>(dx0),
                        dwp, dbp, data_ptr<float>(workspace), sms, st);
  const long long cs = (long long)s.m * s.D;
  return cin_tc_bwd(data_ptr<float>(x0), s.B > 1 ? stride_of(x0, 0) : cs, wp, bp, s.nl, s.hs, s.B, s.m, s.D,
                    data_ptr<float>(d_pooled), data_ptr<char>(saved), data_ptr<float>(dx0),
                    s.B > 1 ? stride_of(dx0, 0) : cs, dwp,
                    dbp, data_ptr<char>(workspace), sms, st);
}
