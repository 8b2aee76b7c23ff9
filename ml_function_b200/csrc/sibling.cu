// Callers / siblings of the hot path that the reference wires from the same layer classes
// (VERDICT r1 items 4, 5, 7): small fp32 kernels, HBM-bound or tiny.
//
//   kon_pool_sum_fwd   SeqBaseLayer.call on a materialised [B,L,k] sequence embedding (BL:45-46) and the
//                      reduce_sum over the pair axis of AFM's AttentionBaseLayer (IL:364)
//   kon_pairs_fwd/bwd  InnerLayer(use_inner=True, use_add=False): the un-summed list of F(F-1)/2 Hadamard
//                      products in itertools.combinations order (IL:61) -- AFM's / IPNN's input
//   kon_pattn_fwd/bwd  ProductAttentionLayer.call (BL:292-311) on explicit [q,k,v] with both mask modes
//                      (BL:299-306): the form SeqFM / BST reuse (MD:292-301)
#include "common.cuh"

namespace kon {
namespace {

// ------------------------------------------------------------------------------------------------
// sum over axis 1, in order l = 0..L-1
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool_sum_kernel(const float* __restrict__ x, long long sb, long long sl, int L, int k, long long total,
                float* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / k;
    const int c = (int)(idx - b * k);
    const float* p = x + b * sb + c;
    float acc = __ldg(p);
    for (int l = 1; l < L; ++l) acc += __ldg(p + (long long)l * sl);
    out[idx] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// pairwise Hadamard products, pair index p <-> (i<j) in itertools.combinations order
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPairFields = 64;

__device__ __forceinline__ int pair_index(int i, int j, int F) {   // i < j
  return i * F - i * (i + 1) / 2 + (j - i - 1);
}

__global__ void __launch_bounds__(256)
pairs_fwd_kernel(const float* __restrict__ v, long long sb, long long sf, int F, int k, int P, long long B,
                 float* __restrict__ out) {
  __shared__ unsigned char s_i[kMaxPairFields * (kMaxPairFields - 1) / 2];
  __shared__ unsigned char s_j[kMaxPairFields * (kMaxPairFields - 1) / 2];
  for (int i = threadIdx.x; i < F; i += blockDim.x)
    for (int j = i + 1; j < F; ++j) {
      const int p = pair_index(i, j, F);
      s_i[p] = (unsigned char)i;
      s_j[p] = (unsigned char)j;
    }
  __syncthreads();
  const long long total = B * P * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % k);
    const long long t = idx / k;
    const int p = (int)(t % P);
    const long long b = t / P;
    const float* row = v + b * sb + c;
    out[idx] = __ldg(row + s_i[p] * sf) * __ldg(row + s_j[p] * sf);
  }
}

// dv[b,f,c] = sum_{j != f} g[b, p(f,j), c] * v[b,j,c], j ascending
__global__ void __launch_bounds__(256)
pairs_bwd_kernel(const float* __restrict__ v, long long sb, long long sf, const float* __restrict__ g, int F,
                 int k, int P, long long B, float* __restrict__ dv) {
  const long long total = B * F * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % k);
    const long long t = idx / k;
    const int f = (int)(t % F);
    const long long b = t / F;
    const float* row = v + b * sb + c;
    const float* grow = g + (b * P) * (long long)k + c;
    float acc = 0.f;
    for (int j = 0; j < F; ++j) {
      if (j == f) continue;
      const int p = j < f ? pair_index(j, f, F) : pair_index(f, j, F);
      acc = fmaf(__ldg(grow + (long long)p * k), __ldg(row + j * sf), acc);
    }
    dv[idx] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// product attention on explicit q, k, v: one CTA per (head, sample) unit, everything in shared memory
// ------------------------------------------------------------------------------------------------
constexpr int kPaThreads = 128;

struct PattnArgs {
  const float *q, *k, *v, *mask, *go;
  float *out, *dq, *dk, *dv;
  long long units;
  int F, d, mask_mode;
  float scale;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// scores -> P (in sP); sT is scratch of the same size (mask mode 1 needs the unmasked scores)
__device__ __forceinline__ void pattn_scores(const PattnArgs& a, const float* sq, const float* sk, float* sP,
                                             float* sT) {
  const int F = a.F, d = a.d;
  float* raw = a.mask_mode == 1 ? sT : sP;
  for (int e = threadIdx.x; e < F * F; e += kPaThreads) {
    const int i = e / F, j = e - i * F;
    float acc = 0.f;
    for (int c = 0; c < d; ++c) acc = fmaf(sq[i * d + c], sk[j * d + c], acc);
    raw[e] = acc * a.scale;                                     // atten_score /= sqrt(d)  (BL:296-297)
  }
  __syncthreads();
  if (a.mask_mode == 1) {                                       // atten_score = matmul(atten_score, mask)  (BL:300-302)
    for (int e = threadIdx.x; e < F * F; e += kPaThreads) {
      const int i = e / F, j = e - i * F;
      float acc = 0.f;
      for (int t = 0; t < F; ++t) acc = fmaf(sT[i * F + t], __ldg(a.mask + t * F + j), acc);
      sP[e] = sigmoidf_(acc);
    }
  } else {
    for (int e = threadIdx.x; e < F * F; e += kPaThreads) {
      float s = sP[e];
      if (a.mask_mode == 2) s += __ldg(a.mask + e) * (-100000.f);   // BL:303-306
      sP[e] = sigmoidf_(s);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPaThreads) pattn_fwd_kernel(const PattnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int F = a.F, d = a.d, Fd = F * d;
  float *sq = sm, *sk = sq + Fd, *sv = sk + Fd, *sP = sv + Fd, *sT = sP + F * F;
  for (long long u = blockIdx.x; u < a.units; u += gridDim.x) {
    const long long base = u * Fd;
    for (int e = threadIdx.x; e < Fd; e += kPaThreads) {
      sq[e] = __ldg(a.q + base + e);
      sk[e] = __ldg(a.k + base + e);
      sv[e] = __ldg(a.v + base + e);
    }
    __syncthreads();
    pattn_scores(a, sq, sk, sP, sT);
    for (int e = threadIdx.x; e < Fd; e += kPaThreads) {        // atten_v = matmul(atten_score, v)  (BL:309)
      const int i = e / d, c = e - i * d;
      float acc = 0.f;
      for (int j = 0; j < F; ++j) acc = fmaf(sP[i * F + j], sv[j * d + c], acc);
      a.out[base + e] = acc;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kPaThreads) pattn_bwd_kernel(const PattnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int F = a.F, d = a.d, Fd = F * d;
  float *sq = sm, *sk = sq + Fd, *sv = sk + Fd, *sg = sv + Fd, *sP = sg + Fd, *sT = sP + F * F;
  for (long long u = blockIdx.x; u < a.units; u += gridDim.x) {
    const long long base = u * Fd;
    for (int e = threadIdx.x; e < Fd; e += kPaThreads) {
      sq[e] = __ldg(a.q + base + e);
      sk[e] = __ldg(a.k + base + e);
      sv[e] = __ldg(a.v + base + e);
      sg[e] = __ldg(a.go + base + e);
    }
    __syncthreads();
    pattn_scores(a, sq, sk, sP, sT);
    // dv[j,c] = sum_i P[i,j] gO[i,c]
    for (int e = threadIdx.x; e < Fd; e += kPaThreads) {
      const int j = e / d, c = e - j * d;
      float acc = 0.f;
      for (int i = 0; i < F; ++i) acc = fmaf(sP[i * F + j], sg[i * d + c], acc);
      a.dv[base + e] = acc;
    }
    // gS'[i,j] = (gO[i,:] . v[j,:]) * P (1 - P)   -> sT
    for (int e = threadIdx.x; e < F * F; e += kPaThreads) {
      const int i = e / F, j = e - i * F;
      float acc = 0.f;
      for (int c = 0; c < d; ++c) acc = fmaf(sg[i * d + c], sv[j * d + c], acc);
      const float p = sP[e];
      sT[e] = acc * p * (1.f - p);
    }
    __syncthreads();
    if (a.mask_mode == 1) {                                     // gS[i,t] = sum_j gS'[i,j] mask[t,j]   -> sP
      for (int e = threadIdx.x; e < F * F; e += kPaThreads) {
        const int i = e / F, t = e - i * F;
        float acc = 0.f;
        for (int j = 0; j < F; ++j) acc = fmaf(sT[i * F + j], __ldg(a.mask + t * F + j), acc);
        sP[e] = acc * a.scale;
      }
    } else {
      for (int e = threadIdx.x; e < F * F; e += kPaThreads) sP[e] = sT[e] * a.scale;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Fd; e += kPaThreads) {
      const int i = e / d, c = e - i * d;
      float aq = 0.f, ak = 0.f;
      for (int j = 0; j < F; ++j) {
        aq = fmaf(sP[i * F + j], sk[j * d + c], aq);            // dq[i,c] = sum_j gS[i,j] k[j,c]
        ak = fmaf(sP[j * F + i], sq[j * d + c], ak);            // dk[i,c] = sum_j gS[j,i] q[j,c]
      }
      a.dq[base + e] = aq;
      a.dk[base + e] = ak;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Keras binary_crossentropy on probabilities (compile(loss=binary_crossentropy), EX un_seq.py:61):
//   p' = clip(p, eps, 1-eps);  bce = -(y log(p'+eps) + (1-y) log(1-p'+eps));  loss = mean over all elements
// (mean over the last axis, then over the batch == mean over all elements for a dense [B,C] tensor).
// Two fixed-order stages -> deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kBceThreads = 256;
constexpr int kBceMaxCtas = 1024;

__device__ __forceinline__ float block_sum_256(float v, float* s_part) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < kBceThreads / 32; ++w) t += s_part[w];
  return t;   // valid in thread 0
}

__global__ void __launch_bounds__(kBceThreads)
bce_partial_kernel(const float* __restrict__ p, const float* __restrict__ y, long long n, float eps,
                   float* __restrict__ part) {
  __shared__ float s_part[kBceThreads / 32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)kBceThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kBceThreads) {
    const float pc = fminf(fmaxf(p[i], eps), 1.f - eps);
    const float yy = y[i];
    acc -= yy * logf(pc + eps) + (1.f - yy) * logf(1.f - pc + eps);
  }
  const float t = block_sum_256(acc, s_part);
  if (threadIdx.x == 0) part[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kBceThreads)
bce_final_kernel(const float* __restrict__ part, int n_part, float inv_n, float* __restrict__ loss) {
  __shared__ float s_part[kBceThreads / 32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n_part; i += kBceThreads) acc += part[i];
  const float t = block_sum_256(acc, s_part);
  if (threadIdx.x == 0) *loss = t * inv_n;
}

// dp = g * d(mean bce)/dp; the clip passes no gradient outside [eps, 1-eps] (tf.clip_by_value)
__global__ void __launch_bounds__(kBceThreads)
bce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y, const float* __restrict__ g, long long n,
               float eps, float inv_n, float* __restrict__ dp) {
  const float gs = (*g) * inv_n;
  for (long long i = blockIdx.x * (long long)kBceThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kBceThreads) {
    const float pv = p[i];
    float d = 0.f;
    if (pv >= eps && pv <= 1.f - eps) {
      const float yy = y[i];
      d = gs * ((1.f - yy) / (1.f - pv + eps) - yy / (pv + eps));
    }
    dp[i] = d;
  }
}

int grid_for(long long total, int sms, int mult = 16) {
  return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)sms * mult));
}

int check_qkv(const DLTensor* q, const DLTensor* k, const DLTensor* v, int dev, long long* units, int* F, int* d) {
  KON_TRY(check_cuda_tensor(k, "k", dev));
  KON_TRY(check_cuda_tensor(v, "v", dev));
  KON_REQUIRE(q->ndim >= 2 && k->ndim == q->ndim && v->ndim == q->ndim, KON_EINVAL, "q, k, v must share one rank >= 2");
  for (int i = 0; i < q->ndim; ++i)
    KON_REQUIRE(q->shape[i] == k->shape[i] && q->shape[i] == v->shape[i], KON_EUNSUPPORTED,
                "q, k, v must have identical shapes (self-attention form)");
  KON_REQUIRE(is_f32(q) && is_f32(k) && is_f32(v) && is_compact(q) && is_compact(k) && is_compact(v), KON_EINVAL,
              "q, k, v must be compact float32");
  *F = (int)q->shape[q->ndim - 2];
  *d = (int)q->shape[q->ndim - 1];
  long long u = 1;
  for (int i = 0; i < q->ndim - 2; ++i) u *= q->shape[i];
  *units = u;
  KON_REQUIRE(*F >= 1 && *F <= 128 && *d >= 1 && *d <= 128, KON_EUNSUPPORTED, "product attention: F, d must be in [1,128]");
  return KON_OK;
}

int check_mask(const DLTensor* mask, int mask_mode, int F, int dev) {
  KON_REQUIRE(mask_mode >= 0 && mask_mode <= 2, KON_EINVAL, "mask_mode must be 0, 1 or 2");
  if (mask_mode == 0) return KON_OK;
  KON_TRY(check_cuda_tensor(mask, "mask", dev));
  KON_REQUIRE(is_f32(mask) && mask->ndim == 2 && mask->shape[0] == F && mask->shape[1] == F && is_compact(mask),
              KON_EINVAL, "mask must be compact float32 [F,F]");
  return KON_OK;
}

}  // namespace
}  // namespace kon

using namespace kon;

extern "C" int kon_pool_sum_fwd(const DLTensor* x, DLTensor* out, void* stream) {
  KON_TRY(check_cuda_tensor(x, "x"));
  const int dev = x->device.device_id;
  KON_TRY(check_cuda_tensor(out, "out", dev));
  KON_REQUIRE(is_f32(x) && x->ndim == 3 && (x->shape[2] == 1 || stride_of(x, 2) == 1), KON_EINVAL,
              "x must be [B,L,k] float32 with a compact last dim");
  KON_REQUIRE(x->shape[1] >= 1, KON_EINVAL, "x has an empty pooled axis");
  KON_REQUIRE(is_f32(out) && out->ndim == 2 && out->shape[0] == x->shape[0] && out->shape[1] == x->shape[2] &&
                  is_compact(out), KON_EINVAL, "out must be compact [B,k] float32");
  const long long total = x->shape[0] * x->shape[2];
  if (total == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pool_sum_kernel<<<grid_for(total, sm_count_of(dev)), 256, 0, st>>>(data_ptr<float>(x), stride_of(x, 0), stride_of(x, 1),
                                                                     (int)x->shape[1], (int)x->shape[2], total,
                                                                     data_ptr<float>(out));
  KON_LAUNCH_CHECK("pool_sum_kernel");
  return KON_OK;
}

static int check_pairs(const DLTensor* v, const DLTensor* o, const char* oname, int* F, int* k, int* P) {
  KON_TRY(check_cuda_tensor(v, "v"));
  KON_TRY(check_cuda_tensor(o, oname, v->device.device_id));
  KON_REQUIRE(is_f32(v) && v->ndim == 3 && (v->shape[2] == 1 || stride_of(v, 2) == 1), KON_EINVAL,
              "v must be [B,F,k] float32 with a compact last dim");
  *F = (int)v->shape[1];
  *k = (int)v->shape[2];
  KON_REQUIRE(*F >= 2 && *F <= kMaxPairFields, KON_EUNSUPPORTED, "pairwise products need 2 <= F <= %d fields", kMaxPairFields);
  *P = *F * (*F - 1) / 2;
  KON_REQUIRE(is_f32(o) && o->ndim == 3 && o->shape[0] == v->shape[0] && o->shape[1] == *P && o->shape[2] == *k &&
                  is_compact(o), KON_EINVAL, "%s must be compact [B, F(F-1)/2, k] float32", oname);
  return KON_OK;
}

extern "C" int kon_pairs_fwd(const DLTensor* v, DLTensor* out, void* stream) {
  int F, k, P;
  KON_TRY(check_pairs(v, out, "out", &F, &k, &P));
  const long long B = v->shape[0];
  if (B == 0 || k == 0) return KON_OK;
  const int dev = v->device.device_id;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pairs_fwd_kernel<<<grid_for(B * P * k, sm_count_of(dev)), 256, 0, st>>>(data_ptr<float>(v), stride_of(v, 0), stride_of(v, 1),
                                                                          F, k, P, B, data_ptr<float>(out));
  KON_LAUNCH_CHECK("pairs_fwd_kernel");
  return KON_OK;
}

extern "C" int kon_pairs_bwd(const DLTensor* v, const DLTensor* g, DLTensor* dv, void* stream) {
  int F, k, P;
  KON_TRY(check_pairs(v, g, "g", &F, &k, &P));
  const int dev = v->device.device_id;
  KON_TRY(check_cuda_tensor(dv, "dv", dev));
  KON_REQUIRE(is_f32(dv) && dv->ndim == 3 && dv->shape[0] == v->shape[0] && dv->shape[1] == F && dv->shape[2] == k &&
                  is_compact(dv), KON_EINVAL, "dv must be compact [B,F,k] float32");
  const long long B = v->shape[0];
  if (B == 0 || k == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pairs_bwd_kernel<<<grid_for(B * F * k, sm_count_of(dev)), 256, 0, st>>>(data_ptr<float>(v), stride_of(v, 0), stride_of(v, 1),
                                                                          data_ptr<float>(g), F, k, P, B, data_ptr<float>(dv));
  KON_LAUNCH_CHECK("pairs_bwd_kernel");
  return KON_OK;
}

extern "C" int kon_pattn_fwd(const DLTensor* q, const DLTensor* k, const DLTensor* v, const DLTensor* mask,
                             DLTensor* out, int32_t use_scale, int32_t mask_mode, void* stream) {
  KON_TRY(check_cuda_tensor(q, "q"));
  const int dev = q->device.device_id;
  PattnArgs a{};
  KON_TRY(check_qkv(q, k, v, dev, &a.units, &a.F, &a.d));
  KON_TRY(check_mask(mask, mask_mode, a.F, dev));
  KON_TRY(check_cuda_tensor(out, "out", dev));
  KON_REQUIRE(is_f32(out) && is_compact(out) && numel(out) == numel(q), KON_EINVAL, "out must be compact float32 like q");
  if (a.units == 0) return KON_OK;
  const size_t smem = ((size_t)3 * a.F * a.d + 2 * (size_t)a.F * a.F) * 4;
  KON_REQUIRE(smem <= 200 * 1024, KON_EUNSUPPORTED, "product attention: F=%d, d=%d needs %zu B of shared memory", a.F, a.d, smem);
  a.q = data_ptr<float>(q); a.k = data_ptr<float>(k); a.v = data_ptr<float>(v);
  a.mask = mask_mode ? data_ptr<float>(mask) : nullptr;
  a.out = data_ptr<float>(out);
  a.mask_mode = mask_mode;
  a.scale = use_scale ? 1.f / sqrtf((float)a.d) : 1.f;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KON_CUDA(cudaFuncSetAttribute(pattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long long>(a.units, (long long)sm_count_of(dev) * 8);
  pattn_fwd_kernel<<<grid, kPaThreads, smem, st>>>(a);
  KON_LAUNCH_CHECK("pattn_fwd_kernel");
  return KON_OK;
}

extern "C" int kon_pattn_bwd(const DLTensor* q, const DLTensor* k, const DLTensor* v, const DLTensor* mask,
                             const DLTensor* g_out, DLTensor* dq, DLTensor* dk, DLTensor* dv, int32_t use_scale,
                             int32_t mask_mode, void* stream) {
  KON_TRY(check_cuda_tensor(q, "q"));
  const int dev = q->device.device_id;
  PattnArgs a{};
  KON_TRY(check_qkv(q, k, v, dev, &a.units, &a.F, &a.d));
  KON_TRY(check_mask(mask, mask_mode, a.F, dev));
  const DLTensor* ts[4] = {g_out, dq, dk, dv};
  const char* names[4] = {"g_out", "dq", "dk", "dv"};
  for (int i = 0; i < 4; ++i) {
    KON_TRY(check_cuda_tensor(ts[i], names[i], dev));
    KON_REQUIRE(is_f32(ts[i]) && is_compact(ts[i]) && numel(ts[i]) == numel(q), KON_EINVAL,
                "%s must be compact float32 like q", names[i]);
  }
  if (a.units == 0) return KON_OK;
  const size_t smem = ((size_t)4 * a.F * a.d + 2 * (size_t)a.F * a.F) * 4;
  KON_REQUIRE(smem <= 200 * 1024, KON_EUNSUPPORTED, "product attention: F=%d, d=%d needs %zu B of shared memory", a.F, a.d, smem);
  a.q = data_ptr<float>(q); a.k = data_ptr<float>(k); a.v = data_ptr<float>(v);
  a.mask = mask_mode ? data_ptr<float>(mask) : nullptr;
  a.go = data_ptr<float>(g_out);
  a.dq = data_ptr<float>(dq); a.dk = data_ptr<float>(dk); a.dv = data_ptr<float>(dv);
  a.mask_mode = mask_mode;
  a.scale = use_scale ? 1.f / sqrtf((float)a.d) : 1.f;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  KON_CUDA(cudaFuncSetAttribute(pattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long long>(a.units, (long long)sm_count_of(dev) * 8);
  pattn_bwd_kernel<<<grid, kPaThreads, smem, st>>>(a);
  KON_LAUNCH_CHECK("pattn_bwd_kernel");
  return KON_OK;
}

static int check_bce(const DLTensor* p, const DLTensor* y, int* dev, long long* n) {
  KON_TRY(check_cuda_tensor(p, "p"));
  *dev = p->device.device_id;
  KON_TRY(check_cuda_tensor(y, "y", *dev));
  KON_REQUIRE(is_f32(p) && is_f32(y) && is_compact(p) && is_compact(y) && numel(p) == numel(y), KON_EINVAL,
              "p and y must be compact float32 tensors of one size");
  *n = numel(p);
  KON_REQUIRE(*n >= 1, KON_EINVAL, "empty loss input");
  return KON_OK;
}

extern "C" size_t kon_bce_workspace_bytes(void) { return (size_t)kBceMaxCtas * 4; }

extern "C" int kon_bce_fwd(const DLTensor* p, const DLTensor* y, DLTensor* loss, DLTensor* workspace, float eps,
                           void* stream) {
  int dev;
  long long n;
  KON_TRY(check_bce(p, y, &dev, &n));
  KON_TRY(check_cuda_tensor(loss, "loss", dev));
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  KON_REQUIRE(is_f32(loss) && numel(loss) == 1, KON_EINVAL, "loss must be float32[1]");
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= kon_bce_workspace_bytes() &&
                  ((uintptr_t)data_ptr<char>(workspace) & 15u) == 0, KON_EWORKSPACE, "workspace too small or misaligned");
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<long long>((n + kBceThreads - 1) / kBceThreads, kBceMaxCtas);
  float* part = data_ptr<float>(workspace);
  bce_partial_kernel<<<grid, kBceThreads, 0, st>>>(data_ptr<float>(p), data_ptr<float>(y), n, eps, part);
  KON_LAUNCH_CHECK("bce_partial_kernel");
  bce_final_kernel<<<1, kBceThreads, 0, st>>>(part, grid, 1.f / (float)n, data_ptr<float>(loss));
  KON_LAUNCH_CHECK("bce_final_kernel");
  return KON_OK;
}

extern "C" int kon_bce_bwd(const DLTensor* p, const DLTensor* y, const DLTensor* g_loss, DLTensor* dp, float eps,
                           void* stream) {
  int dev;
  long long n;
  KON_TRY(check_bce(p, y, &dev, &n));
  KON_TRY(check_cuda_tensor(g_loss, "g_loss", dev));
  KON_TRY(check_cuda_tensor(dp, "dp", dev));
  KON_REQUIRE(is_f32(g_loss) && numel(g_loss) == 1, KON_EINVAL, "g_loss must be float32[1]");
  KON_REQUIRE(is_f32(dp) && is_compact(dp) && numel(dp) == n, KON_EINVAL, "dp must be compact float32 like p");
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<long long>((n + kBceThreads - 1) / kBceThreads, (long long)sm_count_of(dev) * 8);
  bce_bwd_kernel<<<grid, kBceThreads, 0, st>>>(data_ptr<float>(p), data_ptr<float>(y), data_ptr<float>(g_loss), n, eps,
                                               1.f / (float)n, data_ptr<float>(dp));
  KON_LAUNCH_CHECK("bce_bwd_kernel");
  return KON_OK;
}
