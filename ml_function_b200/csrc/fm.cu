// a5-a6: FM second-order term + linear terms, forward and backward, one pass each.
//
// Replaces InnerLayer.call (IL:59-66: 325 tf.multiply + 324 sequential AddV2 for F=26)
// and FmLayer.call (IL:161-170: +26 AddV2 of the broadcast linear terms).
//
//   fwd: out[b,:] = sum_j v[b,j,:] * (sum_{i<j} v[b,i,:])  (+ lin[b,0] + lin[b,1] + ...)
//        The prefix form has no cancellation (unlike 0.5((sum v)^2 - sum v^2)) and, like
//        the reference, only ever adds products of two embeddings.
//   bwd: dv[b,f,:] = g[b,:] * (S[b,:] - v[b,f,:]),  dlin[b,f] = sum_k g[b,k]
//
// HBM-bound: algorithmic bytes per sample are F*k*4 + F*4 + k*4 (fwd) and
// k*4 + 2*F*k*4 + F*4 (bwd).  V lanes-wide 128-bit loads, all F row loads of a sample
// chunk are independent and issued in batches of 8.
#include "common.cuh"

namespace kon {

template <int V>
struct Vec;
template <>
struct Vec<4> {
  float4 v;
  __device__ static Vec zero() { return {make_float4(0.f, 0.f, 0.f, 0.f)}; }
  __device__ static Vec load(const float* p) { return {__ldg(reinterpret_cast<const float4*>(p))}; }
  __device__ static Vec load_rw(const float* p) { return {*reinterpret_cast<const float4*>(p)}; }   // memory this kernel also writes
  __device__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ void fma(const Vec& a, const Vec& b) {
    v.x = fmaf(a.v.x, b.v.x, v.x); v.y = fmaf(a.v.y, b.v.y, v.y);
    v.z = fmaf(a.v.z, b.v.z, v.z); v.w = fmaf(a.v.w, b.v.w, v.w);
  }
  __device__ void add(const Vec& a) { v.x += a.v.x; v.y += a.v.y; v.z += a.v.z; v.w += a.v.w; }
  __device__ void adds(float s) { v.x += s; v.y += s; v.z += s; v.w += s; }
  __device__ float hsum() const { return (v.x + v.y) + (v.z + v.w); }
  // g * (s - a)
  __device__ static Vec gsub(const Vec& g, const Vec& s, const Vec& a) {
    return {make_float4(g.v.x * (s.v.x - a.v.x), g.v.y * (s.v.y - a.v.y), g.v.z * (s.v.z - a.v.z),
                        g.v.w * (s.v.w - a.v.w))};
  }
};
template <>
struct Vec<1> {
  float v;
  __device__ static Vec zero() { return {0.f}; }
  __device__ static Vec load(const float* p) { return {__ldg(p)}; }
  __device__ static Vec load_rw(const float* p) { return {*p}; }
  __device__ void store(float* p) const { *p = v; }
  __device__ void fma(const Vec& a, const Vec& b) { v = fmaf(a.v, b.v, v); }
  __device__ void add(const Vec& a) { v += a.v; }
  __device__ void adds(float s) { v += s; }
  __device__ float hsum() const { return v; }
  __device__ static Vec gsub(const Vec& g, const Vec& s, const Vec& a) { return {g.v * (s.v - a.v)}; }
};

// One thread per (sample, V-wide column chunk).  cpr = chunks per row = k / V.
template <int V>
__global__ void __launch_bounds__(256)
fm_fwd_kernel(const float* __restrict__ v, long long sb, long long sf, const float* __restrict__ lin,
              long long lsb, long long lsf, int Fl, float* __restrict__ out, long long B, int F, int cpr) {
  const long long total = B * cpr;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / cpr;
    const int c = (int)(idx - b * cpr);
    const float* base = v + b * sb + c * V;
    Vec<V> acc = Vec<V>::zero(), prefix = Vec<V>::zero();
    for (int f0 = 0; f0 < F; f0 += 8) {
      Vec<V> r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        r[u] = (f0 + u < F) ? Vec<V>::load(base + (long long)(f0 + u) * sf) : Vec<V>::zero();
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        acc.fma(r[u], prefix);
        prefix.add(r[u]);
      }
    }
    if (lin) {   // Keras Add: ((cross + lin_0) + lin_1) + ...  (IL:166)
      const float* lp = lin + b * lsb;
      for (int f = 0; f < Fl; ++f) acc.adds(__ldg(lp + f * lsf));
    }
    acc.store(out + b * (long long)cpr * V + c * V);
  }
}

// One thread per (sample, chunk); threads of one sample are adjacent lanes (cpr is a power
// of two <= 32 on this path) so dlin's sum over k finishes with shuffles.
// ACC: dv += ... -- the gradient lands in a buffer that already holds another consumer's gradient of the same
// concat buffer (the first MLP layer's input gradient), instead of autograd summing two [B,W] tensors afterwards.
template <int V, bool ACC>
__global__ void __launch_bounds__(256)
fm_bwd_kernel(const float* __restrict__ v, long long sb, long long sf, const float* __restrict__ g,
              float* __restrict__ dv, long long dsb, long long dsf, float* __restrict__ dlin,
              long long dlsb, long long dlsf, int Fl, long long B, int F, int cpr, int cpr_pad) {
  const long long total = B * cpr_pad;
  const long long stride = (long long)gridDim.x * blockDim.x;   // multiple of 32
  for (long long idx0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
       idx0 - (threadIdx.x & 31) < total; idx0 += stride) {
    const long long b = idx0 / cpr_pad;
    const int c = (int)(idx0 - b * cpr_pad);
    const bool on = idx0 < total && c < cpr;
    Vec<V> gv = Vec<V>::zero(), S = Vec<V>::zero();
    const float* base = v + b * sb + c * V;
    if (on) {
      gv = Vec<V>::load(g + b * (long long)cpr * V + c * V);
      for (int f0 = 0; f0 < F; f0 += 8) {
        Vec<V> r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          r[u] = (f0 + u < F) ? Vec<V>::load(base + (long long)(f0 + u) * sf) : Vec<V>::zero();
#pragma unroll
        for (int u = 0; u < 8; ++u) S.add(r[u]);
      }
      float* dbase = dv + b * dsb + c * V;
      for (int f0 = 0; f0 < F; f0 += 8) {   // second read of v hits L1/L2
        Vec<V> r[8], o[ACC ? 8 : 1];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          r[u] = (f0 + u < F) ? Vec<V>::load(base + (long long)(f0 + u) * sf) : Vec<V>::zero();
          if (ACC) o[u] = (f0 + u < F) ? Vec<V>::load_rw(dbase + (long long)(f0 + u) * dsf) : Vec<V>::zero();
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (f0 + u < F) {
            Vec<V> d = Vec<V>::gsub(gv, S, r[u]);
            if (ACC) d.add(o[u]);
            d.store(dbase + (long long)(f0 + u) * dsf);
          }
      }
    }
    if (dlin) {
      float gs = on ? gv.hsum() : 0.f;
      for (int o = 1; o < cpr_pad; o <<= 1) gs += __shfl_xor_sync(0xffffffffu, gs, o);
      if (on && c == 0) {
        float* dl = dlin + b * dlsb;
        for (int f = 0; f < Fl; ++f) dl[f * dlsf] = gs;
      }
    }
  }
}

static int pow2_ge_i(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace kon

using namespace kon;

static int fm_check_v(const DLTensor* v, const char* name, int dev) {
  KON_TRY(check_cuda_tensor(v, name, dev));
  KON_REQUIRE(is_f32(v) && v->ndim == 3 && stride_of(v, 2) == 1, KON_EINVAL,
              "%s must be float32 [B,F,k] with a compact last dim", name);
  return KON_OK;
}

extern "C" int kon_fm_fwd(const DLTensor* v, const DLTensor* lin, DLTensor* out, void* stream) {
  KON_TRY(fm_check_v(v, "v", -1));
  const int dev = v->device.device_id;
  const int64_t B = v->shape[0], F = v->shape[1], k = v->shape[2];
  KON_TRY(check_cuda_tensor(out, "out", dev));
  KON_REQUIRE(is_f32(out) && out->ndim == 2 && out->shape[0] == B && out->shape[1] == k &&
                  is_compact(out),
              KON_EINVAL, "out must be compact float32 [B,k]");
  const float* lp = nullptr;
  long long lsb = 0, lsf = 0;
  int Fl = 0;
  if (lin) {
    KON_TRY(check_cuda_tensor(lin, "lin", dev));
    KON_REQUIRE(is_f32(lin) && lin->ndim == 2 && lin->shape[0] == B && lin->shape[1] >= 1,
                KON_EINVAL, "lin must be float32 [B,Fl], Fl >= 1");
    Fl = (int)lin->shape[1];
    lp = data_ptr<float>(lin);
    lsb = stride_of(lin, 0);
    lsf = stride_of(lin, 1);
  }
  if (B == 0 || k == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long sb = stride_of(v, 0), sf = stride_of(v, 1);
  const float* vp = data_ptr<float>(v);
  float* op = data_ptr<float>(out);
  const int sms = sm_count_of(dev);
  const bool vec = k % 4 == 0 && aligned16(vp) && aligned16(op) && sb % 4 == 0 && sf % 4 == 0;
  ProfileScope ps("fm_fwd_kernel", st);
  if (vec) {
    const int cpr = (int)(k / 4);
    const long long total = B * cpr;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    fm_fwd_kernel<4><<<grid, 256, 0, st>>>(vp, sb, sf, lp, lsb, lsf, Fl, op, B, (int)F, cpr);
  } else {
    const long long total = B * k;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    fm_fwd_kernel<1><<<grid, 256, 0, st>>>(vp, sb, sf, lp, lsb, lsf, Fl, op, B, (int)F, (int)k);
  }
  KON_LAUNCH_CHECK("fm_fwd_kernel");
  return KON_OK;
}

static int fm_bwd_impl(const DLTensor* v, const DLTensor* g, DLTensor* dv, DLTensor* dlin, bool acc,
                       void* stream) {
  KON_TRY(fm_check_v(v, "v", -1));
  const int dev = v->device.device_id;
  const int64_t B = v->shape[0], F = v->shape[1], k = v->shape[2];
  KON_TRY(check_cuda_tensor(g, "g", dev));
  KON_TRY(fm_check_v(dv, "dv", dev));
  KON_REQUIRE(is_f32(g) && g->ndim == 2 && g->shape[0] == B && g->shape[1] == k && is_compact(g),
              KON_EINVAL, "g must be compact float32 [B,k]");
  KON_REQUIRE(dv->shape[0] == B && dv->shape[1] == F && dv->shape[2] == k, KON_EINVAL,
              "dv must have the shape of v");
  float* dlp = nullptr;
  long long dlsb = 0, dlsf = 0;
  int Fl = 0;
  if (dlin) {
    KON_TRY(check_cuda_tensor(dlin, "dlin", dev));
    KON_REQUIRE(is_f32(dlin) && dlin->ndim == 2 && dlin->shape[0] == B && dlin->shape[1] >= 1,
                KON_EINVAL, "dlin must be float32 [B,Fl], Fl >= 1");
    Fl = (int)dlin->shape[1];
    dlp = data_ptr<float>(dlin);
    dlsb = stride_of(dlin, 0);
    dlsf = stride_of(dlin, 1);
  }
  if (B == 0 || k == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long sb = stride_of(v, 0), sf = stride_of(v, 1);
  const long long dsb = stride_of(dv, 0), dsf = stride_of(dv, 1);
  const float* vp = data_ptr<float>(v);
  const float* gp = data_ptr<float>(g);
  float* dvp = data_ptr<float>(dv);
  const int sms = sm_count_of(dev);
  const bool vec = k % 4 == 0 && k <= 128 && aligned16(vp) && aligned16(gp) && aligned16(dvp) &&
                   sb % 4 == 0 && sf % 4 == 0 && dsb % 4 == 0 && dsf % 4 == 0;
  ProfileScope ps("fm_bwd_kernel", st);
  if (vec) {
    const int cpr = (int)(k / 4), cpr_pad = pow2_ge_i(cpr);
    const long long total = B * cpr_pad;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    if (acc)
      fm_bwd_kernel<4, true><<<grid, 256, 0, st>>>(vp, sb, sf, gp, dvp, dsb, dsf, dlp, dlsb, dlsf, Fl, B,
                                                   (int)F, cpr, cpr_pad);
    else
      fm_bwd_kernel<4, false><<<grid, 256, 0, st>>>(vp, sb, sf, gp, dvp, dsb, dsf, dlp, dlsb, dlsf, Fl, B,
                                                    (int)F, cpr, cpr_pad);
  } else {
    KON_REQUIRE(k <= 32, KON_EUNSUPPORTED,
                "FM backward needs k %% 4 == 0 (16-B aligned rows) or k <= 32; got k=%lld",
                (long long)k);
    const int cpr = (int)k, cpr_pad = pow2_ge_i(cpr);
    const long long total = B * cpr_pad;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
    if (acc)
      fm_bwd_kernel<1, true><<<grid, 256, 0, st>>>(vp, sb, sf, gp, dvp, dsb, dsf, dlp, dlsb, dlsf, Fl, B,
                                                   (int)F, cpr, cpr_pad);
    else
      fm_bwd_kernel<1, false><<<grid, 256, 0, st>>>(vp, sb, sf, gp, dvp, dsb, dsf, dlp, dlsb, dlsf, Fl, B,
                                                    (int)F, cpr, cpr_pad);
  }
  KON_LAUNCH_CHECK("fm_bwd_kernel");
  return KON_OK;
}

extern "C" int kon_fm_bwd(const DLTensor* v, const DLTensor* g, DLTensor* dv, DLTensor* dlin,
                          void* stream) {
  return fm_bwd_impl(v, g, dv, dlin, false, stream);
}

extern "C" int kon_fm_bwd_acc(const DLTensor* v, const DLTensor* g, DLTensor* dv, DLTensor* dlin,
                              void* stream) {
  return fm_bwd_impl(v, g, dv, dlin, true, stream);
}
