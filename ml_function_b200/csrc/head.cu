// a12 (heads): skinny dense layer  y = [x1 | x2] W + b  with N = 1 or 2 output units.
//
// Replaces MergeScoreLayer.call (CL:86-100: Flatten + Concatenate + Dense(2)) and the Dense(1)
// logit layers of XDeepFM (IL:323-325 / CL:190).  A [65536 x 912] x [912 x 2] product is not a
// tensor-core problem: library GEMMs fall back to SIMT kernels that take 100-200 us each for the
// forward, dX and dW, and the Concatenate is a full copy.  Here it is what it is, an HBM-bound
// row pass:
//   fwd : one warp per sample; a lane owns 4 consecutive columns per 128-column group (128-bit loads
//         when the rows allow it), W in shared memory, N warp reductions.  Both inputs are read in
//         place (no concat buffer).
//   bwd : same pass: dx = gy W^T written straight into the two gradient buffers, dW = X^T gy and
//         db = sum gy accumulated in registers per lane, then warp -> CTA (warp order) -> one partial
//         per CTA -> fixed-order reduction kernel (deterministic).
#include "common.cuh"

namespace kon {
namespace {

constexpr int kHeadThreads = 256;
constexpr int kHeadWarps = kHeadThreads / 32;

struct HeadIn {
  const float* x1;
  const float* x2;
  long long s1, s2;   // row strides (elements)
  int D1, D;          // columns of x1, total columns
  int vec1, vec2;     // rows of x1 / x2 can move as 128-bit accesses
};

__device__ __forceinline__ float4 head_load(const HeadIn& in, long long r, int d0) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d0 >= in.D) return v;
  const bool first = d0 < in.D1;
  const float* p = first ? in.x1 + r * in.s1 + d0 : in.x2 + r * in.s2 + (d0 - in.D1);
  const int lim = first ? in.D1 : in.D;     // groups never straddle the two inputs (D1 % 4 == 0)
  if ((first ? in.vec1 : in.vec2) && d0 + 3 < lim) return __ldg(reinterpret_cast<const float4*>(p));
  v.x = __ldg(p);
  if (d0 + 1 < lim) v.y = __ldg(p + 1);
  if (d0 + 2 < lim) v.z = __ldg(p + 2);
  if (d0 + 3 < lim) v.w = __ldg(p + 3);
  return v;
}

struct HeadOut {
  float* x1;
  float* x2;
  long long s1, s2;
  int D1, D;
  int vec1, vec2;
};

__device__ __forceinline__ void head_store(const HeadOut& o, long long r, int d0, float4 v) {
  if (d0 >= o.D) return;
  const bool first = d0 < o.D1;
  float* base = first ? o.x1 : o.x2;
  if (base == nullptr) return;
  float* p = first ? base + r * o.s1 + d0 : base + r * o.s2 + (d0 - o.D1);
  const int lim = first ? o.D1 : o.D;
  if ((first ? o.vec1 : o.vec2) && d0 + 3 < lim) {
    *reinterpret_cast<float4*>(p) = v;
    return;
  }
  p[0] = v.x;
  if (d0 + 1 < lim) p[1] = v.y;
  if (d0 + 2 < lim) p[2] = v.z;
  if (d0 + 3 < lim) p[3] = v.w;
}

// W [D][N] -> shared, transposed and zero padded: w_s[n][Dp]
template <int N>
__device__ __forceinline__ void head_stage_w(const float* __restrict__ w, int D, int Dp, float* w_s) {
  for (int i = threadIdx.x; i < N * Dp; i += kHeadThreads) {
    const int n = i / Dp, d = i - n * Dp;
    w_s[i] = d < D ? w[d * N + n] : 0.f;
  }
  __syncthreads();
}

template <int NJ, int N>
__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const HeadIn in, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ y, long long B) {
  constexpr int Dp = NJ * 128;
  extern __shared__ __align__(16) float w_s[];
  head_stage_w<N>(w, in.D, Dp, w_s);
  const int lane = threadIdx.x & 31;
  float bias[N];
#pragma unroll
  for (int n = 0; n < N; ++n) bias[n] = b ? b[n] : 0.f;
  for (long long r = (long long)blockIdx.x * kHeadWarps + (threadIdx.x >> 5); r < B;
       r += (long long)gridDim.x * kHeadWarps) {
    float acc[N];
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = 0.f;
    float4 x[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) x[j] = head_load(in, r, 128 * j + 4 * lane);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d0 = 128 * j + 4 * lane;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const float4 wv = *reinterpret_cast<const float4*>(w_s + n * Dp + d0);
        acc[n] = fmaf(x[j].x, wv.x, acc[n]);
        acc[n] = fmaf(x[j].y, wv.y, acc[n]);
        acc[n] = fmaf(x[j].z, wv.z, acc[n]);
        acc[n] = fmaf(x[j].w, wv.w, acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < N; ++n) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) y[r * N + n] = acc[n] + bias[n];
    }
  }
}

// partial layout per CTA: [N][Dp] dW (transposed) | [N] db
template <int NJ, int N>
__global__ void __launch_bounds__(kHeadThreads, 2)
head_bwd_kernel(const HeadIn in, const HeadOut out, const float* __restrict__ w,
                const float* __restrict__ gy, float* __restrict__ partial, long long B) {
  constexpr int Dp = NJ * 128;
  extern __shared__ __align__(16) float smem[];
  float* w_s = smem;                       // [N][Dp]
  float* acc_s = w_s + N * Dp;             // [N][Dp] + [N]: CTA accumulator (epilogue)
  head_stage_w<N>(w, in.D, Dp, w_s);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float4 dw[NJ][N];
  float db[N];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    db[n] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) dw[j][n] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long r = (long long)blockIdx.x * kHeadWarps + wid; r < B; r += (long long)gridDim.x * kHeadWarps) {
    float g[N];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      g[n] = __ldg(gy + r * N + n);
      db[n] += g[n];
    }
    // all the row's loads are issued before the first store: dx may alias x as far as the compiler
    // knows, so a store in between would serialise the remaining loads behind it
    float4 x[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) x[j] = head_load(in, r, 128 * j + 4 * lane);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d0 = 128 * j + 4 * lane;
      float4 dx = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const float4 wv = *reinterpret_cast<const float4*>(w_s + n * Dp + d0);
        dx.x = fmaf(g[n], wv.x, dx.x); dx.y = fmaf(g[n], wv.y, dx.y);
        dx.z = fmaf(g[n], wv.z, dx.z); dx.w = fmaf(g[n], wv.w, dx.w);
        dw[j][n].x = fmaf(x[j].x, g[n], dw[j][n].x); dw[j][n].y = fmaf(x[j].y, g[n], dw[j][n].y);
        dw[j][n].z = fmaf(x[j].z, g[n], dw[j][n].z); dw[j][n].w = fmaf(x[j].w, g[n], dw[j][n].w);
      }
      head_store(out, r, d0, dx);
    }
  }
  // warps add their fragments into the CTA accumulator in warp order (deterministic)
  for (int ws = 0; ws < kHeadWarps; ++ws) {
    __syncthreads();
    if (wid != ws) continue;
#pragma unroll
    for (int n = 0; n < N; ++n) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float4* p = reinterpret_cast<float4*>(acc_s + n * Dp + 128 * j + 4 * lane);
        float4 v = dw[j][n];
        if (ws) { const float4 o = *p; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
        *p = v;
      }
      if (lane == 0) acc_s[N * Dp + n] = ws ? acc_s[N * Dp + n] + db[n] : db[n];
    }
  }
  __syncthreads();
  float* p = partial + (long long)blockIdx.x * (N * Dp + N);
  for (int i = threadIdx.x; i < N * Dp + N; i += kHeadThreads) p[i] = acc_s[i];
}

// out[i] = sum_p partial[p][i], p in fixed order: 32 columns x 8 slices per CTA
__global__ void __launch_bounds__(256)
head_reduce_kernel(const float* __restrict__ partial, int n_part, int width, int N, int Dp, int D,
                   float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[8][32];
  const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + col;
  const int per = (n_part + 7) / 8;
  const int lo = slice * per, hi = min(n_part, lo + per);
  float a0 = 0.f, a1 = 0.f;
  if (i < width) {
    int p = lo;
    for (; p + 1 < hi; p += 2) {
      a0 += partial[(long long)p * width + i];
      a1 += partial[(long long)(p + 1) * width + i];
    }
    if (p < hi) a0 += partial[(long long)p * width + i];
  }
  sm[slice][col] = a0 + a1;
  __syncthreads();
  if (slice == 0 && i < width) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sm[k][col];
    if (i < N * Dp) {
      const int n = i / Dp, d = i - n * Dp;
      if (d < D) dw[d * N + n] = v;
    } else if (db) {
      db[i - N * Dp] = v;
    }
  }
}

int head_nj(int64_t D) {
  const int nj = (int)((D + 127) / 128);
  return nj <= 1 ? 1 : nj <= 2 ? 2 : nj <= 4 ? 4 : 8;
}
int head_grid(long long B, int sms) {
  return (int)std::max<long long>(1, std::min<long long>((B + kHeadWarps - 1) / kHeadWarps, (long long)sms * 4));
}
int vec_ok(const float* p, int64_t stride) { return (p && aligned16(p) && stride % 4 == 0) ? 1 : 0; }

int head_check(const DLTensor* x1, const DLTensor* x2, const DLTensor* w, HeadIn* in, int64_t* B, int* N) {
  KON_TRY(check_cuda_tensor(x1, "x1"));
  const int dev = x1->device.device_id;
  KON_TRY(check_cuda_tensor(w, "w", dev));
  KON_REQUIRE(is_f32(x1) && x1->ndim == 2 && stride_of(x1, 1) == 1, KON_EINVAL,
              "x1 must be float32 [B,D1] with a compact last dim");
  *B = x1->shape[0];
  in->x1 = data_ptr<float>(x1);
  in->s1 = stride_of(x1, 0);
  in->D1 = (int)x1->shape[1];
  in->x2 = nullptr;
  in->s2 = 0;
  int64_t D = x1->shape[1];
  if (x2) {
    KON_TRY(check_cuda_tensor(x2, "x2", dev));
    KON_REQUIRE(is_f32(x2) && x2->ndim == 2 && x2->shape[0] == *B && stride_of(x2, 1) == 1, KON_EINVAL,
                "x2 must be float32 [B,D2] with a compact last dim");
    KON_REQUIRE(x1->shape[1] % 4 == 0, KON_EUNSUPPORTED, "two-input head needs D1 %% 4 == 0 (got %lld)",
                (long long)x1->shape[1]);
    in->x2 = data_ptr<float>(x2);
    in->s2 = stride_of(x2, 0);
    D += x2->shape[1];
  }
  KON_REQUIRE(D >= 1 && D <= 1024, KON_EUNSUPPORTED, "head width D=%lld outside [1,1024]", (long long)D);
  in->D = (int)D;
  KON_REQUIRE(is_f32(w) && w->ndim == 2 && w->shape[0] == D && is_compact(w) &&
                  (w->shape[1] == 1 || w->shape[1] == 2),
              KON_EINVAL, "w must be compact float32 [D,N] with N in {1,2}");
  *N = (int)w->shape[1];
  in->vec1 = vec_ok(in->x1, in->s1);
  in->vec2 = vec_ok(in->x2, in->s2);
  return KON_OK;
}

}  // namespace
}  // namespace kon

using namespace kon;

#define KON_HEAD_DISPATCH(NJ_, N_, CALL)                         \
  do {                                                           \
    if (N_ == 1) {                                               \
      switch (NJ_) { case 1: CALL(1, 1); break; case 2: CALL(2, 1); break; case 4: CALL(4, 1); break; \
                     default: CALL(8, 1); break; }               \
    } else {                                                     \
      switch (NJ_) { case 1: CALL(1, 2); break; case 2: CALL(2, 2); break; case 4: CALL(4, 2); break; \
                     default: CALL(8, 2); break; }               \
    }                                                            \
  } while (0)

extern "C" int kon_head_fwd(const DLTensor* x1, const DLTensor* x2, const DLTensor* w, const DLTensor* b,
                            DLTensor* y, void* stream) {
  HeadIn in;
  int64_t B;
  int N;
  KON_TRY(head_check(x1, x2, w, &in, &B, &N));
  const int dev = x1->device.device_id;
  KON_TRY(check_cuda_tensor(y, "y", dev));
  KON_REQUIRE(is_f32(y) && y->ndim == 2 && y->shape[0] == B && y->shape[1] == N && is_compact(y), KON_EINVAL,
              "y must be compact float32 [B,N]");
  const float* bp = nullptr;
  if (b) {
    KON_TRY(check_cuda_tensor(b, "b", dev));
    KON_REQUIRE(is_f32(b) && numel(b) == N && is_compact(b), KON_EINVAL, "b must be float32 [N]");
    bp = data_ptr<float>(b);
  }
  if (B == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nj = head_nj(in.D);
  const size_t smem = (size_t)N * nj * 128 * sizeof(float);
  const int grid = head_grid(B, sm_count_of(dev));
  ProfileScope ps("head_fwd_kernel", st);
#define CALL(NJ, NN) \
  head_fwd_kernel<NJ, NN><<<grid, kHeadThreads, smem, st>>>(in, data_ptr<float>(w), bp, data_ptr<float>(y), B)
  KON_HEAD_DISPATCH(nj, N, CALL);
#undef CALL
  KON_LAUNCH_CHECK("head_fwd_kernel");
  return KON_OK;
}

extern "C" size_t kon_head_bwd_workspace_bytes(int64_t batch, int32_t dim, int32_t units, int device_id) {
  const int grid = head_grid(batch, sm_count_of(device_id));
  return (size_t)grid * ((size_t)units * head_nj(dim) * 128 + units) * sizeof(float);
}

extern "C" int kon_head_bwd(const DLTensor* x1, const DLTensor* x2, const DLTensor* w, const DLTensor* gy,
                            DLTensor* dx1, DLTensor* dx2, DLTensor* dw, DLTensor* db, DLTensor* workspace,
                            void* stream) {
  HeadIn in;
  int64_t B;
  int N;
  KON_TRY(head_check(x1, x2, w, &in, &B, &N));
  const int dev = x1->device.device_id;
  KON_TRY(check_cuda_tensor(gy, "gy", dev));
  KON_TRY(check_cuda_tensor(dw, "dw", dev));
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  KON_REQUIRE(is_f32(gy) && gy->ndim == 2 && gy->shape[0] == B && gy->shape[1] == N && is_compact(gy), KON_EINVAL,
              "gy must be compact float32 [B,N]");
  KON_REQUIRE(is_f32(dw) && numel(dw) == (int64_t)in.D * N && is_compact(dw), KON_EINVAL,
              "dw must be compact float32 [D,N]");
  HeadOut out{};
  out.D1 = in.D1;
  out.D = in.D;
  if (dx1) {
    KON_TRY(check_cuda_tensor(dx1, "dx1", dev));
    KON_REQUIRE(is_f32(dx1) && dx1->ndim == 2 && dx1->shape[0] == B && dx1->shape[1] == in.D1 &&
                    stride_of(dx1, 1) == 1,
                KON_EINVAL, "dx1 must be float32 [B,D1] with a compact last dim");
    out.x1 = data_ptr<float>(dx1);
    out.s1 = stride_of(dx1, 0);
    out.vec1 = vec_ok(out.x1, out.s1);
  }
  if (dx2) {
    KON_REQUIRE(x2 != nullptr, KON_EINVAL, "dx2 without x2");
    KON_TRY(check_cuda_tensor(dx2, "dx2", dev));
    KON_REQUIRE(is_f32(dx2) && dx2->ndim == 2 && dx2->shape[0] == B && dx2->shape[1] == in.D - in.D1 &&
                    stride_of(dx2, 1) == 1,
                KON_EINVAL, "dx2 must be float32 [B,D2] with a compact last dim");
    out.x2 = data_ptr<float>(dx2);
    out.s2 = stride_of(dx2, 0);
    out.vec2 = vec_ok(out.x2, out.s2);
  }
  float* dbp = nullptr;
  if (db) {
    KON_TRY(check_cuda_tensor(db, "db", dev));
    KON_REQUIRE(is_f32(db) && numel(db) == N && is_compact(db), KON_EINVAL, "db must be float32 [N]");
    dbp = data_ptr<float>(db);
  }
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nj = head_nj(in.D), Dp = nj * 128;
  const int grid = head_grid(B, sm_count_of(dev));
  const int width = N * Dp + N;
  const size_t need = (size_t)grid * width * sizeof(float);
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= need, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), need);
  float* partial = data_ptr<float>(workspace);
  KON_REQUIRE(((uintptr_t)partial & 15u) == 0, KON_EINVAL, "workspace must be 16-B aligned");
  const size_t smem = ((size_t)2 * N * Dp + N) * sizeof(float);
  ProfileScope ps("head_bwd_kernels", st);
#define CALL(NJ, NN)                                                                                        \
  head_bwd_kernel<NJ, NN><<<grid, kHeadThreads, smem, st>>>(in, out, data_ptr<float>(w), data_ptr<float>(gy), \
                                                            partial, B)
  KON_HEAD_DISPATCH(nj, N, CALL);
#undef CALL
  KON_LAUNCH_CHECK("head_bwd_kernel");
  head_reduce_kernel<<<(width + 31) / 32, 256, 0, st>>>(partial, grid, width, N, Dp, in.D, data_ptr<float>(dw), dbp);
  KON_LAUNCH_CHECK("head_reduce_kernel");
  return KON_OK;
}
