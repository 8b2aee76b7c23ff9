// a1-a3: multi-table embedding gather / bag-sum (forward) + the row-wise sparse optimizers (f1).
// The backward (a4: routing sort + segmented sum) lives in embed_bwd.cu.
//
// Replaces SparseEmbed.call (IL:225-242: 26 Keras Embedding gathers, Flatten, optional
// Add over fields) and SeqBaseLayer.call (BL:45-46: sum over the bag axis).
//
// HBM-bound integer routing + fp32 payload.  Design:
//   forward : persistent CTAs; the id tile of a CTA iteration is staged in shared memory
//             with a 1-D bulk async copy (TMA engine, mbarrier completion), double
//             buffered, so the dependent chain id -> row address never stalls a warp on
//             global memory; LPR lanes x 128-bit cover one row, 4 rows in flight per
//             thread; stores are fully coalesced ([B,F,dim] bags are consecutive rows).
#include "common.cuh"
#include "embed_common.cuh"

namespace kon {

#ifndef KON_EMB_U
#define KON_EMB_U 4
#endif
constexpr int kFwdThreads = 256;
#ifndef KON_EMB_TILE
#define KON_EMB_TILE 512
#endif
#ifndef KON_EMB_CTAS
#define KON_EMB_CTAS 4
#endif
constexpr int kTileIds = KON_EMB_TILE;   // ids staged per CTA iteration (per buffer): small tiles keep the
                                         // persistent CTAs balanced (832 tiles of 2048 over 592 resident CTAs = 2 uneven waves)

// Sharded embeddings over NVLink peer memory: the batch axis of `out` / `d_out` is
// split in `rows`-sample slabs, slab q living in the memory of rank q (base[q], mapped into this
// process with cudaIpcOpenMemHandle).  n == 0: the ordinary single-buffer call.
struct PeerTable {
  float* base[kMaxPeers];
  long long rows;       // samples per peer slab
  int n;                // peers (0 = off)
  int skip_invalid;     // forward: leave the destination untouched for out-of-range ids (row-wise
                        // shards: the id belongs to another rank, which writes the row itself)
  int use_col;          // forward: local field f lands at column col[f] (floats) of the destination row instead
  int col[kMaxFields];  // of f * out_stride_f -- owners whose fields are not adjacent in the model's field order
};

// -------------------------------------------------------------------------------------
// forward, vector path: dim % 4 == 0, 16-B aligned rows
// -------------------------------------------------------------------------------------
template <typename IdT, int LPR, bool PEER>
__global__ void __launch_bounds__(kFwdThreads)
embed_fwd_vec_kernel(const float4* __restrict__ arena, const IdT* __restrict__ ids,
                     const __grid_constant__ FieldTable ft, int F, int L, int vec_per_row,
                     long long n_bags, int bags_per_tile, float* __restrict__ out,
                     long long out_sb, long long out_sf, int* __restrict__ oob, unsigned int f_magic,
                     const __grid_constant__ PeerTable pt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ long long s_off[kMaxFields + 1];

  IdT* sid[2];
  const int tile_ids = bags_per_tile * L;
  sid[0] = reinterpret_cast<IdT*>(smem_raw);
  sid[1] = sid[0] + tile_ids;

  const int tid = threadIdx.x;
  for (int i = tid; i <= F; i += kFwdThreads) s_off[i] = ft.off[i];
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const long long n_tiles = (n_bags + bags_per_tile - 1) / bags_per_tile;
  constexpr int G = kFwdThreads / LPR;   // lane groups per CTA
  const int g = tid / LPR;
  const int lane = tid % LPR;
  const bool lane_on = lane < vec_per_row;

  // Stage tile `t` into buffer `buf`.  Full tiles use the bulk async engine; a ragged last
  // tile (byte count not a multiple of 16) is copied by the threads themselves.
  auto stage = [&](long long t, int buf) {
    const long long bag0 = t * bags_per_tile;
    const long long nb = min((long long)bags_per_tile, n_bags - bag0);
    const uint32_t bytes = (uint32_t)(nb * L * sizeof(IdT));
    const IdT* src = ids + bag0 * L;
    if ((bytes & 15u) == 0) {
      if (tid == 0) {
        mbar_expect_tx(&bar[buf], bytes);
        bulk_g2s(sid[buf], src, bytes, &bar[buf]);
      }
    } else {
      for (int i = tid; i < (int)(nb * L); i += kFwdThreads) sid[buf][i] = src[i];
      __syncthreads();
      if (tid == 0) mbar_arrive(&bar[buf]);
    }
  };

  long long t = blockIdx.x;
  if (t < n_tiles) stage(t, 0);
  uint32_t phase[2] = {0, 0};
  int buf = 0;
  for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
    const long long tn = t + gridDim.x;
    if (tn < n_tiles) stage(tn, buf ^ 1);   // buffer buf^1 was released by the barrier below
    mbar_wait(&bar[buf], phase[buf]);
    phase[buf] ^= 1;

    const long long bag0 = t * bags_per_tile;
    const int nb = (int)min((long long)bags_per_tile, n_bags - bag0);
    const long long b0 = bag0 / F;
    const int f0 = (int)(bag0 - b0 * F);
    const IdT* s = sid[buf];

    if (L == 1) {
      constexpr int U = KON_EMB_U;
      // jj / F by multiplication: jj < tile + F < 65536, f_magic = ceil(2^24 / F) -> exact
      auto divF = [&](int jj) { return (int)(((unsigned long long)(unsigned)jj * f_magic) >> 24); };
      for (int j0 = g; j0 < nb; j0 += G * U) {
        float4 r[U];
        bool ok[U];
        int fs[U], qs[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * G;
          ok[u] = false;
          r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          fs[u] = qs[u] = 0;
          if (j < nb) {
            const int jj = f0 + j;
            qs[u] = divF(jj);
            const int f = jj - qs[u] * F;
            fs[u] = f;
            const long long id = (long long)s[j];
            const long long rows = s_off[f + 1] - s_off[f];
            ok[u] = true;
            if (id >= 0 && id < rows) {
              if (lane_on) r[u] = ldg_stream_f4(arena + (s_off[f] + id) * vec_per_row + lane);
            } else {
              if (PEER && pt.skip_invalid) ok[u] = false;
              else if (lane == 0 && oob) atomicAdd(oob, 1);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * G;
          if (ok[u] && lane_on) {
            (void)j;
            long long b = b0 + qs[u];
            float* dst = out;
            if (PEER) {   // slab q of the batch axis lives on rank q: 16-B stores straight over NVLink
              const long long q = b / pt.rows;
              dst = pt.base[q];
              b -= q * pt.rows;
            }
            const long long fcol = (PEER && pt.use_col) ? (long long)pt.col[fs[u]] : fs[u] * out_sf;
            *reinterpret_cast<float4*>(dst + b * out_sb + fcol + lane * 4) = r[u];
          }
        }
      }
    } else {
      for (int j = g; j < nb; j += G) {
        const int jj = f0 + j;
        const long long b = b0 + jj / F;
        const int f = jj % F;
        const long long base = s_off[f];
        const long long rows = s_off[f + 1] - base;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const IdT* sb = s + (long long)j * L;
        for (int l0 = 0; l0 < L; l0 += 4) {
          float4 r[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (l0 + u < L) {
              const long long id = (long long)sb[l0 + u];
              if (id >= 0 && id < rows) {
                if (lane_on) r[u] = ldg_stream_f4(arena + (base + id) * vec_per_row + lane);
              } else if (lane == 0 && oob) {
                atomicAdd(oob, 1);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {   // order l = 0..L-1 (BL:46)
            acc.x += r[u].x; acc.y += r[u].y; acc.z += r[u].z; acc.w += r[u].w;
          }
        }
        if (lane_on) *reinterpret_cast<float4*>(out + b * out_sb + f * out_sf + lane * 4) = acc;
      }
    }
    __syncthreads();   // everyone is done with sid[buf] before it is restaged
  }
}

// -------------------------------------------------------------------------------------
// forward, scalar path: any dim (the dim-1 linear tables, IL:219-222), optional sum over
// fields in order f = 0..F-1 (Keras Add, IL:233-234)
// -------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(256)
embed_fwd_scalar_kernel(const float* __restrict__ arena, const IdT* __restrict__ ids,
                        const __grid_constant__ FieldTable ft, int F, int L, int dim,
                        long long n_out_rows, float* __restrict__ out, long long out_sb,
                        long long out_sf, int sum_fields, int* __restrict__ oob) {
  __shared__ long long s_off[kMaxFields + 1];
  for (int i = threadIdx.x; i <= F; i += blockDim.x) s_off[i] = ft.off[i];
  __syncthreads();
  const long long total = n_out_rows * dim;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long orow = idx / dim;
    const int col = (int)(idx - orow * dim);
    auto bag_sum = [&](long long b, int f) {
      const IdT* p = ids + (b * F + f) * (long long)L;
      const long long base = s_off[f];
      const long long rows = s_off[f + 1] - base;
      float acc = 0.f;
      for (int l = 0; l < L; ++l) {
        const long long id = (long long)p[l];
        float v = 0.f;
        if (id >= 0 && id < rows) v = __ldg(arena + (base + id) * dim + col);
        else if (col == 0 && oob) atomicAdd(oob, 1);
        acc = (l == 0) ? v : acc + v;
      }
      return acc;
    };
    if (sum_fields) {
      float tot = bag_sum(orow, 0);
      for (int f = 1; f < F; ++f) tot += bag_sum(orow, f);
      out[orow * out_sb + col] = tot;
    } else {
      const long long b = orow / F;
      const int f = (int)(orow - b * F);
      out[b * out_sb + f * out_sf + col] = bag_sum(b, f);
    }
  }
}

// -------------------------------------------------------------------------------------
// sparse optimizers
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_sgd_kernel(float* __restrict__ arena, const int* __restrict__ rows,
                 const float* __restrict__ grads, const int* __restrict__ n_unique, int dim,
                 float lr, float l2) {
  const long long total = (long long)(*n_unique) * dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long u = i / dim;
    const int c = (int)(i - u * dim);
    float* w = arena + (long long)rows[u] * dim + c;
    const float wv = *w;
    *w = wv - lr * (grads[i] + 2.f * l2 * wv);
  }
}

// VEC = 4: dim % 4 == 0, one thread per 128-bit group (rows are 16-B aligned); VEC = 1: any dim.
// 32-bit index math (n_unique * dim < 2^31 is checked by the host), shift instead of a division when
// the groups-per-row count is a power of two.
template <int VEC>
__global__ void __launch_bounds__(256)
embed_adam_kernel(float* __restrict__ arena, float* __restrict__ m, float* __restrict__ v,
                  const int* __restrict__ rows, const float* __restrict__ grads,
                  const int* __restrict__ n_unique, int dim, float lr_t, float b1, float b2,
                  float eps, float l2, const int* __restrict__ step_dev, float lr) {
  if (step_dev) {   // step counter lives on the device (CUDA-graph replay): bias correction here
    const float t = (float)(*step_dev);
    lr_t = lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
  }
  const unsigned gpr = (unsigned)(dim / VEC);                 // groups per row
  const int shift = (gpr & (gpr - 1)) == 0 ? __ffs(gpr) - 1 : -1;
  const unsigned total = (unsigned)(*n_unique) * gpr;
  const float c1 = 1.f - b1, c2 = 1.f - b2, tl2 = 2.f * l2;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned u = shift >= 0 ? (i >> shift) : i / gpr;
    const unsigned c = i - u * gpr;
    const size_t o = (size_t)rows[u] * gpr + c;
    if (VEC == 4) {
      float4 wv = reinterpret_cast<float4*>(arena)[o];
      float4 mv = reinterpret_cast<float4*>(m)[o];
      float4 vv = reinterpret_cast<float4*>(v)[o];
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(grads) + i);
      float* w_ = &wv.x; float* m_ = &mv.x; float* v_ = &vv.x; const float* g_ = &g4.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float gq = g_[e] + tl2 * w_[e];
        m_[e] = b1 * m_[e] + c1 * gq;
        v_[e] = b2 * v_[e] + c2 * gq * gq;
        w_[e] = w_[e] - lr_t * m_[e] / (sqrtf(v_[e]) + eps);
      }
      reinterpret_cast<float4*>(m)[o] = mv;
      reinterpret_cast<float4*>(v)[o] = vv;
      reinterpret_cast<float4*>(arena)[o] = wv;
    } else {
      const float wv = arena[o];
      const float gq = grads[i] + tl2 * wv;
      const float mn = b1 * m[o] + c1 * gq;
      const float vn = b2 * v[o] + c2 * gq * gq;
      m[o] = mn;
      v[o] = vn;
      arena[o] = wv - lr_t * mn / (sqrtf(vn) + eps);
    }
  }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
template <typename IdT, bool PEER>
static int launch_fwd_vec(int lpr, int grid, size_t smem, cudaStream_t st, const float4* arena,
                          const IdT* ids, const FieldTable& ft, int F, int L, int vpr,
                          long long n_bags, int bpt, float* out, long long sb, long long sf,
                          int* oob, const PeerTable& pt) {
#define KON_FWD_CASE(N)                                                                         \
  case N:                                                                                       \
    embed_fwd_vec_kernel<IdT, N, PEER><<<grid, kFwdThreads, smem, st>>>(                        \
        arena, ids, ft, F, L, vpr, n_bags, bpt, out, sb, sf, oob,                               \
        (unsigned)(((1u << 24) + F - 1) / F), pt);                                              \
    break;
  ProfileScope ps(PEER ? "embed_fwd_peer_kernel" : "embed_fwd_vec_kernel", st);
  switch (lpr) {
    KON_FWD_CASE(1)
    KON_FWD_CASE(2)
    KON_FWD_CASE(4)
    KON_FWD_CASE(8)
    KON_FWD_CASE(16)
    KON_FWD_CASE(32)
    default:
      return fail(KON_EUNSUPPORTED, "embedding dim too large for the vector path");
  }
#undef KON_FWD_CASE
  KON_LAUNCH_CHECK("embed_fwd_vec_kernel");
  return KON_OK;
}

}  // namespace kon

using namespace kon;

extern "C" int kon_embed_fwd(const DLTensor* arena, const DLTensor* ids,
                             const int64_t* field_row_offset, int32_t n_fields, DLTensor* out,
                             DLTensor* oob, int32_t flags, void* stream) {
  KON_TRY(check_cuda_tensor(arena, "arena"));
  const int dev = arena->device.device_id;
  IdsView v;
  FieldTable ft;
  KON_TRY(parse_common(ids, field_row_offset, n_fields, dev, &v, &ft));
  KON_TRY(check_cuda_tensor(out, "out", dev));
  KON_REQUIRE(is_f32(arena) && arena->ndim == 2 && is_compact(arena), KON_EINVAL,
              "arena must be compact [R,dim] float32");
  KON_REQUIRE(ft.off[n_fields] <= arena->shape[0], KON_EINVAL,
              "field_row_offset[F]=%lld exceeds arena rows %lld", (long long)ft.off[n_fields],
              (long long)arena->shape[0]);
  const int64_t dim = arena->shape[1];
  const bool sum_fields = (flags & KON_EMBED_SUM_FIELDS) != 0;
  KON_REQUIRE(is_f32(out), KON_EINVAL, "out must be float32");
  long long sb, sf;
  if (sum_fields) {
    KON_REQUIRE(out->ndim == 2 && out->shape[0] == v.B && out->shape[1] == dim &&
                    stride_of(out, 1) == 1,
                KON_EINVAL, "with KON_EMBED_SUM_FIELDS out must be [B,dim]");
    sb = stride_of(out, 0);
    sf = 0;
  } else {
    KON_REQUIRE(out->ndim == 3 && out->shape[0] == v.B && out->shape[1] == v.F &&
                    out->shape[2] == dim && (dim == 1 || stride_of(out, 2) == 1),
                KON_EINVAL, "out must be [B,F,dim] with a compact last dim");
    sb = stride_of(out, 0);
    sf = stride_of(out, 1);
  }
  int* oob_p = nullptr;
  if (oob) {
    KON_TRY(check_cuda_tensor(oob, "oob", dev));
    KON_REQUIRE(is_i32(oob) && numel(oob) >= 1, KON_EINVAL, "oob must be int32[1]");
    oob_p = data_ptr<int>(oob);
  }
  if (v.B == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = sm_count_of(dev);
  float* outp = data_ptr<float>(out);
  const float* ap = data_ptr<float>(arena);

  const bool vec_ok = !sum_fields && dim % 4 == 0 && dim <= 128 && aligned16(ap) &&
                      aligned16(outp) && sb % 4 == 0 && sf % 4 == 0 &&
                      aligned16(data_ptr<char>(ids));
  if (vec_ok) {
    const int vpr = (int)(dim / 4);
    const int lpr = pow2_ge(vpr);
    int bpt = (int)((kTileIds / v.L) & ~3LL);
    if (bpt < 4) bpt = 4;
    const long long n_bags = v.B * v.F;
    const long long n_tiles = (n_bags + bpt - 1) / bpt;
    const size_t idsz = v.i64 ? 8 : 4;
    const size_t smem = 2 * (size_t)bpt * v.L * idsz;
    int grid = (int)std::min<long long>(n_tiles, (long long)sms * KON_EMB_CTAS);   // = resident CTAs: one wave, tiles strided
    PeerTable pt{};
    if (v.i64)
      return launch_fwd_vec<long long, false>(lpr, grid, smem, st, reinterpret_cast<const float4*>(ap),
                                              data_ptr<long long>(ids), ft, (int)v.F, (int)v.L, vpr,
                                              n_bags, bpt, outp, sb, sf, oob_p, pt);
    return launch_fwd_vec<int, false>(lpr, grid, smem, st, reinterpret_cast<const float4*>(ap),
                                      data_ptr<int>(ids), ft, (int)v.F, (int)v.L, vpr, n_bags, bpt,
                                      outp, sb, sf, oob_p, pt);
  }
  const long long n_out_rows = sum_fields ? v.B : v.B * v.F;
  const long long total = n_out_rows * dim;
  int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 16);
  if (v.i64)
    embed_fwd_scalar_kernel<long long><<<grid, 256, 0, st>>>(
        ap, data_ptr<long long>(ids), ft, (int)v.F, (int)v.L, (int)dim, n_out_rows, outp, sb, sf,
        sum_fields ? 1 : 0, oob_p);
  else
    embed_fwd_scalar_kernel<int><<<grid, 256, 0, st>>>(ap, data_ptr<int>(ids), ft, (int)v.F,
                                                       (int)v.L, (int)dim, n_out_rows, outp, sb,
                                                       sf, sum_fields ? 1 : 0, oob_p);
  KON_LAUNCH_CHECK("embed_fwd_scalar_kernel");
  return KON_OK;
}

// Sharded forward over peer memory: this rank gathers the rows of ITS tables for the GLOBAL batch
// and stores each row straight into the concat buffer of the rank that owns the sample.
extern "C" int kon_embed_fwd_peer(const DLTensor* arena, const DLTensor* ids,
                                  const int64_t* field_row_offset, int32_t n_fields,
                                  void* const* peer_out, int32_t n_peers, int64_t rows_per_peer,
                                  int64_t out_stride_b, int64_t out_stride_f, DLTensor* oob,
                                  int32_t flags, void* stream) {
  return kon_embed_fwd_peer_cols(arena, ids, field_row_offset, n_fields, peer_out, n_peers, rows_per_peer,
                                 out_stride_b, out_stride_f, nullptr, oob, flags, stream);
}

extern "C" int kon_embed_fwd_peer_cols(const DLTensor* arena, const DLTensor* ids,
                                       const int64_t* field_row_offset, int32_t n_fields,
                                       void* const* peer_out, int32_t n_peers, int64_t rows_per_peer,
                                       int64_t out_stride_b, int64_t out_stride_f, const int32_t* field_col,
                                       DLTensor* oob, int32_t flags, void* stream) {
  KON_TRY(check_cuda_tensor(arena, "arena"));
  const int dev = arena->device.device_id;
  IdsView v;
  FieldTable ft;
  KON_TRY(parse_common(ids, field_row_offset, n_fields, dev, &v, &ft));
  KON_REQUIRE(is_f32(arena) && arena->ndim == 2 && is_compact(arena), KON_EINVAL,
              "arena must be compact [R,dim] float32");
  KON_REQUIRE(ft.off[n_fields] <= arena->shape[0], KON_EINVAL,
              "field_row_offset[F]=%lld exceeds arena rows %lld", (long long)ft.off[n_fields],
              (long long)arena->shape[0]);
  KON_REQUIRE(peer_out != nullptr && n_peers >= 1 && n_peers <= kMaxPeers, KON_EINVAL,
              "n_peers=%d outside [1,%d]", n_peers, kMaxPeers);
  KON_REQUIRE(rows_per_peer >= 1 && v.B <= rows_per_peer * n_peers, KON_EINVAL,
              "ids has %lld samples, peers hold %lld x %d", (long long)v.B, (long long)rows_per_peer,
              n_peers);
  KON_REQUIRE(v.L == 1, KON_EUNSUPPORTED, "the peer exchange takes [B,F] ids (one id per field)");
  const int64_t dim = arena->shape[1];
  KON_REQUIRE(dim % 4 == 0 && dim <= 128 && aligned16(data_ptr<float>(arena)) &&
                  out_stride_b % 4 == 0 && out_stride_f % 4 == 0 && aligned16(data_ptr<char>(ids)),
              KON_EUNSUPPORTED, "the peer exchange needs dim %% 4 == 0 and 16-B aligned rows");
  PeerTable pt{};
  pt.n = n_peers;
  pt.rows = rows_per_peer;
  pt.skip_invalid = (flags & KON_EMBED_SKIP_INVALID) ? 1 : 0;
  for (int q = 0; q < n_peers; ++q) {
    KON_REQUIRE(peer_out[q] != nullptr && aligned16(peer_out[q]), KON_EINVAL,
                "peer_out[%d] is NULL or not 16-B aligned", q);
    pt.base[q] = static_cast<float*>(peer_out[q]);
  }
  if (field_col) {
    pt.use_col = 1;
    for (int f = 0; f < n_fields; ++f) {
      KON_REQUIRE(field_col[f] >= 0 && field_col[f] % 4 == 0, KON_EINVAL, "field_col[%d]=%d must be a non-negative multiple of 4 floats", f, field_col[f]);
      pt.col[f] = field_col[f];
    }
  }
  int* oob_p = nullptr;
  if (oob) {
    KON_TRY(check_cuda_tensor(oob, "oob", dev));
    KON_REQUIRE(is_i32(oob) && numel(oob) >= 1, KON_EINVAL, "oob must be int32[1]");
    oob_p = data_ptr<int>(oob);
  }
  if (v.B == 0) return KON_OK;
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int sms = sm_count_of(dev);
  const int vpr = (int)(dim / 4);
  const int lpr = pow2_ge(vpr);
  const int bpt = (int)(kTileIds & ~3);
  const long long n_bags = v.B * v.F;
  const long long n_tiles = (n_bags + bpt - 1) / bpt;
  const size_t smem = 2 * (size_t)bpt * (v.i64 ? 8 : 4);
  const int grid = (int)std::min<long long>(n_tiles, (long long)sms * KON_EMB_CTAS);
  const float4* ap = reinterpret_cast<const float4*>(data_ptr<float>(arena));
  if (v.i64)
    return launch_fwd_vec<long long, true>(lpr, grid, smem, st, ap, data_ptr<long long>(ids), ft,
                                           (int)v.F, 1, vpr, n_bags, bpt, nullptr, out_stride_b,
                                           out_stride_f, oob_p, pt);
  return launch_fwd_vec<int, true>(lpr, grid, smem, st, ap, data_ptr<int>(ids), ft, (int)v.F, 1, vpr,
                                   n_bags, bpt, nullptr, out_stride_b, out_stride_f, oob_p, pt);
}

static int check_sparse_update(const DLTensor* arena, const DLTensor* unique_rows,
                               const DLTensor* grads, const DLTensor* n_unique) {
  KON_TRY(check_cuda_tensor(arena, "arena"));
  const int dev = arena->device.device_id;
  KON_TRY(check_cuda_tensor(unique_rows, "unique_rows", dev));
  KON_TRY(check_cuda_tensor(grads, "grads", dev));
  KON_TRY(check_cuda_tensor(n_unique, "n_unique", dev));
  KON_REQUIRE(is_f32(arena) && arena->ndim == 2 && is_compact(arena), KON_EINVAL,
              "arena must be compact float32 [R,dim]");
  KON_REQUIRE(is_f32(grads) && grads->ndim == 2 && grads->shape[1] == arena->shape[1] &&
                  is_compact(grads),
              KON_EINVAL, "grads must be compact float32 [N,dim]");
  KON_REQUIRE(is_i32(unique_rows) && is_compact(unique_rows) &&
                  numel(unique_rows) >= grads->shape[0],
              KON_EINVAL, "unique_rows must be compact int32 [>=N]");
  KON_REQUIRE(is_i32(n_unique), KON_EINVAL, "n_unique must be int32[1]");
  return KON_OK;
}

extern "C" int kon_embed_sgd(DLTensor* arena, const DLTensor* unique_rows, const DLTensor* grads,
                             const DLTensor* n_unique, float lr, float l2, void* stream) {
  KON_TRY(check_sparse_update(arena, unique_rows, grads, n_unique));
  const int dev = arena->device.device_id;
  DeviceGuard guard(dev);
  const long long total = grads->shape[0] * grads->shape[1];
  if (total == 0) return KON_OK;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count_of(dev) * 16);
  embed_sgd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      data_ptr<float>(arena), data_ptr<int>(unique_rows), data_ptr<float>(grads),
      data_ptr<int>(n_unique), (int)arena->shape[1], lr, l2);
  KON_LAUNCH_CHECK("embed_sgd_kernel");
  return KON_OK;
}

static int embed_adam_impl(DLTensor* arena, DLTensor* m, DLTensor* v, const DLTensor* unique_rows,
                           const DLTensor* grads, const DLTensor* n_unique, float lr, float beta1,
                           float beta2, float eps, float l2, int32_t step, const DLTensor* step_dev,
                           void* stream);

extern "C" int kon_embed_adam(DLTensor* arena, DLTensor* m, DLTensor* v,
                              const DLTensor* unique_rows, const DLTensor* grads,
                              const DLTensor* n_unique, float lr, float beta1, float beta2,
                              float eps, float l2, int32_t step, void* stream) {
  return embed_adam_impl(arena, m, v, unique_rows, grads, n_unique, lr, beta1, beta2, eps, l2, step, nullptr,
                         stream);
}

extern "C" int kon_embed_adam_devstep(DLTensor* arena, DLTensor* m, DLTensor* v,
                                      const DLTensor* unique_rows, const DLTensor* grads,
                                      const DLTensor* n_unique, float lr, float beta1, float beta2,
                                      float eps, float l2, const DLTensor* step, void* stream) {
  KON_TRY(check_cuda_tensor(step, "step"));
  KON_REQUIRE(is_i32(step) && numel(step) >= 1, KON_EINVAL, "step must be a device int32[1]");
  return embed_adam_impl(arena, m, v, unique_rows, grads, n_unique, lr, beta1, beta2, eps, l2, 1, step, stream);
}

static int embed_adam_impl(DLTensor* arena, DLTensor* m, DLTensor* v, const DLTensor* unique_rows,
                           const DLTensor* grads, const DLTensor* n_unique, float lr, float beta1,
                           float beta2, float eps, float l2, int32_t step, const DLTensor* step_dev,
                           void* stream) {
  KON_TRY(check_sparse_update(arena, unique_rows, grads, n_unique));
  const int dev = arena->device.device_id;
  KON_TRY(check_cuda_tensor(m, "m", dev));
  KON_TRY(check_cuda_tensor(v, "v", dev));
  KON_REQUIRE(is_f32(m) && is_f32(v) && numel(m) == numel(arena) && numel(v) == numel(arena) &&
                  is_compact(m) && is_compact(v),
              KON_EINVAL, "m and v must be compact float32 like arena");
  KON_REQUIRE(step >= 1, KON_EINVAL, "step must be >= 1");
  DeviceGuard guard(dev);
  const long long total = grads->shape[0] * grads->shape[1];
  if (total == 0) return KON_OK;
  const float lr_t = lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
  KON_REQUIRE(total < 0x7fffffffLL, KON_EUNSUPPORTED, "more than 2^31-1 elements in one sparse update");
  const int dim_i = (int)arena->shape[1];
  const bool vec = dim_i % 4 == 0 && aligned16(data_ptr<float>(arena)) && aligned16(data_ptr<float>(m)) &&
                   aligned16(data_ptr<float>(v)) && aligned16(data_ptr<float>(grads));
  const long long work = vec ? total / 4 : total;
  const int grid = (int)std::max<long long>(1, std::min<long long>((work + 255) / 256, (long long)sm_count_of(dev) * 16));
#define KON_ADAM_ARGS                                                                               \
  data_ptr<float>(arena), data_ptr<float>(m), data_ptr<float>(v), data_ptr<int>(unique_rows),       \
      data_ptr<float>(grads), data_ptr<int>(n_unique), dim_i, lr_t, beta1, beta2, eps, l2,          \
      step_dev ? data_ptr<int>(step_dev) : nullptr, lr
  if (vec)
    embed_adam_kernel<4><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(KON_ADAM_ARGS);
  else
    embed_adam_kernel<1><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(KON_ADAM_ARGS);
#undef KON_ADAM_ARGS
  KON_LAUNCH_CHECK("embed_adam_kernel");
  return KON_OK;
}
