// a1-a4: multi-table embedding gather / bag-sum (forward) and deterministic
// sort-then-segment scatter-add (backward) + sparse row optimizers.
//
// Replaces SparseEmbed.call (IL:225-242: 26 Keras Embedding gathers, Flatten, optional
// Add over fields), SeqBaseLayer.call (BL:45-46: sum over the bag axis) and the implicit
// IndexedSlices -> unique -> unsorted_segment_sum of Model.fit.
//
// HBM-bound integer routing + fp32 payload.  Design:
//   forward : persistent CTAs; the id tile of a CTA iteration is staged in shared memory
//             with a 1-D bulk async copy (TMA engine, mbarrier completion), double
//             buffered, so the dependent chain id -> row address never stalls a warp on
//             global memory; LPR lanes x 128-bit cover one row, 4 rows in flight per
//             thread; stores are fully coalesced ([B,F,dim] bags are consecutive rows).
//   backward: key = arena row (uint32), value = lookup position; LSD radix sort on the
//             significant bits only; run heads by inclusive scan; a two-level windowed
//             segmented reduction (16 lookups per lane group, 64 groups per CTA) whose
//             association order depends only on the sorted positions -> bit-reproducible.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

namespace kon {

#ifndef KON_EMB_U
#define KON_EMB_U 4
#endif
#ifndef KON_EMB_WIN
#define KON_EMB_WIN 16
#endif
#ifndef KON_EMB_RED_MINB_LIN
#define KON_EMB_RED_MINB_LIN 3   // fused first-order gradient (kon_embed_bwd_pair)
#endif
#ifndef KON_EMB_RED_NB_LIN
#define KON_EMB_RED_NB_LIN 4     // 8 spills (80-register cap at 3 CTAs/SM): 91 us instead of 61 us at 1.7 M lookups
#endif
#ifndef KON_EMB_RED_NB
#define KON_EMB_RED_NB 8
#endif
#ifndef KON_EMB_RED_MINB
#define KON_EMB_RED_MINB 3
#endif
constexpr int kMaxFields = 256;
constexpr int kFwdThreads = 256;
#ifndef KON_EMB_TILE
#define KON_EMB_TILE 512
#endif
#ifndef KON_EMB_CTAS
#define KON_EMB_CTAS 4
#endif
constexpr int kTileIds = KON_EMB_TILE;   // ids staged per CTA iteration (per buffer): small tiles keep the
                                         // persistent CTAs balanced (832 tiles of 2048 over 592 resident CTAs = 2 uneven waves)
constexpr int kMaxBagLen = 512;

struct FieldTable {
  int64_t off[kMaxFields + 1];
};

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
// backward
// -------------------------------------------------------------------------------------
template <typename IdT>
__global__ void __launch_bounds__(256)
embed_keys_kernel(const IdT* __restrict__ ids, const __grid_constant__ FieldTable ft, int F, int L,
                  long long n, uint32_t sentinel, uint32_t* __restrict__ keys,
                  uint32_t* __restrict__ vals) {
  __shared__ long long s_off[kMaxFields + 1];
  for (int i = threadIdx.x; i <= F; i += blockDim.x) s_off[i] = ft.off[i];
  __syncthreads();
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n;
       p += (long long)gridDim.x * blockDim.x) {
    const long long bag = p / L;
    const int f = (int)(bag % F);
    const long long id = (long long)ids[p];
    const long long rows = s_off[f + 1] - s_off[f];
    keys[p] = (id >= 0 && id < rows) ? (uint32_t)(s_off[f] + id) : sentinel;
    // the payload of the sort is the (sample, field) pair itself, packed b * 256 + f (F <= 256, B < 2^24: checked
    // by the host), so that the segmented reduction addresses its gradient row with a shift and a mask -- the two
    // 64-bit divisions per lookup it used to spend on p -> (b, f) were most of its 146 instructions per lookup
    vals[p] = (uint32_t)(((bag / F) << 8) | (uint32_t)f);
  }
}

struct RunHead {
  const uint32_t* keys;
  __host__ __device__ int operator()(int i) const {
    return (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
  }
};

constexpr int kWin = KON_EMB_WIN;  // sorted lookups per lane group
constexpr int kRedThreads = 256;
// per-window / per-CTA meta bits
constexpr int kHasHead = 1;        // first run continues a run of the previous window
constexpr int kHeadEnded = 2;      // that run ends inside this window
constexpr int kHasTail = 4;        // last run started here and continues in the next window

struct BwdArgs {
  const float* d_out;
  long long sb, sf;   // strides of d_out dims 0 / 1 (elements)
  int F, L, dim, vec_per_row;
  long long n;        // lookups
  const uint32_t* keys;   // sorted
  const uint32_t* vals;   // sorted with the keys
  const int* segidx;      // inclusive scan of run heads (1-based run id)
  uint32_t sentinel;
  int* unique_rows;
  float* grads;
  int* n_unique;
  float* cta_head;  // [n_cta, dim]
  float* cta_tail;  // [n_cta, dim]
  int* cta_meta;    // [n_cta]
  // fused first-order gradient (kon_embed_bwd_pair): a second, one-float-per-lookup gradient reduced over the
  // same routing in the same pass (the dim-1 "linear" tables are looked up with the same ids)
  const float* d1;
  long long sb1, sf1;
  float* grads1;      // [n_unique]
  float* cta_head1;   // [n_cta]
  float* cta_tail1;   // [n_cta]
  // peer mode (n_peers > 0): sample b's gradient row is read from rank b / peer_rows over NVLink
  int n_peers;
  long long peer_rows;
  const float* peer[kMaxPeers];
};

// One lane group (LPR lanes x float4) reduces one window of kWin sorted lookups; the CTA
// then stitches runs that cross window boundaries through shared memory, and leaves at
// most one head and one tail partial per CTA for embed_fixup_kernel.
// NB = gradient rows loaded per lane group before the first one is consumed.  8 covers HBM latency; rows that
// come over NVLink (peer mode) have ~3x the latency, so that instantiation keeps a whole 16-lookup window in flight.
// LIN: also reduce the one-float gradient a.d1 (see BwdArgs) -- one extra 4-byte load per lookup (the same address
// for the LPR lanes of a group), one extra accumulator.
template <int LPR, int NB, bool LIN = false>
__global__ void __launch_bounds__(kRedThreads, NB > 8 ? 2 : (LIN ? KON_EMB_RED_MINB_LIN : KON_EMB_RED_MINB))
embed_reduce_kernel(const BwdArgs a) {
  constexpr int G = kRedThreads / LPR;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_head = reinterpret_cast<float4*>(smem_raw);          // [G][LPR]
  float4* s_tail = s_head + G * LPR;                             // [G][LPR]
  __shared__ int s_meta[G];
  __shared__ int s_tailseg[G];
  __shared__ uint32_t s_tailkey[G];
  __shared__ float s_head1[LIN ? G : 1];
  __shared__ float s_tail1[LIN ? G : 1];

  const int tid = threadIdx.x;
  const int g = tid / LPR, lane = tid % LPR;
  const bool lane_on = lane < a.vec_per_row;
  const long long cta_lo = (long long)blockIdx.x * G * kWin;
  const long long lo = cta_lo + (long long)g * kWin;
  const long long hi = min(a.n, lo + kWin);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);

  int meta = 0;
  float4 head = zero, tail = zero;
  float head1 = 0.f, tail1 = 0.f;
  int tail_seg = 0;
  uint32_t tail_key = 0;

  auto emit_final = [&](int seg, uint32_t key, float4 v, float v1) {
    if (key == a.sentinel) return;   // out-of-range ids carry no gradient
    if (lane_on)
      *reinterpret_cast<float4*>(a.grads + (long long)(seg - 1) * a.dim + lane * 4) = v;
    if (lane == 0) {
      a.unique_rows[seg - 1] = (int)key;
      if (LIN) a.grads1[seg - 1] = v1;
    }
  };

  if (lo < a.n) {
    const bool starts = (lo == 0) || (a.segidx[lo] != a.segidx[lo - 1]);
    const bool ends = (hi == a.n) || (a.segidx[hi] != a.segidx[hi - 1]);
    if (!starts) meta |= kHasHead;
    int cur = a.segidx[lo];
    uint32_t cur_key = a.keys[lo];
    bool first_run = true;
    float4 acc = zero;
    float acc1 = 0.f;
    const int cnt = (int)(hi - lo);
    for (int i0 = 0; i0 < cnt; i0 += NB) {
      float4 r[NB];
      float r1[LIN ? NB : 1];
      int sg[NB];
      uint32_t ky[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        r[u] = zero;
        if (LIN) r1[u] = 0.f;
        sg[u] = cur;
        ky[u] = cur_key;
        if (i0 + u < cnt) {
          const long long i = lo + i0 + u;
          sg[u] = a.segidx[i];
          ky[u] = a.keys[i];
          const uint32_t bf = a.vals[i];               // b * 256 + f, packed by embed_keys_kernel
          long long b = bf >> 8;
          const int f = (int)(bf & 255u);
          const float* src = a.d_out;
          if (a.n_peers) {
            const long long q = b / a.peer_rows;
            src = a.peer[q];
            b -= q * a.peer_rows;
          }
          if (lane_on)
            r[u] = ldg_stream_f4(
                reinterpret_cast<const float4*>(src + b * a.sb + f * a.sf) + lane);
          if (LIN) r1[u] = __ldg(a.d1 + b * a.sb1 + f * a.sf1);
        }
      }
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        if (i0 + u < cnt) {
          if (sg[u] != cur) {   // run [.., i-1] is complete at its right end
            if (first_run && !starts) {
              head = acc;
              head1 = acc1;
              meta |= kHeadEnded;
            } else {
              emit_final(cur, cur_key, acc, acc1);
            }
            first_run = false;
            acc = zero;
            acc1 = 0.f;
            cur = sg[u];
            cur_key = ky[u];
          }
          acc.x += r[u].x; acc.y += r[u].y; acc.z += r[u].z; acc.w += r[u].w;
          if (LIN) acc1 += r1[u];
        }
      }
    }
    // last run of the window
    if (first_run && !starts) {
      head = acc;
      head1 = acc1;
      if (ends) meta |= kHeadEnded;
    } else if (ends) {
      emit_final(cur, cur_key, acc, acc1);
    } else {
      tail = acc;
      tail1 = acc1;
      tail_seg = cur;
      tail_key = cur_key;
      meta |= kHasTail;
    }
    if (hi == a.n && lane == 0) {
      const int nseg = a.segidx[a.n - 1];
      *a.n_unique = nseg - ((a.keys[a.n - 1] == a.sentinel) ? 1 : 0);
    }
  }
  s_head[g * LPR + lane] = head;
  s_tail[g * LPR + lane] = tail;
  if (lane == 0) {
    s_meta[g] = meta;
    s_tailseg[g] = tail_seg;
    s_tailkey[g] = tail_key;
    if (LIN) {
      s_head1[g] = head1;
      s_tail1[g] = tail1;
    }
  }
  __syncthreads();

  // ---- stitch inside the CTA ---------------------------------------------------------
  // (a) a run that started in window g (tail) walks right through the heads of g+1..
  if (meta & kHasTail) {
    float4 acc = tail;
    float acc1 = tail1;
    int w = g + 1;
    bool ended = false;
    for (; w < G; ++w) {
      const int m = s_meta[w];
      if (!(m & kHasHead)) break;   // cannot happen while the run continues; defensive
      const float4 h = s_head[w * LPR + lane];
      acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
      if (LIN) acc1 += s_head1[w];
      if (m & kHeadEnded) { ended = true; break; }
    }
    if (ended) {
      emit_final(tail_seg, tail_key, acc, acc1);
    } else {   // runs off the CTA: it is the CTA's tail partial
      if (lane_on)
        *reinterpret_cast<float4*>(a.cta_tail + (long long)blockIdx.x * a.dim + lane * 4) = acc;
      if (LIN && lane == 0) a.cta_tail1[blockIdx.x] = acc1;
    }
  }
  // (b) the run entering the CTA from the left: group 0 walks it
  if (g == 0) {
    int cmeta = 0;
    if (s_meta[0] & kHasHead) {
      cmeta |= kHasHead;
      float4 acc = s_head[lane];
      float acc1 = LIN ? s_head1[0] : 0.f;
      bool ended = (s_meta[0] & kHeadEnded) != 0;
      for (int w = 1; w < G && !ended; ++w) {
        const int m = s_meta[w];
        if (!(m & kHasHead)) break;   // window w is empty (past n): the run ended with n
        const float4 h = s_head[w * LPR + lane];
        acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
        if (LIN) acc1 += s_head1[w];
        if (m & kHeadEnded) ended = true;
      }
      if (ended) cmeta |= kHeadEnded;
      if (lane_on)
        *reinterpret_cast<float4*>(a.cta_head + (long long)blockIdx.x * a.dim + lane * 4) = acc;
      if (LIN && lane == 0) a.cta_head1[blockIdx.x] = acc1;
    }
    // does some run leave the CTA on the right?  It is the tail of the last non-empty
    // window, or a head run that never ended.
    bool leaves = false;
    for (int w = G - 1; w >= 0; --w) {
      const int m = s_meta[w];
      const long long wlo = cta_lo + (long long)w * kWin;
      if (wlo >= a.n) continue;
      if (m & kHasTail) leaves = true;
      else if ((m & kHasHead) && !(m & kHeadEnded)) leaves = true;
      break;
    }
    if (leaves) {
      // a tail partial exists only if the leaving run STARTED in this CTA
      bool started_here = !((cmeta & kHasHead) && !(cmeta & kHeadEnded));
      if (started_here) cmeta |= kHasTail;
    }
    if (lane == 0) a.cta_meta[blockIdx.x] = cmeta;
  }
}

// Runs that cross CTA boundaries: one lane group per CTA that owns a tail partial.
template <int LPR, bool LIN = false>
__global__ void __launch_bounds__(kRedThreads) embed_fixup_kernel(const BwdArgs a, int n_cta) {
  constexpr int G = kRedThreads / LPR;
  const int c = blockIdx.x * G + threadIdx.x / LPR;
  const int lane = threadIdx.x % LPR;
  if (c >= n_cta) return;
  if (!(a.cta_meta[c] & kHasTail)) return;
  const bool lane_on = lane < a.vec_per_row;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float acc1 = 0.f;
  if (lane_on) acc = *reinterpret_cast<const float4*>(a.cta_tail + (long long)c * a.dim + lane * 4);
  if (LIN) acc1 = a.cta_tail1[c];
  for (int w = c + 1; w < n_cta; ++w) {
    const int m = a.cta_meta[w];
    if (!(m & kHasHead)) break;
    if (lane_on) {
      const float4 h = *reinterpret_cast<const float4*>(a.cta_head + (long long)w * a.dim + lane * 4);
      acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
    }
    if (LIN) acc1 += a.cta_head1[w];
    if (m & kHeadEnded) break;
  }
  // the run's identity: last sorted lookup of CTA c
  const long long last = min(a.n, (long long)(c + 1) * G * kWin) - 1;
  const int seg = a.segidx[last];
  const uint32_t key = a.keys[last];
  if (key == a.sentinel) return;
  if (lane_on) *reinterpret_cast<float4*>(a.grads + (long long)(seg - 1) * a.dim + lane * 4) = acc;
  if (lane == 0) {
    a.unique_rows[seg - 1] = (int)key;
    if (LIN) a.grads1[seg - 1] = acc1;
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
struct IdsView {
  int64_t B, F, L;
  bool i64;
};

static int parse_common(const DLTensor* ids, const int64_t* field_row_offset, int32_t n_fields,
                        int device, IdsView* v, FieldTable* ft) {
  KON_TRY(check_cuda_tensor(ids, "ids", device));
  KON_REQUIRE(field_row_offset != nullptr, KON_EINVAL, "field_row_offset is NULL");
  KON_REQUIRE(n_fields >= 1 && n_fields <= kMaxFields, KON_EUNSUPPORTED,
              "n_fields=%d outside [1,%d]", n_fields, kMaxFields);
  KON_REQUIRE(ids->ndim == 2 || ids->ndim == 3, KON_EINVAL, "ids must be [B,F] or [B,F,L]");
  KON_REQUIRE(is_i32(ids) || is_i64(ids), KON_EINVAL, "ids must be int32 or int64");
  KON_REQUIRE(is_compact(ids), KON_EINVAL, "ids must be compact row-major");
  v->B = ids->shape[0];
  v->F = ids->shape[1];
  v->L = ids->ndim == 3 ? ids->shape[2] : 1;
  v->i64 = is_i64(ids);
  KON_REQUIRE(v->F == n_fields, KON_EINVAL, "ids has %lld fields, n_fields=%d", (long long)v->F,
              n_fields);
  KON_REQUIRE(v->L >= 1 && v->L <= kMaxBagLen, KON_EUNSUPPORTED, "bag length %lld outside [1,%d]",
              (long long)v->L, kMaxBagLen);
  for (int f = 0; f <= n_fields; ++f) {
    ft->off[f] = field_row_offset[f];
    KON_REQUIRE(f == 0 || ft->off[f] >= ft->off[f - 1], KON_EINVAL,
                "field_row_offset must be non-decreasing");
  }
  KON_REQUIRE(ft->off[0] >= 0, KON_EINVAL, "field_row_offset[0] < 0");
  return KON_OK;
}

static int pow2_ge(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

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

// ---- backward workspace layout ----------------------------------------------------------
namespace {
struct BwdLayout {
  size_t keys_in, vals_in, keys_out, vals_out, segidx, cta_head, cta_tail, cta_meta, cub, cta_lin, total;
  size_t cub_bytes;
  int n_cta;
  int lpr, vpr, dim_pad;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int bwd_layout(int64_t n, int32_t dim, BwdLayout* l) {
  // the segmented reduction works on float4 lanes; dims that are not a multiple of 4 are
  // handled by the caller padding d_out (not needed by any reference configuration except
  // the dim-1 linear tables, which take the dim==1 scalar route below).
  l->vpr = (dim + 3) / 4;
  l->lpr = pow2_ge(l->vpr);
  if (l->lpr > 32) return -1;
  const int G = kRedThreads / l->lpr;
  l->n_cta = (int)((n + (int64_t)G * kWin - 1) / ((int64_t)G * kWin));
  if (l->n_cta < 1) l->n_cta = 1;
  size_t sort_bytes = 0, scan_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n, 0, 32);
  auto it = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), RunHead{nullptr});
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, it, (int*)nullptr, (int)n);
  l->cub_bytes = std::max(sort_bytes, scan_bytes);
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes, 256);
    return r;
  };
  l->keys_in = take((size_t)n * 4);
  l->vals_in = take((size_t)n * 4);
  l->keys_out = take((size_t)n * 4);
  l->vals_out = take((size_t)n * 4);
  l->segidx = take((size_t)n * 4);
  l->cta_head = take((size_t)l->n_cta * l->lpr * 16);
  l->cta_tail = take((size_t)l->n_cta * l->lpr * 16);
  l->cta_meta = take((size_t)l->n_cta * 4);
  l->cub = take(l->cub_bytes);
  l->cta_lin = take((size_t)l->n_cta * 8);   // kon_embed_bwd_pair: per-CTA head / tail partials of the 1-float gradient
  l->total = o;
  return 0;
}
}  // namespace

extern "C" size_t kon_embed_bwd_workspace_bytes(int64_t n_lookups, int32_t dim) {
  BwdLayout l;
  if (n_lookups <= 0) return 256;
  if (n_lookups > 0x7fffffffLL || bwd_layout(n_lookups, dim == 1 ? 4 : dim, &l) != 0) return 0;
  if (dim == 1) return l.total + 2 * align_up((size_t)n_lookups * 16, 256);
  return l.total;
}

// dim == 1 (linear tables) reuses the float4 machinery by treating each gradient as a
// one-lane row: the reduce kernel needs 16-B rows, so dim==1 is routed through a padded
// copy.  This small kernel spreads [B,F] -> [B,F,4] (x, 0, 0, 0) and the inverse.
__global__ void pad1_kernel(const float* __restrict__ src, long long sb, long long sf, int F,
                            long long n_bags, float4* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_bags;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / F;
    const int f = (int)(i - b * F);
    dst[i] = make_float4(src[b * sb + f * sf], 0.f, 0.f, 0.f);
  }
}

// grads [N,4] (column 0 valid) -> [N,1]; only the n_unique leading rows carry data
__global__ void __launch_bounds__(256)
unpad1_kernel(const float4* __restrict__ src, const int* __restrict__ n_unique, long long n,
              float* __restrict__ dst) {
  const long long nu = min((long long)*n_unique, n);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nu;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i].x;
}

namespace {
// where the upstream gradient rows live: one local [B,F,dim] view, or per-rank slabs over NVLink
struct GradSrc {
  const float* p = nullptr;
  long long sb = 0, sf = 0;
  int64_t dim = 0;
  int device = 0;
  int n_peers = 0;
  long long peer_rows = 0;
  const float* peer[kMaxPeers] = {};
  // kon_embed_bwd_pair: the fused one-float gradient and its output
  const float* lin = nullptr;
  long long lin_sb = 0, lin_sf = 0;
  float* lin_grads = nullptr;
};
}  // namespace

static int embed_bwd_core(const GradSrc& src, const DLTensor* ids, const int64_t* field_row_offset,
                          int32_t n_fields, DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                          DLTensor* workspace, int reuse_sort, void* stream);

static int embed_bwd_impl(const DLTensor* d_out, const DLTensor* ids, const int64_t* field_row_offset,
                          int32_t n_fields, DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                          DLTensor* workspace, int reuse_sort, void* stream) {
  KON_TRY(check_cuda_tensor(d_out, "d_out"));
  KON_TRY(check_cuda_tensor(ids, "ids", d_out->device.device_id));
  KON_REQUIRE(is_f32(d_out) && d_out->ndim == 3 && (ids->ndim == 2 || ids->ndim == 3) &&
                  d_out->shape[0] == ids->shape[0] && d_out->shape[1] == ids->shape[1],
              KON_EINVAL, "d_out must be float32 [B,F,dim]");
  GradSrc src;
  src.p = data_ptr<float>(d_out);
  src.sb = stride_of(d_out, 0);
  src.sf = stride_of(d_out, 1);
  src.dim = d_out->shape[2];
  src.device = d_out->device.device_id;
  KON_REQUIRE(src.dim == 1 || stride_of(d_out, 2) == 1, KON_EINVAL, "d_out last dim must be compact");
  return embed_bwd_core(src, ids, field_row_offset, n_fields, unique_rows, grads, n_unique, workspace,
                        reuse_sort, stream);
}

extern "C" int kon_embed_bwd(const DLTensor* d_out, const DLTensor* ids,
                             const int64_t* field_row_offset, int32_t n_fields,
                             DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                             DLTensor* workspace, void* stream) {
  return embed_bwd_impl(d_out, ids, field_row_offset, n_fields, unique_rows, grads, n_unique, workspace, 0,
                        stream);
}

extern "C" int kon_embed_bwd_reuse(const DLTensor* d_out, const DLTensor* ids,
                                   const int64_t* field_row_offset, int32_t n_fields,
                                   DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                                   DLTensor* workspace, void* stream) {
  return embed_bwd_impl(d_out, ids, field_row_offset, n_fields, unique_rows, grads, n_unique, workspace, 1,
                        stream);
}

// Two gradients over one routing: the embedding tables' [B,F,dim] gradient and the first-order ("linear", dim-1)
// tables' [B,F,1] gradient (any strides; a sum-pooled first-order term has stride_f = 0) when both tables were
// looked up with the same ids and per-field row counts -- the reference's FeatureInput(useLinear=True) (DP:65-76).
// One pass over the sorted lookups yields both; unique_rows / n_unique are shared.  Replaces a second
// pad + reduce + fixup + unpad chain (63 us at 1.7 M lookups).
extern "C" int kon_embed_bwd_pair(const DLTensor* d_out, const DLTensor* d_lin, const DLTensor* ids,
                                  const int64_t* field_row_offset, int32_t n_fields,
                                  DLTensor* unique_rows, DLTensor* grads, DLTensor* grads_lin,
                                  DLTensor* n_unique, DLTensor* workspace, int32_t reuse_sort, void* stream) {
  KON_TRY(check_cuda_tensor(d_out, "d_out"));
  const int dev = d_out->device.device_id;
  KON_TRY(check_cuda_tensor(d_lin, "d_lin", dev));
  KON_TRY(check_cuda_tensor(ids, "ids", dev));
  KON_TRY(check_cuda_tensor(grads_lin, "grads_lin", dev));
  KON_REQUIRE(is_f32(d_out) && d_out->ndim == 3 && ids->ndim == 2 && d_out->shape[0] == ids->shape[0] &&
                  d_out->shape[1] == ids->shape[1] && d_out->shape[2] >= 4 && d_out->shape[2] % 4 == 0,
              KON_EINVAL, "d_out must be float32 [B,F,dim], dim a multiple of 4, ids [B,F]");
  KON_REQUIRE(is_f32(d_lin) && d_lin->ndim == 3 && d_lin->shape[0] == ids->shape[0] &&
                  d_lin->shape[1] == ids->shape[1] && d_lin->shape[2] == 1,
              KON_EINVAL, "d_lin must be float32 [B,F,1]");
  KON_REQUIRE(is_f32(grads_lin) && is_compact(grads_lin) && numel(grads_lin) >= numel(ids), KON_EINVAL,
              "grads_lin must be compact float32 [>=N,1]");
  KON_REQUIRE(stride_of(d_out, 2) == 1, KON_EINVAL, "d_out last dim must be compact");
  GradSrc src;
  src.p = data_ptr<float>(d_out);
  src.sb = stride_of(d_out, 0);
  src.sf = stride_of(d_out, 1);
  src.dim = d_out->shape[2];
  src.device = dev;
  src.lin = data_ptr<float>(d_lin);
  src.lin_sb = stride_of(d_lin, 0);
  src.lin_sf = stride_of(d_lin, 1);
  src.lin_grads = data_ptr<float>(grads_lin);
  return embed_bwd_core(src, ids, field_row_offset, n_fields, unique_rows, grads, n_unique, workspace,
                        reuse_sort ? 1 : 0, stream);
}

// Routing only: builds the (arena row, position) keys, sorts them and scans the run heads into the front
// of `workspace`, where kon_embed_bwd_reuse / kon_embed_bwd_peer(reuse_sort = 1) pick them up.  The routing
// depends on the ids alone, so a trainer runs it on a side stream at the START of the step, off the
// critical path of the backward.
extern "C" int kon_embed_sort(const DLTensor* ids, const int64_t* field_row_offset, int32_t n_fields,
                              DLTensor* workspace, void* stream) {
  KON_TRY(check_cuda_tensor(ids, "ids"));
  GradSrc src;
  src.dim = 4;
  src.device = ids->device.device_id;
  return embed_bwd_core(src, ids, field_row_offset, n_fields, nullptr, nullptr, nullptr, workspace, 2, stream);
}

// Sharded backward over peer memory: the owner of the tables reads the gradient row of sample b
// from the gradient buffer of rank b / rows_per_peer while it reduces the sorted segments.
extern "C" int kon_embed_bwd_peer(const void* const* peer_d_out, int32_t n_peers,
                                  int64_t rows_per_peer, int64_t stride_b, int64_t stride_f,
                                  int32_t dim, const DLTensor* ids, const int64_t* field_row_offset,
                                  int32_t n_fields, DLTensor* unique_rows, DLTensor* grads,
                                  DLTensor* n_unique, DLTensor* workspace, int32_t reuse_sort,
                                  void* stream) {
  KON_TRY(check_cuda_tensor(ids, "ids"));
  KON_REQUIRE(peer_d_out != nullptr && n_peers >= 1 && n_peers <= kMaxPeers, KON_EINVAL,
              "n_peers=%d outside [1,%d]", n_peers, kMaxPeers);
  KON_REQUIRE(ids->ndim == 2, KON_EUNSUPPORTED, "the peer exchange takes [B,F] ids");
  KON_REQUIRE(rows_per_peer >= 1 && ids->shape[0] <= rows_per_peer * n_peers, KON_EINVAL,
              "ids has %lld samples, peers hold %lld x %d", (long long)ids->shape[0],
              (long long)rows_per_peer, n_peers);
  KON_REQUIRE(dim >= 4 && dim % 4 == 0, KON_EUNSUPPORTED, "the peer exchange needs dim %% 4 == 0");
  GradSrc src;
  src.sb = stride_b;
  src.sf = stride_f;
  src.dim = dim;
  src.device = ids->device.device_id;
  src.n_peers = n_peers;
  src.peer_rows = rows_per_peer;
  for (int q = 0; q < n_peers; ++q) {
    KON_REQUIRE(peer_d_out[q] != nullptr && aligned16(peer_d_out[q]), KON_EINVAL,
                "peer_d_out[%d] is NULL or not 16-B aligned", q);
    src.peer[q] = static_cast<const float*>(peer_d_out[q]);
  }
  src.p = src.peer[0];
  return embed_bwd_core(src, ids, field_row_offset, n_fields, unique_rows, grads, n_unique, workspace,
                        reuse_sort ? 1 : 0, stream);
}

static int embed_bwd_core(const GradSrc& src, const DLTensor* ids, const int64_t* field_row_offset,
                          int32_t n_fields, DLTensor* unique_rows, DLTensor* grads, DLTensor* n_unique,
                          DLTensor* workspace, int reuse_sort, void* stream) {
  const int dev = src.device;
  IdsView v;
  FieldTable ft;
  KON_TRY(parse_common(ids, field_row_offset, n_fields, dev, &v, &ft));
  const bool sort_only = reuse_sort == 2;     // kon_embed_sort: routing only (keys, radix sort, run-head scan)
  if (!sort_only) {
    KON_TRY(check_cuda_tensor(unique_rows, "unique_rows", dev));
    KON_TRY(check_cuda_tensor(grads, "grads", dev));
    KON_TRY(check_cuda_tensor(n_unique, "n_unique", dev));
  }
  KON_TRY(check_cuda_tensor(workspace, "workspace", dev));
  const int64_t dim = src.dim;
  KON_REQUIRE(dim % 4 == 0 || dim == 1, KON_EUNSUPPORTED,
              "embedding dim must be 1 or a multiple of 4 (got %lld)", (long long)dim);
  const int64_t n = v.B * v.F * v.L;
  KON_REQUIRE(n <= 0x7fffffffLL, KON_EUNSUPPORTED, "more than 2^31-1 lookups per call");
  KON_REQUIRE(v.B < (1LL << 24), KON_EUNSUPPORTED, "batch of %lld samples: the routing packs (sample, field) in 32 bits, B < 2^24",
              (long long)v.B);
  const int64_t total_rows = ft.off[n_fields];
  KON_REQUIRE(total_rows < 0x7fffffffLL, KON_EUNSUPPORTED,
              "arena with >= 2^31-1 rows: unique_rows is int32 (shard the tables over ranks / arenas)");
  if (!sort_only) {
    KON_REQUIRE(is_i32(unique_rows) && numel(unique_rows) >= n && is_compact(unique_rows),
                KON_EINVAL, "unique_rows must be compact int32 [>=N]");
    KON_REQUIRE(is_f32(grads) && grads->ndim == 2 && grads->shape[0] >= n && grads->shape[1] == dim &&
                    is_compact(grads),
                KON_EINVAL, "grads must be compact float32 [>=N,dim]");
    KON_REQUIRE(is_i32(n_unique) && numel(n_unique) >= 1, KON_EINVAL, "n_unique must be int32[1]");
  }
  DeviceGuard guard(dev);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    if (!sort_only) KON_CUDA(cudaMemsetAsync(data_ptr<int>(n_unique), 0, 4, st));
    return KON_OK;
  }
  const int rdim = dim == 1 ? 4 : (int)dim;   // row width seen by the reduce kernels
  BwdLayout l;
  KON_REQUIRE(bwd_layout(n, rdim, &l) == 0, KON_EUNSUPPORTED, "embedding dim %lld too large",
              (long long)dim);
  size_t need = l.total;
  size_t pad_off = 0;
  if (dim == 1) {   // padded d_out copy + padded grads live behind the regular layout
    pad_off = need;
    need += align_up((size_t)v.B * v.F * 16, 256) + align_up((size_t)n * 16, 256);
  }
  KON_REQUIRE(is_u8(workspace) && (size_t)numel(workspace) >= need, KON_EWORKSPACE,
              "workspace has %lld bytes, need %zu", (long long)numel(workspace), need);
  char* ws = data_ptr<char>(workspace);
  KON_REQUIRE(((uintptr_t)ws & 255u) == 0, KON_EINVAL, "workspace must be 256-B aligned");
  uint32_t* keys_in = (uint32_t*)(ws + l.keys_in);
  uint32_t* vals_in = (uint32_t*)(ws + l.vals_in);
  uint32_t* keys_out = (uint32_t*)(ws + l.keys_out);
  uint32_t* vals_out = (uint32_t*)(ws + l.vals_out);
  int* segidx = (int*)(ws + l.segidx);
  const int sms = sm_count_of(dev);
  const uint32_t sentinel = (uint32_t)total_rows;
  int end_bit = 1;
  while (end_bit < 32 && (total_rows >> end_bit) != 0) ++end_bit;

  const int kgrid = (int)std::min<long long>((n + 255) / 256, (long long)sms * 16);
  if (reuse_sort != 1) {
  if (v.i64)
    embed_keys_kernel<long long><<<kgrid, 256, 0, st>>>(data_ptr<long long>(ids), ft, (int)v.F,
                                                        (int)v.L, n, sentinel, keys_in, vals_in);
  else
    embed_keys_kernel<int><<<kgrid, 256, 0, st>>>(data_ptr<int>(ids), ft, (int)v.F, (int)v.L, n,
                                                  sentinel, keys_in, vals_in);
  KON_LAUNCH_CHECK("embed_keys_kernel");

  size_t cub_bytes = l.cub_bytes;
  {
    ProfileScope ps("embed_bwd_sort", st);
    KON_CUDA(cub::DeviceRadixSort::SortPairs(ws + l.cub, cub_bytes, keys_in, keys_out, vals_in,
                                             vals_out, (int)n, 0, end_bit, st));
  }
  cub_bytes = l.cub_bytes;
  auto it = thrust::make_transform_iterator(thrust::counting_iterator<int>(0), RunHead{keys_out});
  KON_CUDA(cub::DeviceScan::InclusiveSum(ws + l.cub, cub_bytes, it, segidx, (int)n, st));
  }   // reuse_sort != 1
  if (sort_only) return KON_OK;

  BwdArgs a;
  a.d_out = src.p;
  a.sb = src.sb;
  a.sf = src.sf;
  a.n_peers = src.n_peers;
  a.peer_rows = src.peer_rows;
  for (int q = 0; q < kMaxPeers; ++q) a.peer[q] = src.peer[q];
  a.F = (int)v.F;
  a.L = (int)v.L;
  a.dim = rdim;
  a.vec_per_row = l.vpr;
  a.n = n;
  a.keys = keys_out;
  a.vals = vals_out;
  a.segidx = segidx;
  a.sentinel = sentinel;
  a.unique_rows = data_ptr<int>(unique_rows);
  a.grads = data_ptr<float>(grads);
  a.n_unique = data_ptr<int>(n_unique);
  a.cta_head = (float*)(ws + l.cta_head);
  a.cta_tail = (float*)(ws + l.cta_tail);
  a.cta_meta = (int*)(ws + l.cta_meta);
  a.d1 = src.lin;
  a.sb1 = src.lin_sb;
  a.sf1 = src.lin_sf;
  a.grads1 = src.lin_grads;
  a.cta_head1 = (float*)(ws + l.cta_lin);
  a.cta_tail1 = a.cta_head1 + l.n_cta;
  float* padded_grads = nullptr;
  if (dim == 1) {
    float4* padded = (float4*)(ws + pad_off);
    padded_grads = (float*)(ws + pad_off + align_up((size_t)v.B * v.F * 16, 256));
    const long long nb = v.B * v.F;
    pad1_kernel<<<(int)std::min<long long>((nb + 255) / 256, (long long)sms * 16), 256, 0, st>>>(
        a.d_out, a.sb, a.sf, (int)v.F, nb, padded);
    KON_LAUNCH_CHECK("pad1_kernel");
    a.d_out = (const float*)padded;
    a.sb = v.F * 4;
    a.sf = 4;
    a.grads = padded_grads;
  } else {
    KON_REQUIRE(aligned16(a.d_out) && a.sb % 4 == 0 && a.sf % 4 == 0, KON_EINVAL,
                "d_out rows must be 16-B aligned");
  }
  const size_t smem = 2 * (size_t)kRedThreads * 16;
  ProfileScope ps_red("embed_reduce_kernel", st);
#define KON_RED_CASE(N)                                                                  \
  case N:                                                                                \
    if (a.d1 != nullptr) {                                                               \
      embed_reduce_kernel<N, KON_EMB_RED_NB_LIN, true><<<l.n_cta, kRedThreads, smem, st>>>(a); \
      KON_LAUNCH_CHECK("embed_reduce_kernel");                                           \
      embed_fixup_kernel<N, true>                                                        \
          <<<(l.n_cta + kRedThreads / N - 1) / (kRedThreads / N), kRedThreads, 0, st>>>(a, l.n_cta); \
      KON_LAUNCH_CHECK("embed_fixup_kernel");                                            \
      break;                                                                             \
    }                                                                                    \
    if (a.n_peers > 1) embed_reduce_kernel<N, kWin><<<l.n_cta, kRedThreads, smem, st>>>(a);  \
    else embed_reduce_kernel<N, KON_EMB_RED_NB><<<l.n_cta, kRedThreads, smem, st>>>(a);  \
    KON_LAUNCH_CHECK("embed_reduce_kernel");                                             \
    embed_fixup_kernel<N>                                                                \
        <<<(l.n_cta + kRedThreads / N - 1) / (kRedThreads / N), kRedThreads, 0, st>>>(a, l.n_cta); \
    KON_LAUNCH_CHECK("embed_fixup_kernel");                                              \
    break;
  switch (l.lpr) {
    KON_RED_CASE(1)
    KON_RED_CASE(2)
    KON_RED_CASE(4)
    KON_RED_CASE(8)
    KON_RED_CASE(16)
    KON_RED_CASE(32)
    default:
      return fail(KON_EUNSUPPORTED, "embedding dim too large");
  }
#undef KON_RED_CASE
  if (dim == 1) {   // compact [N,4] -> [N,1]  (a 2-D memcpy with 4-byte rows takes 160 us for 1.7 M rows)
    unpad1_kernel<<<(int)std::min<long long>((n + 255) / 256, (long long)sms * 16), 256, 0, st>>>(
        reinterpret_cast<const float4*>(padded_grads), a.n_unique, n, data_ptr<float>(grads));
    KON_LAUNCH_CHECK("unpad1_kernel");
  }
  return KON_OK;
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
