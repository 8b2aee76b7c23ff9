// a8: xDeepFM CIN on the 5th-gen tensor cores (KON_CIN_BF16): bf16 operands, fp32 accumulate.
//
// GEMM view of one CIN layer (IL:310-322): rows r = (b,d), M = B*D;  reduction index
// c = h*m + i (IL:317-318), K = H_prev*m;  z[r,o] = sum_c A[r,c] W[c,o] + bias[o] with the
// rank-1 operand A[r,c] = pre[r,h] * x0[r,i], which the reference materialises ([D,B,m,H],
// 21.8 GB at B = 65536) and this kernel never stores anywhere but tensor memory.
//
// Kernel skeleton shared by the forward GEMM and the weight-gradient GEMM ("stream" kernel):
//   * one persistent CTA per SM, 10 warps:
//       warps 0-3 / 4-7 : two 128-lane sub-tiles.  Each thread owns one TMEM lane (a row of the
//                         A operand); it computes its bf16x2 products with packed HMUL2 and
//                         writes them with tcgen05.st into a small TMEM ring (A never touches
//                         shared memory or HBM).  After the K loop the same warps run the
//                         epilogue (tcgen05.ld of their lanes).
//       warp 8          : one elected thread issues tcgen05.mma (A from TMEM, B from smem),
//                         M = 128 per sub-tile, N = N_pad <= 208, K = 16 per instruction.
//                         Both sub-tiles consume every B stage -> half the L2->SM operand traffic.
//       warp 9          : one elected thread streams the B operand with 1-D bulk async copies
//                         (UBLKCP) of pre-packed, descriptor-ready blocks through an 8-stage
//                         mbarrier ring.
//   * TMEM map (512 columns): D0 [0,N) | A ring 0 [208,256) | D1 [256,256+N) | A ring 1 [464,512)
//   * no swizzle: operands are stored as 8x8 core matrices (128 contiguous bytes); the layouts
//     and descriptor fields were pinned on hardware with tools/tc_probe.cu.
#include "cin_tc_common.cuh"

namespace kon {

using namespace tcs;

namespace {

// ---------------------------------------------------------------------------------------------
// Weight packing: W[C,N] fp32 (Keras Conv1D kernel, row c = h*m+i) -> bf16 core-matrix blocks.
// ---------------------------------------------------------------------------------------------
// Forward B operand, K-major (B[n][k], k = c): per k-step s (16 c's) one block of 2*N8*128 bytes:
//   offset(s, c%16, n) = s*blk + (c%16/8)*N8*128 + (n/8)*128 + (n%8)*16 + (c%8)*2
__global__ void __launch_bounds__(256)
cin_pack_w_fwd_kernel(const float* __restrict__ W, int C, int N, int N8, int nk,
                      __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)nk * 16 * N8 * 8;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // consecutive idx -> consecutive output elements
    const int c8 = (int)(idx & 7);
    const int n7 = (int)((idx >> 3) & 7);
    long long t = idx >> 6;
    const int n8 = (int)(t % N8);
    t /= N8;
    const int kg = (int)(t & 1);
    const long long s = t >> 1;
    const long long c = s * 16 + kg * 8 + c8;
    const int n = n8 * 8 + n7;
    float v = 0.f;
    if (c < C && n < N) v = W[c * N + n];
    out[idx] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------
// The stream kernel
// ---------------------------------------------------------------------------------------------
struct FwdArgs {
  const float* x0;              // [B, m, D] fp32, batch stride x0_sb (elements), rows compact
  long long x0_sb;
  const unsigned short* pre;    // [B, Hp, D] bf16 bits, or nullptr on layer 1 (pre == x0)
  const unsigned char* wpack;   // nk blocks of kblk bytes
  const float* bias;            // [N]
  unsigned short* zt;           // out: [B, N, D] bf16 (the reference's `pre_` layout, IL:320)
  float* pooled;                // [B, pooled_stride], this layer's columns start at pooled_col0
  int pooled_stride, pooled_col0;
  long long rows;               // B*D
  int D, Hp, N, N8, nk;
  uint32_t kblk;                // bytes per k-step block of wpack
  long long n_pairs;            // ceil(rows / (128*kSub))
};

// Row warps come in kFwdSets sets (cf. cin_dw2_tc_kernel): with kNS == 2 A slots per sub-tile, set s
// fills slot s for the groups g = s, s+2, ..., so the wait -> HMUL2 -> tcgen05.st -> wait::st ->
// arrive chain of one group overlaps with the other set's chain.
#ifndef KON_FWD_SETS
#define KON_FWD_SETS 1
#endif
constexpr int kFwdSets = KON_FWD_SETS;
constexpr int kFwdProd = kProdWarps * kFwdSets;
constexpr int kFwdThreads = 32 * (kFwdProd + 2);
static_assert(kFwdSets == 1 || (kFwdSets == 2 && kNS == 2), "two sets <-> two A slots");

template <int MF>
__global__ void __launch_bounds__(kFwdThreads, 1) cin_fwd_tc_kernel(const FwdArgs a) {
  constexpr int LCM = lcm_(16, MF);
  constexpr int PK = LCM / 16;   // k-steps per period
  constexpr int PH = LCM / MF;   // feature maps (h) per period
  static_assert(MF % 2 == 0, "field count must be even (bf16x2 pairs never straddle an h)");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Barriers bars;
  __shared__ float s_bias[kMaxN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kMaxN; i += kFwdThreads) s_bias[i] = i < a.N ? a.bias[i] : 0.f;

  stream_init(bars, tid, warp, 1, kFwdProd);
  const uint32_t tmem = bars.tmem_base;

  if (warp < kFwdProd) {
    // ================= producers + epilogue =================================================
    const int set = warp / kProdWarps, w8 = warp % kProdWarps;
    const int sub = w8 >> 2;
    const uint32_t lane_base = (uint32_t)((w8 & 3) * 32) << 16;
    const uint32_t colD = sub ? kColD1 : kColD0;
    const uint32_t colA = sub ? kColA1 : kColA0;
    const int n_periods = (a.Hp + PH - 1) / PH;
    SlotWriter sw;                 // kFwdSets == 1
    uint32_t my_phase = 0;         // kFwdSets == 2: phase of this set's slot
    int gcnt = 0, gpar = 0;        // k-steps into the current group / parity of the current group
    uint32_t tile_it = 0;
    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x, ++tile_it) {
      const long long r = pair * (128 * kSub) + sub * 128 + (w8 & 3) * 32 + lane;
      const bool valid = r < a.rows;
      const long long b = valid ? r / a.D : 0;
      const int d = valid ? (int)(r - b * a.D) : 0;
      const float* xrow = a.x0 + b * a.x0_sb + d;                      // x0[b,i,d] = xrow[i*D]
      uint32_t x2[MF / 2];
#pragma unroll
      for (int i = 0; i < MF / 2; ++i) {
        const float lo = valid ? __ldg(xrow + (long long)(2 * i) * a.D) : 0.f;
        const float hi = valid ? __ldg(xrow + (long long)(2 * i + 1) * a.D) : 0.f;
        x2[i] = tc::pack_bf16(lo, hi);
      }
      // pre[r,h] for the PH feature maps of a period, fetched one period ahead.  The raw loads
      // land in distinct registers and are only touched (broadcast to bf16x2) a period later, so
      // the L2 latency hides behind ~13 k-steps of work; out-of-range h reads a clamped address
      // and is zeroed by the mask instead of branching around the load.
      const unsigned short* prow = a.pre ? a.pre + b * (long long)a.Hp * a.D + d : nullptr;
      auto load_raw = [&](int h) -> uint32_t {
        const int hc = min(h, a.Hp - 1);
        if (prow) return (uint32_t)__ldg(prow + (long long)hc * a.D);
        return __float_as_uint(__ldg(xrow + (long long)hc * a.D));
      };
      auto to_bcast = [&](uint32_t raw, int h) -> uint32_t {
        if (!valid || h >= a.Hp) return 0u;
        return prow ? bf16_bcast_raw((unsigned short)raw) : bf16_bcast(__uint_as_float(raw));
      };
      uint32_t cur[PH], nraw[PH];
#pragma unroll
      for (int j = 0; j < PH; ++j) nraw[j] = load_raw(j);

      int ks_global = 0;     // k-step index within this tile
      int per = 0;
      // one k-step: 8 packed products of this lane's row (j = k-step within the period: static)
#define KON_FWD_KSTEP(J, W)                                                         \
  _Pragma("unroll") for (int q = 0; q < 8; ++q) {                                   \
    const int cl = 16 * (J) + 2 * q;                                                \
    (W)[q] = hmul2_bf16(cur[cl / MF], x2[(cl % MF) / 2]);                           \
  }
      auto refresh = [&]() {
#pragma unroll
        for (int j = 0; j < PH; ++j) cur[j] = to_bcast(nraw[j], per * PH + j);
#pragma unroll
        for (int j = 0; j < PH; ++j) nraw[j] = load_raw((per + 1) * PH + j);
        ++per;
      };
      if (kFwdSets == 1) {
        // ---- fast path: kG periods = kG*PK k-steps = PK groups; every slot wait / hand-over sits at
        // a compile-time position of the unrolled body (the generic loop below spends 3x the
        // instructions of the 8 HMUL2 + STTM on per-k-step slot bookkeeping)
        while (sw.cnt == 0 && ks_global + kG * PK <= a.nk) {
#pragma unroll
          for (int pp = 0; pp < kG; ++pp) {
            refresh();
#pragma unroll
            for (int j = 0; j < PK; ++j) {
              const int jj = pp * PK + j;                 // static
              if (jj % kG == 0) {
                mbar_wait(&bars.a_empty[sub][sw.slot], sw.phase ^ 1);
                tc::fence_after();
              }
              uint32_t w[8];
              KON_FWD_KSTEP(j, w)
              tc::st8(tmem + lane_base + colA + sw.slot * (8 * kG) + 8 * (jj % kG), w);
              if (jj % kG == kG - 1) {
                tc::wait_st();
                tc::fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.a_full[sub][sw.slot]);
                if (++sw.slot == kNS) { sw.slot = 0; sw.phase ^= 1; }
              }
            }
          }
          ks_global += kG * PK;
        }
      }
      // ---- generic path (tail of the K loop, partial periods, two-set mode) ---------------------
      for (; per < n_periods;) {
        refresh();
#pragma unroll
        for (int j = 0; j < PK; ++j) {
          if (ks_global < a.nk) {
            ++ks_global;
            const bool last = ks_global == a.nk;
            if (kFwdSets == 1 || gpar == set) {
              uint32_t w[8];
              KON_FWD_KSTEP(j, w)
              if (kFwdSets == 1) {
                sw.put(bars, sub, tmem + lane_base + colA, w, last, lane);
              } else {
                if (gcnt == 0) {
                  mbar_wait(&bars.a_empty[sub][set], my_phase ^ 1);
                  tc::fence_after();
                }
                tc::st8(tmem + lane_base + colA + set * (8 * kG) + 8 * gcnt, w);
                if (gcnt == kG - 1 || last) {
                  tc::wait_st();
                  tc::fence_before();
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&bars.a_full[sub][set]);
                  my_phase ^= 1;
                }
              }
            }
            if (kFwdSets == 2) {
              if (++gcnt == kG || last) { gcnt = 0; gpar ^= 1; }
            }
          }
        }
      }
#undef KON_FWD_KSTEP
      // ---- epilogue: z = D + bias; pooled = sum_o z; z^T (bf16) -> zt ---------------------
      mbar_wait(&bars.d_full, tile_it & 1);
      tc::fence_after();
      float rsum = 0.f;
      unsigned short* zrow = a.zt + b * (long long)a.N * a.D + d;       // zt[b,o,d] = zrow[o*D]
      const int n_full = a.N & ~15;
      for (int o0 = (kFwdSets == 2 ? set * 16 : 0); o0 < n_full; o0 += 16 * kFwdSets) {
        uint32_t v[16];
        tc::ld16(tmem + lane_base + colD + o0, v);
        tc::wait_ld();
#pragma unroll
        for (int q = 0; q < 16; q += 2) {
          const float z0 = __uint_as_float(v[q]) + s_bias[o0 + q];
          const float z1 = __uint_as_float(v[q + 1]) + s_bias[o0 + q + 1];
          rsum += z0;
          rsum += z1;
          const uint32_t pk = tc::pack_bf16(z0, z1);
          if (valid) {
            zrow[(long long)(o0 + q) * a.D] = (unsigned short)(pk & 0xffffu);
            zrow[(long long)(o0 + q + 1) * a.D] = (unsigned short)(pk >> 16);
          }
        }
      }
      for (int o0 = n_full; o0 < a.N && (kFwdSets == 1 || set == 0); o0 += 8) {     // tail: N % 16 in {8} or ragged
        uint32_t v[8];
        tc::ld8(tmem + lane_base + colD + o0, v);
        tc::wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (o0 + q < a.N) {
            const float z = __uint_as_float(v[q]) + s_bias[o0 + q];
            rsum += z;
            if (valid) zrow[(long long)(o0 + q) * a.D] = (unsigned short)(tc::pack_bf16(z, 0.f) & 0xffffu);
          }
        }
      }
      if (valid) {
        float* pp = a.pooled + b * a.pooled_stride + a.pooled_col0 + d;
        if (kFwdSets == 1) *pp = rsum;
        else atomicAdd(pp, rsum);     // two addends onto a zeroed cell: order-independent, deterministic
      }
      tc::fence_before();   // our tcgen05.ld are complete (wait_ld) before the next tile's a_full arrive
      // two sets: the set that owns slot 0 must not let the next tile's first MMA (which overwrites
      // the accumulator) start before the OTHER set has finished reading its columns
      if (kFwdSets == 2) tc::named_bar_sync(1 + sub, 32 * kProdWarps / kSub * kFwdSets);
    }
  } else if (warp == kFwdProd) {
    // ================= MMA issuer ===========================================================
    {
      const long long n_items = a.n_pairs > blockIdx.x ? (a.n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      stream_mma_role(bars, smem, tmem, n_items, a.nk, a.kblk, tc::idesc_bf16(128, a.N8 * 8, 0, 0),
                      (uint32_t)a.N8 * 128, 128);
    }
  } else {
    // ================= B loader =============================================================
    if (lane == 0) {
      const long long n_items = a.n_pairs > blockIdx.x ? (a.n_pairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      stream_loader_role(bars, smem, n_items, a.nk, a.kblk, [&](long long) { return a.wpack; });
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == kFwdProd) tc::tmem_dealloc(tmem, 512);
}


// ---------------------------------------------------------------------------------------------
// Last layer, forward.  z_L feeds nothing but reduce_sum(z, -1) (IL:322) -- there is no next layer
// to consume it as `pre` -- so the pooled output needs no GEMM at all:
//     pool_L[r] = sum_o (sum_c A[r,c] W[c,o] + bias[o]) = sum_h pre[r,h] * (sum_i x0[r,i] wsum[h,i]) + sum_o bias[o]
// with wsum[h,i] = sum_o W[h*m+i, o]: a 200x26 mat-vec per row (the backward already uses the same
// identity, cin_last_da_kernel).  HBM-bound on the 419 MB of `pre` instead of a 2.18-TFLOP GEMM, in
// fp32 (x0 is not even rounded to bf16), and z_L^T -- 419 MB nobody reads -- is never written.
// ---------------------------------------------------------------------------------------------
constexpr int kPoolTRow = 28;   // wsum rows padded to 28 floats: 16-byte aligned rows, 2 zero columns

// T[h*28 + i] = sum_o W[(h*m+i)*N + o] (one warp per entry); warp 0 of block 0 also leaves sum_o bias[o]
__global__ void __launch_bounds__(256)
cin_fwd_wsum_kernel(const float* __restrict__ W, const float* __restrict__ bias, int C, int N, int m, int Hp8,
                    float* __restrict__ T, float* __restrict__ bsum) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long total = (long long)Hp8 * kPoolTRow;
  for (long long t = w0; t < total; t += nw) {
    const int h = (int)(t / kPoolTRow), i = (int)(t - (long long)h * kPoolTRow);
    const long long c = (long long)h * m + i;
    float acc = 0.f;
    if (i < m && c < C)
      for (int o = lane; o < N; o += 32) acc += W[c * N + o];
    acc = warp_sum(acc);
    if (lane == 0) T[t] = acc;
  }
  if (w0 == 0) {
    float acc = 0.f;
    for (int o = lane; o < N; o += 32) acc += bias[o];
    acc = warp_sum(acc);
    if (lane == 0) *bsum = acc;
  }
}

struct LastPoolArgs {
  const float* x0;              // [B, m, D] fp32, batch stride x0_sb
  long long x0_sb;
  const unsigned short* pre;    // [B, Hp, D] bf16 bits (z^T of the layer before)
  const float* T;               // [Hp8, 28]
  const float* bsum;            // [1]
  float* pooled;
  int pooled_stride, pooled_col0;
  long long rows;               // B*D
  int D, Hp, Hp8;
};

__device__ __forceinline__ void pool_ffma2(float2& acc, const float2 a, const float2 b) {
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(acc.x), "f"(acc.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(rc));
}

constexpr int kPoolThreads = 256;
constexpr int kPoolRows = 2;    // rows per thread: every shared-memory read of wsum serves two rows

template <int MF>
__global__ void __launch_bounds__(kPoolThreads, 2) cin_last_pool_kernel(const LastPoolArgs a) {
  static_assert(MF <= kPoolTRow && MF % 2 == 0, "field count");
  extern __shared__ __align__(16) float sT[];                  // [Hp8][28]
  for (int i = threadIdx.x; i < a.Hp8 * kPoolTRow; i += kPoolThreads) sT[i] = a.T[i];
  __syncthreads();
  const float bs = *a.bsum;
  constexpr int TILE = kPoolThreads * kPoolRows;
  for (long long base = (long long)blockIdx.x * TILE; base < a.rows; base += (long long)gridDim.x * TILE) {
    float2 x[kPoolRows][kPoolTRow / 2];
    const unsigned short* prow[kPoolRows];
    bool valid[kPoolRows];
    long long bb[kPoolRows];
    int dd[kPoolRows];
#pragma unroll
    for (int u = 0; u < kPoolRows; ++u) {
      const long long r = base + u * kPoolThreads + threadIdx.x;
      valid[u] = r < a.rows;
      const long long b = valid[u] ? r / a.D : 0;
      const int d = valid[u] ? (int)(r - b * a.D) : 0;
      bb[u] = b;
      dd[u] = d;
      const float* xrow = a.x0 + b * a.x0_sb + d;              // x0[b,i,d] = xrow[i*D]
      prow[u] = a.pre + b * (long long)a.Hp * a.D + d;         // pre[b,h,d] = prow[h*D]
#pragma unroll
      for (int i = 0; i < kPoolTRow / 2; ++i) {
        x[u][i].x = (2 * i < MF && valid[u]) ? __ldg(xrow + (long long)(2 * i) * a.D) : 0.f;
        x[u][i].y = (2 * i + 1 < MF && valid[u]) ? __ldg(xrow + (long long)(2 * i + 1) * a.D) : 0.f;
      }
    }
    float acc[kPoolRows];
    unsigned short praw[kPoolRows][8];
#pragma unroll
    for (int u = 0; u < kPoolRows; ++u) {
      acc[u] = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) praw[u][q] = __ldg(prow[u] + (long long)min(q, a.Hp - 1) * a.D);
    }
    for (int h0 = 0; h0 < a.Hp8; h0 += 8) {
      float pv[kPoolRows][8];
#pragma unroll
      for (int u = 0; u < kPoolRows; ++u) {
#pragma unroll
        for (int q = 0; q < 8; ++q) pv[u][q] = (h0 + q < a.Hp) ? bf16_to_f32(praw[u][q]) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q)      // next 8 feature maps, one iteration ahead (clamped address)
          praw[u][q] = __ldg(prow[u] + (long long)min(h0 + 8 + q, a.Hp - 1) * a.D);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4* trow = reinterpret_cast<const float4*>(sT + (h0 + q) * kPoolTRow);
        float2 s2[kPoolRows];
#pragma unroll
        for (int u = 0; u < kPoolRows; ++u) s2[u] = make_float2(0.f, 0.f);
#pragma unroll
        for (int z = 0; z < kPoolTRow / 4; ++z) {
          const float4 t = trow[z];                            // warp-uniform address: one broadcast read
#pragma unroll
          for (int u = 0; u < kPoolRows; ++u) {
            pool_ffma2(s2[u], make_float2(t.x, t.y), x[u][2 * z]);
            pool_ffma2(s2[u], make_float2(t.z, t.w), x[u][2 * z + 1]);
          }
        }
#pragma unroll
        for (int u = 0; u < kPoolRows; ++u) acc[u] = fmaf(pv[u][q], s2[u].x + s2[u].y, acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kPoolRows; ++u)
      if (valid[u]) a.pooled[bb[u] * a.pooled_stride + a.pooled_col0 + dd[u]] = acc[u] + bs;
  }
}


// Same computation on warp-level tensor-core MMAs: one warp owns a 16-row tile (the 16 coordinates d of one sample
// when D == 16) and computes t[h,d] = sum_i wsum[h,i] x0[d,i] as m16n8k16 bf16 MMAs with M = feature maps (13 tiles
// of 16), N = coordinates (2 tiles of 8), K = 26 fields padded to 32.  In this orientation a thread's accumulator
// pair (h, d..d+1) sits on ONE 32-bit word of `pre` ([B,H,D] bf16: d is the contiguous axis), so the multiply by
// pre[b,h,d] costs one 4-byte load per two products; the wsum fragments are pre-arranged in shared memory so that
// each A fragment is one conflict-free LDS.128.  ~390 warp instructions per sample (FFMA2 kernel: ~2250, the first
// MMA version with M = coordinates: 1350): the kernel streams the 419 MB of `pre` at the pace of its loads.
// (Legacy mma.sync on purpose: 13 GFLOP of side work.)
constexpr int kPoolMmaWarps = 8;

__device__ __forceinline__ void pool_mma_16816(float* c, const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

template <int MF>
__global__ void __launch_bounds__(32 * kPoolMmaWarps) cin_last_pool_mma_kernel(const LastPoolArgs a) {
  static_assert(MF <= 32, "field count");
  extern __shared__ __align__(16) uint4 sA[];                   // [n_mt][2 k-steps][32 lanes]: A fragments of wsum
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_mt = a.Hp8 / 16 + ((a.Hp8 % 16) ? 1 : 0);
  for (int e = tid; e < n_mt * 2 * 32; e += blockDim.x) {
    const int l = e & 31, ks = (e >> 5) & 1, mt = e >> 6;
    const int gg = l >> 2, tt = l & 3;
    auto w2 = [&](int h, int i) -> uint32_t {                   // (wsum[h][i], wsum[h][i+1]) as bf16x2, zero outside
      float lo = 0.f, hi = 0.f;
      if (h < a.Hp8) {
        if (i < MF) lo = a.T[h * kPoolTRow + i];
        if (i + 1 < MF) hi = a.T[h * kPoolTRow + i + 1];
      }
      return tc::pack_bf16(lo, hi);
    };
    const int h0 = 16 * mt + gg, i0 = 16 * ks + 2 * tt;
    sA[e] = make_uint4(w2(h0, i0), w2(h0 + 8, i0), w2(h0, i0 + 8), w2(h0 + 8, i0 + 8));
  }
  __syncthreads();
  const float bs = *a.bsum;
  const int g = lane >> 2, t = lane & 3;
  const int tiles_per_b = a.D / 16;
  const long long n_tiles = a.rows / 16;
  for (long long tile = (long long)blockIdx.x * kPoolMmaWarps + warp; tile < n_tiles;
       tile += (long long)gridDim.x * kPoolMmaWarps) {
    const long long b = tile / tiles_per_b;
    const int d0 = (int)(tile - b * tiles_per_b) * 16;
    const float* xb = a.x0 + b * a.x0_sb + d0;                  // x0[b,i,d0+d] = xb[i*D + d]
    const unsigned short* pb = a.pre + b * (long long)a.Hp * a.D + d0;   // pre[b,h,d0+d] = pb[h*D + d]
    // B fragments of x0^T (k = field i, n = coordinate d): b0 = (i = 16ks+2t, +1 ; d = 8nt+g), b1 = fields +8
    uint32_t bf[2][2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int i = 16 * ks + 2 * t + 8 * q;
          const int d = 8 * nt + g;
          const float lo = i < MF ? __ldg(xb + (long long)i * a.D + d) : 0.f;
          const float hi = i + 1 < MF ? __ldg(xb + (long long)(i + 1) * a.D + d) : 0.f;
          bf[nt][ks][q] = tc::pack_bf16(lo, hi);
        }
    float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};                 // [nt][d = 8nt + 2t, +1], partial over this lane's h rows
#pragma unroll 13
    for (int mt = 0; mt < n_mt; ++mt) {
      const int hA = min(16 * mt + g, a.Hp - 1), hB = min(16 * mt + g + 8, a.Hp - 1);   // rows >= Hp: wsum is 0 there
      uint32_t pw[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        pw[nt][0] = __ldg(reinterpret_cast<const uint32_t*>(pb + (long long)hA * a.D + 8 * nt + 2 * t));
        pw[nt][1] = __ldg(reinterpret_cast<const uint32_t*>(pb + (long long)hB * a.D + 8 * nt + 2 * t));
      }
      const uint4 a0 = sA[(mt * 2 + 0) * 32 + lane], a1 = sA[(mt * 2 + 1) * 32 + lane];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        pool_mma_16816(c, a0, bf[nt][0][0], bf[nt][0][1]);
        pool_mma_16816(c, a1, bf[nt][1][0], bf[nt][1][1]);
        acc[nt][0] = fmaf(__uint_as_float(pw[nt][0] << 16), c[0], acc[nt][0]);
        acc[nt][1] = fmaf(__uint_as_float(pw[nt][0] & 0xffff0000u), c[1], acc[nt][1]);
        acc[nt][0] = fmaf(__uint_as_float(pw[nt][1] << 16), c[2], acc[nt][0]);
        acc[nt][1] = fmaf(__uint_as_float(pw[nt][1] & 0xffff0000u), c[3], acc[nt][1]);
      }
    }
    // the 8 lanes with the same t hold disjoint feature maps of the same coordinates: fixed-order butterfly over g
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v = acc[nt][q];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[nt][q] = v;
      }
    if (g == 0) {
      float* pp = a.pooled + b * a.pooled_stride + a.pooled_col0 + d0;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        pp[8 * nt + 2 * t] = acc[nt][0] + bs;
        pp[8 * nt + 2 * t + 1] = acc[nt][1] + bs;
      }
    }
  }
}

}  // namespace

size_t cin_tc_saved_bytes(int64_t B, int m, int D, const int32_t* hs, int nl) {
  TcLayout L;
  if (tc_layout(B, m, D, hs, nl, 148, &L) != 0) return 256;
  return L.saved_total;
}

size_t cin_tc_workspace_bytes(int64_t B, int m, int D, const int32_t* hs, int nl, int sms) {
  TcLayout L;
  if (tc_layout(B, m, D, hs, nl, sms, &L) != 0) return 256;
  return L.work_total;
}

int cin_tc_fwd(const float* x0, long long x0_sb, const float* const* w, const float* const* bias, int nl,
               const int32_t* hs, int64_t B, int m, int D, float* pooled, void* saved,
               void* workspace, int sms, cudaStream_t st) {
  TcLayout L;
  KON_TRY(tc_check_layout(tc_layout(B, m, D, hs, nl, sms, &L), m, D));
  KON_REQUIRE(((uintptr_t)workspace & 255u) == 0 && ((uintptr_t)saved & 255u) == 0, KON_EINVAL,
              "saved / workspace must be 256-B aligned");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned char* sv = static_cast<unsigned char*>(saved);
  const long long rows = B * D;
  const size_t smem = stream_smem_bytes();
  static bool attr_done = false;
  if (!attr_done) {
    KON_CUDA(cudaFuncSetAttribute(cin_fwd_tc_kernel<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  for (int l = 0; l < nl; ++l) {
    const long long tot = (long long)L.nk[l] * 16 * L.N8[l] * 8;
    cin_pack_w_fwd_kernel<<<(int)std::min<long long>((tot + 255) / 256, (long long)sms * 8), 256, 0, st>>>(
        w[l], L.Hp[l] * m, L.N[l], L.N8[l], L.nk[l], reinterpret_cast<__nv_bfloat16*>(ws + L.wpack_off[l]));
    KON_LAUNCH_CHECK("cin_pack_w_fwd_kernel");
  }
  if (kFwdSets == 2) KON_CUDA(cudaMemsetAsync(pooled, 0, (size_t)B * nl * D * 4, st));
  // last layer (nl >= 2): pooled output straight from pre and the column sums of W (cin_last_pool_kernel)
  const bool pool_shortcut = tc_last_layer_pooled_only(nl);
  for (int l = 0; l < nl; ++l) {
    if (pool_shortcut && l == nl - 1) {
      const int Hp8 = (L.Hp[l] + 7) / 8 * 8;
      float* T = reinterpret_cast<float*>(ws + L.tail_off);
      float* bsum = T + (size_t)Hp8 * kPoolTRow;
      cin_fwd_wsum_kernel<<<(int)std::min<long long>(((long long)Hp8 * kPoolTRow * 32 + 255) / 256, (long long)sms * 8),
                            256, 0, st>>>(w[l], bias[l], L.Hp[l] * m, L.N[l], m, Hp8, T, bsum);
      KON_LAUNCH_CHECK("cin_fwd_wsum_kernel");
      LastPoolArgs z;
      z.x0 = x0;
      z.x0_sb = x0_sb;
      z.pre = reinterpret_cast<const unsigned short*>(sv + L.zt_off[l - 1]);
      z.T = T;
      z.bsum = bsum;
      z.pooled = pooled;
      z.pooled_stride = nl * D;
      z.pooled_col0 = l * D;
      z.rows = rows;
      z.D = D;
      z.Hp = L.Hp[l];
      z.Hp8 = Hp8;
      if (D % 16 == 0 && rows % 16 == 0) {
        const long long tiles = rows / 16;
        const int grid = (int)std::min<long long>((tiles + kPoolMmaWarps - 1) / kPoolMmaWarps, (long long)sms * 6);
        ProfileScope ps("cin_last_pool_kernel", st);
        const int n_mt = (Hp8 + 15) / 16;
        cin_last_pool_mma_kernel<26><<<grid, 32 * kPoolMmaWarps, (size_t)n_mt * 2 * 32 * 16, st>>>(z);
      } else {
        const long long tiles = (rows + kPoolThreads * kPoolRows - 1) / (kPoolThreads * kPoolRows);
        const int grid = (int)std::min<long long>(tiles, (long long)sms * 2);
        ProfileScope ps("cin_last_pool_kernel", st);
        cin_last_pool_kernel<26><<<grid, kPoolThreads, (size_t)Hp8 * kPoolTRow * 4, st>>>(z);
      }
      KON_LAUNCH_CHECK("cin_last_pool_kernel");
      continue;
    }
    FwdArgs a;
    a.x0 = x0;
    a.x0_sb = x0_sb;
    a.pre = l == 0 ? nullptr : reinterpret_cast<const unsigned short*>(sv + L.zt_off[l - 1]);
    a.wpack = ws + L.wpack_off[l];
    a.bias = bias[l];
    a.zt = reinterpret_cast<unsigned short*>(sv + L.zt_off[l]);
    a.pooled = pooled;
    a.pooled_stride = nl * D;
    a.pooled_col0 = l * D;
    a.rows = rows;
    a.D = D;
    a.Hp = L.Hp[l];
    a.N = L.N[l];
    a.N8 = L.N8[l];
    a.nk = L.nk[l];
    a.kblk = 2u * L.N8[l] * 128u;
    a.n_pairs = (rows + 128 * kSub - 1) / (128 * kSub);
    const int grid = (int)std::min<long long>(a.n_pairs, sms);
    {
      ProfileScope ps("cin_fwd_tc_kernel", st);
      cin_fwd_tc_kernel<26><<<grid, kFwdThreads, smem, st>>>(a);
    }
    KON_LAUNCH_CHECK("cin_fwd_tc_kernel");
  }
  return KON_OK;
}

}  // namespace kon
