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
#include <cuda_bf16.h>

#include "tc_ptx.cuh"

namespace kon {

namespace {

#ifndef KON_TC_SUB
#define KON_TC_SUB 2
#endif
constexpr int kSub = KON_TC_SUB;           // 128-row sub-tiles per CTA sharing every B stage
constexpr int kProdWarps = 4 * kSub;
constexpr int kTcThreads = 32 * (kProdWarps + 2);
#ifndef KON_TC_G
#define KON_TC_G 3
#endif
#ifndef KON_TC_NS
#define KON_TC_NS (KON_TC_SUB == 2 ? 2 : 12)
#endif
#ifndef KON_TC_S
#define KON_TC_S 8
#endif
constexpr int kG = KON_TC_G;     // k-steps (of 16) per A slot and per B stage
constexpr int kNS = KON_TC_NS;   // A slots per sub-tile (8*kG*kNS <= 48 TMEM columns)
constexpr int kS = KON_TC_S;     // B stages
constexpr int kMaxN = 208;
// TMEM columns: kSub == 2: D0 [0,208) A0 [208,256) D1 [256,464) A1 [464,512);  kSub == 1: D0 [0,208) A0 [208,496)
constexpr uint32_t kColD0 = 0, kColA0 = 208, kColD1 = 256, kColA1 = 464;
static_assert(8 * kG * kNS <= (kSub == 2 ? 48 : 304), "A ring does not fit in tensor memory");

__host__ __device__ constexpr int gcd_(int a, int b) { return b == 0 ? a : gcd_(b, a % b); }
__host__ __device__ constexpr int lcm_(int a, int b) { return a / gcd_(a, b) * b; }

__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf16_bcast(float x) {
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  const uint32_t u = *reinterpret_cast<const unsigned short*>(&h);
  return u | (u << 16);
}
__device__ __forceinline__ uint32_t bf16_bcast_raw(unsigned short u) { return (uint32_t)u | ((uint32_t)u << 16); }
__device__ __forceinline__ float bf16_to_f32(unsigned short u) { return __uint_as_float((uint32_t)u << 16); }

struct Barriers {
  uint64_t b_full[kS], b_empty[kS];
  uint64_t a_full[kSub][kNS], a_empty[kSub][kNS];
  uint64_t d_full;
  uint32_t tmem_base;
};

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
  const float* x0;              // [B, m, D] fp32 compact
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

template <int MF>
__global__ void __launch_bounds__(kTcThreads, 1) cin_fwd_tc_kernel(const FwdArgs a) {
  constexpr int LCM = lcm_(16, MF);
  constexpr int PK = LCM / 16;   // k-steps per period
  constexpr int PH = LCM / MF;   // feature maps (h) per period
  static_assert(MF % 2 == 0, "field count must be even (bf16x2 pairs never straddle an h)");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ Barriers bars;
  __shared__ float s_bias[kMaxN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kMaxN; i += kTcThreads) s_bias[i] = i < a.N ? a.bias[i] : 0.f;

  if (tid == 0) {
    for (int i = 0; i < kS; ++i) { mbar_init(&bars.b_full[i], 1); mbar_init(&bars.b_empty[i], 1); }
    for (int s = 0; s < kSub; ++s)
      for (int i = 0; i < kNS; ++i) { mbar_init(&bars.a_full[s][i], 4); mbar_init(&bars.a_empty[s][i], 1); }
    mbar_init(&bars.d_full, 1);
    fence_mbar_init();
  }
  if (warp == kProdWarps) tc::tmem_alloc(&bars.tmem_base, 512);
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem = bars.tmem_base;

  const int n_groups = (a.nk + kG - 1) / kG;
  const uint32_t stage_bytes = kG * a.kblk;

  if (warp < kProdWarps) {
    // ================= producers + epilogue =================================================
    const int sub = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t colD = sub ? kColD1 : kColD0;
    const uint32_t colA = sub ? kColA1 : kColA0;
    const int n_periods = (a.Hp + PH - 1) / PH;
    uint32_t a_it = 0;   // A-slot uses so far (ring position)
    uint32_t tile_it = 0;
    for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x, ++tile_it) {
      const long long r = pair * (128 * kSub) + sub * 128 + (warp & 3) * 32 + lane;
      const bool valid = r < a.rows;
      const long long b = valid ? r / a.D : 0;
      const int d = valid ? (int)(r - b * a.D) : 0;
      const float* xrow = a.x0 + b * (long long)MF * a.D + d;          // x0[b,i,d] = xrow[i*D]
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
      int cnt = 0;           // k-steps written into the current slot
      for (int per = 0; per < n_periods; ++per) {
#pragma unroll
        for (int j = 0; j < PH; ++j) cur[j] = to_bcast(nraw[j], per * PH + j);
#pragma unroll
        for (int j = 0; j < PH; ++j) nraw[j] = load_raw((per + 1) * PH + j);
#pragma unroll
        for (int j = 0; j < PK; ++j) {
          if (ks_global < a.nk) {
            const uint32_t slot = a_it % kNS;
            if (cnt == 0) {
              mbar_wait(&bars.a_empty[sub][slot], ((a_it / kNS) & 1) ^ 1);
              tc::fence_after();
            }
            uint32_t w[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              constexpr int dummy = 0;
              (void)dummy;
              const int cl = 16 * j + 2 * q;            // compile-time after unrolling
              w[q] = hmul2_bf16(cur[cl / MF], x2[(cl % MF) / 2]);
            }
            tc::st8(tmem + lane_base + colA + slot * (8 * kG) + 8 * cnt, w);
            ++cnt;
            ++ks_global;
            if (cnt == kG || ks_global == a.nk) {
              tc::wait_st();
              tc::fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bars.a_full[sub][slot]);
              cnt = 0;
              ++a_it;
            }
          }
        }
      }
      // ---- epilogue: z = D + bias; pooled = sum_o z; z^T (bf16) -> zt ---------------------
      mbar_wait(&bars.d_full, tile_it & 1);
      tc::fence_after();
      float rsum = 0.f;
      unsigned short* zrow = a.zt + b * (long long)a.N * a.D + d;       // zt[b,o,d] = zrow[o*D]
      const int n_full = a.N & ~15;
      for (int o0 = 0; o0 < n_full; o0 += 16) {
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
      for (int o0 = n_full; o0 < a.N; o0 += 8) {     // tail: N % 16 in {8} or ragged
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
      if (valid) a.pooled[b * a.pooled_stride + a.pooled_col0 + d] = rsum;
      tc::fence_before();   // our tcgen05.ld are complete (wait_ld) before the next tile's a_full arrive
    }
  } else if (warp == kProdWarps) {
    // ================= MMA issuer ===========================================================
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_bf16(128, a.N8 * 8, 0, 0);
      const uint32_t lbo = (uint32_t)a.N8 * 128, sbo = 128;
      uint32_t a_it = 0, b_it = 0;
      for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        for (int g = 0; g < n_groups; ++g) {
          const int gk = min(kG, a.nk - g * kG);
          const uint32_t bs = b_it % kS;
          mbar_wait(&bars.b_full[bs], (b_it / kS) & 1);
          tc::fence_after();
          const uint64_t bdesc0 = tc::smem_desc(smem_u32(smem + bs * stage_bytes), lbo, sbo);
          const uint32_t slot = a_it % kNS;
          const uint32_t apar = (a_it / kNS) & 1;
#pragma unroll
          for (int sub = 0; sub < kSub; ++sub) {
            mbar_wait(&bars.a_full[sub][slot], apar);
            tc::fence_after();
            const uint32_t dcol = tmem + (sub ? kColD1 : kColD0);
            const uint32_t acol = tmem + (sub ? kColA1 : kColA0) + slot * (8 * kG);
            for (int ks = 0; ks < gk; ++ks)
              tc::mma_ts(dcol, acol + 8 * ks, bdesc0 + (uint64_t)((ks * a.kblk) >> 4), idesc,
                         (g > 0 || ks > 0) ? 1u : 0u);
            tc::commit(&bars.a_empty[sub][slot]);
          }
          tc::commit(&bars.b_empty[bs]);
          ++a_it;
          ++b_it;
        }
        tc::commit(&bars.d_full);
      }
    }
  } else {
    // ================= B loader =============================================================
    if (lane == 0) {
      uint32_t b_it = 0;
      for (long long pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        for (int g = 0; g < n_groups; ++g) {
          const int gk = min(kG, a.nk - g * kG);
          const uint32_t bs = b_it % kS;
          mbar_wait(&bars.b_empty[bs], ((b_it / kS) & 1) ^ 1);
          const uint32_t bytes = gk * a.kblk;
          mbar_expect_tx(&bars.b_full[bs], bytes);
          bulk_g2s(smem + bs * stage_bytes, a.wpack + (size_t)g * stage_bytes, bytes, &bars.b_full[bs]);
          ++b_it;
        }
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == kProdWarps) tc::tmem_dealloc(tmem, 512);
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

struct TcLayout {
  int nl;
  int Hp[KON_CIN_MAX_LAYERS], N[KON_CIN_MAX_LAYERS], N8[KON_CIN_MAX_LAYERS], nk[KON_CIN_MAX_LAYERS];
  size_t wpack_off[KON_CIN_MAX_LAYERS], wpack_bytes[KON_CIN_MAX_LAYERS];
  size_t zt_off[KON_CIN_MAX_LAYERS];
  size_t saved_total, work_total;
};

int tc_layout(int64_t B, int m, int D, const int32_t* hs, int nl, TcLayout* L) {
  if (m != 26) return -1;
  const int LCM = lcm_(16, m), PH = LCM / m, PK = LCM / 16;
  L->nl = nl;
  size_t so = 0, wo = 0;
  int hp = m;
  for (int l = 0; l < nl; ++l) {
    if (hs[l] > kMaxN) return -2;
    L->Hp[l] = hp;
    L->N[l] = hs[l];
    L->N8[l] = (hs[l] + 7) / 8;
    L->nk[l] = (hp + PH - 1) / PH * PK;
    L->wpack_bytes[l] = (size_t)L->nk[l] * 2 * L->N8[l] * 128;
    L->wpack_off[l] = wo;
    wo = align256(wo + L->wpack_bytes[l]);
    L->zt_off[l] = so;
    so = align256(so + (size_t)B * hs[l] * D * 2);
    hp = hs[l];
  }
  L->saved_total = so ? so : 256;
  L->work_total = wo ? wo : 256;
  return 0;
}

}  // namespace

size_t cin_tc_saved_bytes(int64_t B, int m, int D, const int32_t* hs, int nl) {
  TcLayout L;
  if (tc_layout(B, m, D, hs, nl, &L) != 0) return 256;
  return L.saved_total;
}

size_t cin_tc_workspace_bytes(int64_t B, int m, int D, const int32_t* hs, int nl, int sms) {
  TcLayout L;
  if (tc_layout(B, m, D, hs, nl, &L) != 0) return 256;
  return L.work_total;
}

int cin_tc_fwd(const float* x0, const float* const* w, const float* const* bias, int nl,
               const int32_t* hs, int64_t B, int m, int D, float* pooled, void* saved,
               void* workspace, int sms, cudaStream_t st) {
  TcLayout L;
  const int rc = tc_layout(B, m, D, hs, nl, &L);
  KON_REQUIRE(rc != -1, KON_EUNSUPPORTED, "KON_CIN_BF16 supports m = 26 fields (got %d); use KON_CIN_FP32", m);
  KON_REQUIRE(rc == 0, KON_EUNSUPPORTED, "KON_CIN_BF16 supports layer sizes <= %d", kMaxN);
  KON_REQUIRE(((uintptr_t)workspace & 255u) == 0 && ((uintptr_t)saved & 255u) == 0, KON_EINVAL,
              "saved / workspace must be 256-B aligned");
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  unsigned char* sv = static_cast<unsigned char*>(saved);
  const long long rows = B * D;
  const size_t smem = (size_t)kS * kG * 2 * ((kMaxN + 7) / 8) * 128;
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
  for (int l = 0; l < nl; ++l) {
    FwdArgs a;
    a.x0 = x0;
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
    cin_fwd_tc_kernel<26><<<grid, kTcThreads, smem, st>>>(a);
    KON_LAUNCH_CHECK("cin_fwd_tc_kernel");
  }
  return KON_OK;
}

int cin_tc_bwd(const float*, const float* const*, const float* const*, int, const int32_t*, int64_t,
               int, int, const float*, const void*, float*, float* const*, float* const*, void*, int,
               cudaStream_t) {
  return fail(KON_EUNSUPPORTED, "KON_CIN_BF16 backward not built yet");
}

}  // namespace kon
