// Shared between attn.cu (fp32 CUDA-core path) and attn_tc.cu (bf16 tensor-core path).
#pragma once
#include "common.cuh"

namespace kon {

struct AttnDims {
  long long B;
  int F, kin, H;
  int use_scale, use_ln, use_res, relu;
  float ln_eps;
  // element strides of y / gy over (head, sample, field); the last dim (d) is compact.  Compact [H,B,F,d]
  // unless the caller hands in a permuted window (bf16 path only): e.g. [B,F,H,d] so that the next
  // attention layer reads [B,F,H*d] in place, or [B,H,F,d] = the flattened heads MergeScoreLayer reads.
  long long ysh, ysb, ysf, gsh, gsb, gsf;
};

// sigmoid with the fast exponential and reciprocal (relative error ~2e-7: two orders below the 1e-5 gate)

// bf16 mma.sync path (KON_ATTN_BF16): F <= 32, kin in {16,32,48,64}, d == 8
bool attn_tc_supported(const AttnDims& p, int DH);
bool attn_tc_bwd_supported(const AttnDims& p, int DH);
int attn_tc_bwd(const float* x, const float* wq, const float* wk, const float* wr, const float* gamma,
                const float* beta, const float* gy, float* dx, float* partial, int max_grid,
                const AttnDims& p, int sms, int* grid_used, cudaStream_t st);
int attn_tc_fwd(const float* x, const float* wq, const float* wk, const float* wr, const float* gamma,
                const float* beta, float* y, const AttnDims& p, int sms, cudaStream_t st);

}  // namespace kon
