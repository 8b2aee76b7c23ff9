/* oracle/kon_oracle_c.c -- TEST INFRASTRUCTURE ONLY (see oracle/kon_oracle.py): a plain-C restatement of the
 * integer / byte-exact parts of the hot path and of the fp32 op ORDER of the FM and cross layers, independent
 * of torch.  It cross-checks the torch-CPU oracle (tests/test_oracle_cpu.py) and is never linked into, imported
 * by, or executed from the product (ml_function_b200/).  PARITY UNPINNED: like kon_oracle.py it restates the
 * reference's TensorFlow op stream (TensorFlow is not installable here), it is not output of the reference.
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: every product and sum is rounded to
 * fp32 exactly where the reference's separate Mul / AddV2 ops round).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* SparseEmbed.call, IL:225-242: out[b,f,:] = T_f[ids[b,f], :]  (Keras Embedding = gather after an int cast).
 * tables: all fields back to back [R,k]; offs[F+1] row offsets; ids [B,F] int32; out [B,F,k]. */
void kon_c_embed_gather(const float* tables, const int64_t* offs, const int32_t* ids, int64_t B, int F, int k,
                        float* out) {
  for (int64_t b = 0; b < B; ++b)
    for (int f = 0; f < F; ++f)
      memcpy(out + (b * F + f) * k, tables + (offs[f] + ids[b * F + f]) * k, sizeof(float) * (size_t)k);
}

/* Embedding backward (implicit in Model.fit: IndexedSlices -> unique + unsorted_segment_sum): per arena row the
 * gradient rows are added in ascending sample order; the touched rows come out ascending.
 * d_out [B,F,k]; unique_rows [<=B*F]; grads [<=B*F,k]; returns the number of unique rows. */
int64_t kon_c_embed_grad(const int64_t* offs, const int32_t* ids, const float* d_out, int64_t B, int F, int k,
                         int32_t* unique_rows, float* grads) {
  const int64_t R = offs[F];
  float* dense = (float*)calloc((size_t)R * k, sizeof(float));
  unsigned char* hit = (unsigned char*)calloc((size_t)R, 1);
  for (int f = 0; f < F; ++f)                /* one table at a time, samples ascending (= np.add.at order) */
    for (int64_t b = 0; b < B; ++b) {
      const int64_t r = offs[f] + ids[b * F + f];
      hit[r] = 1;
      for (int c = 0; c < k; ++c) dense[r * k + c] += d_out[(b * F + f) * k + c];
    }
  int64_t n = 0;
  for (int64_t r = 0; r < R; ++r)
    if (hit[r]) {
      unique_rows[n] = (int32_t)r;
      memcpy(grads + n * k, dense + r * k, sizeof(float) * (size_t)k);
      ++n;
    }
  free(dense);
  free(hit);
  return n;
}

/* FmLayer.call (IL:161-170) over InnerLayer.call (IL:59-66) in the reference's op order: the pairwise products
 * v_i*v_j in itertools.combinations order are Keras-Add-ed left to right (((p01 + p02) + p03) + ...), then the
 * linear terms are added one by one.  v [B,F,k], lin [B,F], out [B,k]. */
void kon_c_fm(const float* v, const float* lin, int64_t B, int F, int k, float* out) {
  for (int64_t b = 0; b < B; ++b)
    for (int c = 0; c < k; ++c) {
      float acc = 0.f;
      int first = 1;
      for (int i = 0; i < F; ++i)
        for (int j = i + 1; j < F; ++j) {
          const float p = v[(b * F + i) * k + c] * v[(b * F + j) * k + c];
          acc = first ? p : acc + p;
          first = 0;
        }
      if (lin)
        for (int f = 0; f < F; ++f) acc = acc + lin[b * F + f];
      out[b * k + c] = acc;
    }
}

/* CrossLayer.call (IL:275-282): x_{l+1} = x0 * (x_l . w_l) + x_l + b_l.  The dot is accumulated in double
 * (the reference's MatMul order is a BLAS detail); the three-term update rounds like the reference's
 * BatchMatMul, AddV2, AddV2.  x0 [B,D], w, b [L,D], out [B,D]. */
void kon_c_cross(const float* x0, const float* w, const float* b, int64_t B, int D, int L, float* out) {
  float* xl = (float*)malloc(sizeof(float) * (size_t)D);
  for (int64_t r = 0; r < B; ++r) {
    const float* x = x0 + r * D;
    memcpy(xl, x, sizeof(float) * (size_t)D);
    for (int l = 0; l < L; ++l) {
      double s = 0.0;
      for (int d = 0; d < D; ++d) s += (double)xl[d] * (double)w[l * D + d];
      const float sf = (float)s;
      for (int d = 0; d < D; ++d) {
        const float t = x[d] * sf;
        const float u = t + xl[d];
        xl[d] = u + b[l * D + d];
      }
    }
    memcpy(out + r * D, xl, sizeof(float) * (size_t)D);
  }
  free(xl);
}

/* CIN.call (IL:310-327), per layer l:  z[b,d,o] = sum_{h,i} W_l[h*m + i, o] * pre[b,h,d] * x0[b,i,d] + bias_l[o]
 * (outer product over the field axes per embedding coordinate d, IL:316; channel index h*m + i from the transpose
 * + reshape at IL:317-318; Conv1D(kernel 1) = per-position Dense, IL:308), pre' = z^T (IL:320), pooled over the
 * feature maps o (IL:322).  x0 [B,m,D]; w[l] [H_{l-1}*m, H_l] concatenated in `w`; bias concatenated in `bias`;
 * hs[L] layer sizes; pooled [B, L*D].  Accumulation in double (the BatchMatMul / Conv1D orders are BLAS details). */
void kon_c_cin(const float* x0, const float* w, const float* bias, const int32_t* hs, int L, int64_t B, int m, int D,
               float* pooled) {
  int hmax = m;
  for (int l = 0; l < L; ++l)
    if (hs[l] > hmax) hmax = hs[l];
  float* pre = (float*)malloc(sizeof(float) * (size_t)hmax * D);
  float* nxt = (float*)malloc(sizeof(float) * (size_t)hmax * D);
  for (int64_t b = 0; b < B; ++b) {
    const float* xb = x0 + b * m * D;
    int H = m;
    memcpy(pre, xb, sizeof(float) * (size_t)m * D);
    const float* wl = w;
    const float* bl = bias;
    for (int l = 0; l < L; ++l) {
      const int N = hs[l];
      for (int d = 0; d < D; ++d) {
        double pool = 0.0;
        for (int o = 0; o < N; ++o) {
          double z = 0.0;
          for (int h = 0; h < H; ++h)
            for (int i = 0; i < m; ++i)
              z += (double)wl[(size_t)(h * m + i) * N + o] * (double)pre[h * D + d] * (double)xb[i * D + d];
          const float zf = (float)(z + (double)bl[o]);
          nxt[o * D + d] = zf;
          pool += (double)zf;
        }
        pooled[b * (int64_t)L * D + (int64_t)l * D + d] = (float)pool;
      }
      wl += (size_t)H * m * N;
      bl += N;
      H = N;
      memcpy(pre, nxt, sizeof(float) * (size_t)N * D);
    }
  }
  free(pre);
  free(nxt);
}

/* The AutoInt block: MultHeadAttentionLayer.call (BL:356-377) + ProductAttentionLayer.call (BL:292-311) wrapped by
 * DnnLayer.call (CL:201-226):  q = X Wq, k = X Wk, v = X Wk (BL:360: key_w, value_w is never read),
 * score = sigmoid(q k^T / sqrt(d)) (BL:296-297, 286), a = score v, LayerNorm over d (eps 1e-3, biased variance,
 * BL:331), y = ReLU(X Wr + a) (CL:212, 216).  x [B,F,kin]; w* [kin,H,d]; y [H,B,F,d]. */
#include <math.h>
void kon_c_autoint_block(const float* x, const float* wq, const float* wk, const float* wr, const float* gamma,
                         const float* beta, int64_t B, int F, int kin, int H, int d, float* y) {
  double* q = (double*)malloc(sizeof(double) * (size_t)F * d * 3);
  double* k = q + (size_t)F * d;
  double* r = k + (size_t)F * d;
  const double scale = 1.0 / sqrt((double)d);
  for (int64_t b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      for (int f = 0; f < F; ++f)
        for (int e = 0; e < d; ++e) {
          double sq = 0, sk = 0, sr = 0;
          for (int c = 0; c < kin; ++c) {
            const double xv = x[(b * F + f) * kin + c];
            sq += xv * wq[(c * H + h) * d + e];
            sk += xv * wk[(c * H + h) * d + e];
            sr += xv * wr[(c * H + h) * d + e];
          }
          q[f * d + e] = sq; k[f * d + e] = sk; r[f * d + e] = sr;
        }
      for (int i = 0; i < F; ++i) {
        double a[64];
        for (int e = 0; e < d; ++e) a[e] = 0;
        for (int j = 0; j < F; ++j) {
          double s = 0;
          for (int e = 0; e < d; ++e) s += q[i * d + e] * k[j * d + e];
          const double p = 1.0 / (1.0 + exp(-s * scale));
          for (int e = 0; e < d; ++e) a[e] += p * k[j * d + e];
        }
        double mean = 0, var = 0;
        for (int e = 0; e < d; ++e) mean += a[e];
        mean /= d;
        for (int e = 0; e < d; ++e) var += (a[e] - mean) * (a[e] - mean);
        var /= d;
        const double rstd = 1.0 / sqrt(var + 1e-3);
        for (int e = 0; e < d; ++e) {
          const double v = (a[e] - mean) * rstd * gamma[e] + beta[e] + r[i * d + e];
          y[(((int64_t)h * B + b) * F + i) * d + e] = (float)(v > 0 ? v : 0);
        }
      }
    }
  free(q);
}
