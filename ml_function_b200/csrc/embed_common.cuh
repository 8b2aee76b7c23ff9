// Shared by embed.cu (gather forward, row-wise optimizers) and embed_bwd.cu (routing + segmented sum):
// the per-field row offsets of an arena and the [B,F] / [B,F,L] id view of a call.
#pragma once

#include "common.cuh"

namespace kon {

constexpr int kMaxFields = 256;
constexpr int kMaxBagLen = 512;

struct FieldTable {
  int64_t off[kMaxFields + 1];
};

struct IdsView {
  int64_t B, F, L;
  bool i64;
};

inline int parse_common(const DLTensor* ids, const int64_t* field_row_offset, int32_t n_fields,
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

inline int pow2_ge(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace kon
