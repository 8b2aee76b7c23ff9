// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace kon {
size_t cin_tc_saved_bytes(int64_t, int, int, const int32_t*, int) { return 4; }
size_t cin_tc_workspace_bytes(int64_t, int, int, const int32_t*, int, int) { return 4; }
int cin_tc_fwd(const float*, const float* const*, const float* const*, int, const int32_t*, int64_t,
               int, int, float*, void*, void*, int, cudaStream_t) {
  return fail(KON_EUNSUPPORTED, "KON_CIN_BF16 not built yet");
}
int cin_tc_bwd(const float*, const float* const*, const float* const*, int, const int32_t*, int64_t,
               int, int, const float*, const void*, float*, float* const*, float* const*, void*, int,
               cudaStream_t) {
  return fail(KON_EUNSUPPORTED, "KON_CIN_BF16 not built yet");
}
}  // namespace kon
