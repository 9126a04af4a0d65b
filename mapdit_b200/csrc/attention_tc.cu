// placeholder until the tcgen05 attention kernel lands (dispatch falls back to the CUDA-core kernel)
#include "common.cuh"
bool mapdit_attn_tc_supported(int, int) { return false; }
int mapdit_attn_tc_fwd(const void*, void*, int, int, int, int, void*) {
  mapdit_set_error("attn_tc_fwd: not built");
  return MAPDIT_ERR_UNSUPPORTED;
}
