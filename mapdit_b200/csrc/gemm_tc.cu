#include "common.cuh"
extern "C" int mapdit_gemm_bf16(const mapdit_gemm_args*, void*) {
  mapdit_set_error("gemm_bf16: not built");
  return MAPDIT_ERR_UNSUPPORTED;
}
