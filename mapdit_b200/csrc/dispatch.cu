// Entry points that choose between the tcgen05 kernels and the CUDA-core kernels.
#include "common.cuh"

int mapdit_attn_simt_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, int dtype, void* stream);
int mapdit_attn_tc_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream);
bool mapdit_attn_tc_supported(int tokens, int hd);

extern "C" int mapdit_cos_attn_fwd(const void* qkv, void* o, float* lse, int n_samples, int tokens, int heads, int head_dim, int dtype,
                                   void* stream) {
  MAPDIT_REQUIRE(qkv && o && n_samples > 0 && tokens > 0 && heads > 0, "cos_attn_fwd: bad args");
  // the tcgen05 kernel relies on the cosine-attention logit bound (fixed softmax max); plain dot-product attention
  // (use_cosine_attention=False) takes the running-max CUDA-core kernel
  if (dtype == MAPDIT_BF16 && mapdit_attn_tc_supported(tokens, head_dim) && !(mapdit_variant() & MAPDIT_VAR_DOT_ATTN))
    return mapdit_attn_tc_fwd(qkv, o, lse, n_samples, tokens, heads, head_dim, stream);
  return mapdit_attn_simt_fwd(qkv, o, lse, n_samples, tokens, heads, head_dim, dtype, stream);
}
