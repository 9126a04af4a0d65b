// Entry points that choose between the tcgen05 kernels and the CUDA-core kernels.
#include "common.cuh"

int mapdit_attn_simt_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, int dtype, void* stream);
int mapdit_attn_tc_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream);
bool mapdit_attn_tc_supported(int tokens, int hd);
int mapdit_attn_tc2_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream);
bool mapdit_attn_tc2_supported(int tokens, int hd);
int mapdit_attn_mma_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, void* stream);
bool mapdit_attn_mma_supported(int tokens, int hd);
int g_mapdit_attn_v2 = 1;  // runtime option "attn_v2": the one-CTA-per-SM ping-pong kernel where the shape qualifies

extern "C" int mapdit_cos_attn_fwd(const void* qkv, void* o, float* lse, int n_samples, int tokens, int heads, int head_dim, int dtype,
                                   void* stream) {
  MAPDIT_REQUIRE(qkv && o && n_samples > 0 && tokens > 0 && heads > 0, "cos_attn_fwd: bad args");
  // the tcgen05 kernel relies on the cosine-attention logit bound (fixed softmax max); plain dot-product attention
  // (use_cosine_attention=False) takes the running-max CUDA-core kernel
  const bool cosine = !(mapdit_variant() & MAPDIT_VAR_DOT_ATTN);
  if (dtype == MAPDIT_BF16 && cosine && g_mapdit_attn_v2 && mapdit_attn_tc2_supported(tokens, head_dim))
    return mapdit_attn_tc2_fwd(qkv, o, lse, n_samples, tokens, heads, head_dim, stream);
  if (dtype == MAPDIT_BF16 && cosine && mapdit_attn_tc_supported(tokens, head_dim))
    return mapdit_attn_tc_fwd(qkv, o, lse, n_samples, tokens, heads, head_dim, stream);
  // head_dim 72 (DiT-XL) and plain dot-product attention: warp-level tensor-core MMAs with a running max
  if (dtype == MAPDIT_BF16 && mapdit_attn_mma_supported(tokens, head_dim))
    return mapdit_attn_mma_fwd(qkv, o, lse, n_samples, tokens, heads, head_dim, stream);
  return mapdit_attn_simt_fwd(qkv, o, lse, n_samples, tokens, heads, head_dim, dtype, stream);
}
