// CUDA-core flash attention used by the fp32 parity mode (mode a) and as the generic kernel for
// shapes the tcgen05 kernel does not cover.  qkv is [N*T, 3D] with q|k|v column thirds and
// head-major features (src/layers/attention.py:37-41); q,k are already L2-normalised, so the
// logits are q·k/sqrt(hd) (attention.py:47).  One thread owns one query row: q and the output
// accumulator live in registers, K/V stream through shared memory in 32-key blocks and are read
// as warp-broadcast float4s.  Softmax is the exact online form (running max), fp32 throughout.
#include "common.cuh"

namespace {
constexpr int KB = 32;  // keys per smem block

template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ o, float* __restrict__ lse,
                                                        int tokens, int heads, float scale) {
  __shared__ __align__(16) float Ks[KB][HD];
  __shared__ __align__(16) float Vs[KB][HD];
  const int n = blockIdx.z, h = blockIdx.y;
  const int D = heads * HD;
  const int qi = blockIdx.x * 128 + threadIdx.x;
  const bool active = qi < tokens;
  const T* base = qkv + (size_t)n * tokens * 3 * D;
  float q[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    q[d] = active ? ld_act(base + (size_t)qi * 3 * D + h * HD + d) * scale : 0.f;
    acc[d] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int k0 = 0; k0 < tokens; k0 += KB) {
    __syncthreads();
    for (int i = threadIdx.x; i < KB * HD; i += 128) {
      int j = i / HD, d = i - j * HD;
      int key = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (key < tokens) {
        kv = ld_act(base + (size_t)key * 3 * D + D + h * HD + d);
        vv = ld_act(base + (size_t)key * 3 * D + 2 * D + h * HD + d);
      }
      Ks[j][d] = kv;
      Vs[j][d] = vv;
    }
    __syncthreads();
    float s[KB];
    float bmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
        a = __fmaf_rn(q[d], kk.x, a);
        a = __fmaf_rn(q[d + 1], kk.y, a);
        a = __fmaf_rn(q[d + 2], kk.z, a);
        a = __fmaf_rn(q[d + 3], kk.w, a);
      }
      s[j] = (k0 + j < tokens) ? a : -INFINITY;
      bmax = fmaxf(bmax, s[j]);
    }
    float mnew = fmaxf(mrun, bmax);
    float corr = expf(mrun - mnew);  // first block: exp(-inf) = 0
    lrun *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      float p = expf(s[j] - mnew);
      lrun += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
        acc[d] = __fmaf_rn(p, vv.x, acc[d]);
        acc[d + 1] = __fmaf_rn(p, vv.y, acc[d + 1]);
        acc[d + 2] = __fmaf_rn(p, vv.z, acc[d + 2]);
        acc[d + 3] = __fmaf_rn(p, vv.w, acc[d + 3]);
      }
    }
    mrun = mnew;
  }
  if (active) {
    float inv = 1.0f / lrun;
    T* orow = o + ((size_t)n * tokens + qi) * D + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) st_act(orow + d, acc[d] * inv);
    if (lse) lse[((size_t)n * tokens + qi) * heads + h] = mrun + logf(lrun);
  }
}

template <typename T>
int launch(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, cudaStream_t s) {
  dim3 grid((tokens + 127) / 128, heads, n);
  float scale = 1.0f / sqrtf((float)hd);
  switch (hd) {
    case 64: attn_simt_kernel<T, 64><<<grid, 128, 0, s>>>((const T*)qkv, (T*)o, lse, tokens, heads, scale); break;
    case 72: attn_simt_kernel<T, 72><<<grid, 128, 0, s>>>((const T*)qkv, (T*)o, lse, tokens, heads, scale); break;
    case 32: attn_simt_kernel<T, 32><<<grid, 128, 0, s>>>((const T*)qkv, (T*)o, lse, tokens, heads, scale); break;
    default: mapdit_set_error("cos_attn_fwd: unsupported head_dim %d", hd); return MAPDIT_ERR_UNSUPPORTED;
  }
  return MAPDIT_OK;
}
}  // namespace

int mapdit_attn_simt_fwd(const void* qkv, void* o, float* lse, int n, int tokens, int heads, int hd, int dtype, void* stream) {
  int rc = (dtype == MAPDIT_F32) ? launch<float>(qkv, o, lse, n, tokens, heads, hd, (cudaStream_t)stream)
                                 : launch<bf16>(qkv, o, lse, n, tokens, heads, hd, (cudaStream_t)stream);
  if (rc != MAPDIT_OK) return rc;
  MAPDIT_LAUNCH_CHECK("attn_simt_fwd");
  return MAPDIT_OK;
}
