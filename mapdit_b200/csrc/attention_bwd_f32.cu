// CUDA-core backward of cosine attention (fp32 math; fp32 parity mode and generic fallback).
//   P = softmax(q k^T * scale) recomputed from the saved log-sum-exp L;  delta_i = dO_i . O_i
//   dV_j = sum_i P_ij dO_i ;  dS_ij = P_ij (dO_i . V_j - delta_i) ;  dQ_i = scale sum_j dS_ij K_j ;  dK_j = scale sum_i dS_ij Q_i
// Kernel A: one thread per query row (dQ, delta).  Kernel B: two threads per key row, each owning half of the
// head dimension (dK, dV).  Both are deterministic (no atomics).
#include "common.cuh"

namespace {
constexpr int KB = 32;

template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ o, const T* __restrict__ dout,
                                                          const float* __restrict__ lse, T* __restrict__ dqkv,
                                                          float* __restrict__ delta, int tokens, int heads, float scale) {
  __shared__ __align__(16) float Ks[KB][HD];
  __shared__ __align__(16) float Vs[KB][HD];
  const int n = blockIdx.z, h = blockIdx.y, D = heads * HD;
  const int qi = blockIdx.x * 128 + threadIdx.x;
  const bool active = qi < tokens;
  const size_t row = (size_t)n * tokens + (active ? qi : 0);
  const T* base = qkv + (size_t)n * tokens * 3 * D;
  float q[HD], g[HD], dq[HD];
  float dl = 0.f;
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    q[d] = active ? ld_act(qkv + row * 3 * D + h * HD + d) * scale : 0.f;
    g[d] = active ? ld_act(dout + row * D + h * HD + d) : 0.f;
    float ov = active ? ld_act(o + row * D + h * HD + d) : 0.f;
    dl = fmaf(g[d], ov, dl);
    dq[d] = 0.f;
  }
  const float L = active ? lse[row * heads + h] : 0.f;
  if (active) delta[row * heads + h] = dl;
  for (int k0 = 0; k0 < tokens; k0 += KB) {
    __syncthreads();
    for (int i = threadIdx.x; i < KB * HD; i += 128) {
      int j = i / HD, d = i - j * HD;
      int key = k0 + j;
      Ks[j][d] = key < tokens ? ld_act(base + (size_t)key * 3 * D + D + h * HD + d) : 0.f;
      Vs[j][d] = key < tokens ? ld_act(base + (size_t)key * 3 * D + 2 * D + h * HD + d) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < KB; ++j) {
      if (k0 + j >= tokens) break;
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
        float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d]);
        s = fmaf(q[d], kk.x, s); s = fmaf(q[d + 1], kk.y, s); s = fmaf(q[d + 2], kk.z, s); s = fmaf(q[d + 3], kk.w, s);
        dp = fmaf(g[d], vv.x, dp); dp = fmaf(g[d + 1], vv.y, dp); dp = fmaf(g[d + 2], vv.z, dp); dp = fmaf(g[d + 3], vv.w, dp);
      }
      float ds = expf(s - L) * (dp - dl);
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        float4 kk = *reinterpret_cast<const float4*>(&Ks[j][d]);
        dq[d] = fmaf(ds, kk.x, dq[d]); dq[d + 1] = fmaf(ds, kk.y, dq[d + 1]);
        dq[d + 2] = fmaf(ds, kk.z, dq[d + 2]); dq[d + 3] = fmaf(ds, kk.w, dq[d + 3]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int d = 0; d < HD; ++d) st_act(dqkv + row * 3 * D + h * HD + d, dq[d] * scale);
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ delta,
                                                           T* __restrict__ dqkv, int tokens, int heads, float scale) {
  constexpr int HH = HD / 2;
  __shared__ __align__(16) float Qs[KB][HD];
  __shared__ __align__(16) float Gs[KB][HD];
  __shared__ float Ls[KB], Ds[KB];
  const int n = blockIdx.z, h = blockIdx.y, D = heads * HD;
  const int half = threadIdx.x & 1;
  const int kj = blockIdx.x * 64 + (threadIdx.x >> 1);
  const bool active = kj < tokens;
  const T* base = qkv + (size_t)n * tokens * 3 * D;
  const size_t krow = (size_t)n * tokens + (active ? kj : 0);
  float kk[HH], vv[HH], dk[HH], dv[HH];
#pragma unroll
  for (int d = 0; d < HH; ++d) {
    kk[d] = active ? ld_act(qkv + krow * 3 * D + D + h * HD + half * HH + d) : 0.f;
    vv[d] = active ? ld_act(qkv + krow * 3 * D + 2 * D + h * HD + half * HH + d) : 0.f;
    dk[d] = 0.f;
    dv[d] = 0.f;
  }
  for (int q0 = 0; q0 < tokens; q0 += KB) {
    __syncthreads();
    for (int i = threadIdx.x; i < KB * HD; i += 128) {
      int j = i / HD, d = i - j * HD;
      int qi = q0 + j;
      size_t r = (size_t)n * tokens + qi;
      Qs[j][d] = qi < tokens ? ld_act(base + (size_t)qi * 3 * D + h * HD + d) * scale : 0.f;
      Gs[j][d] = qi < tokens ? ld_act(dout + r * D + h * HD + d) : 0.f;
    }
    if (threadIdx.x < KB) {
      int qi = q0 + threadIdx.x;
      size_t r = (size_t)n * tokens + qi;
      Ls[threadIdx.x] = qi < tokens ? lse[r * heads + h] : INFINITY;
      Ds[threadIdx.x] = qi < tokens ? delta[r * heads + h] : 0.f;
    }
    __syncthreads();
    for (int j = 0; j < KB; ++j) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < HH; ++d) {
        s = fmaf(Qs[j][half * HH + d], kk[d], s);
        dp = fmaf(Gs[j][half * HH + d], vv[d], dp);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      dp += __shfl_xor_sync(0xffffffffu, dp, 1);
      float p = expf(s - Ls[j]);  // 0 for padded queries (L = +inf)
      float ds = p * (dp - Ds[j]);
#pragma unroll
      for (int d = 0; d < HH; ++d) {
        dv[d] = fmaf(p, Gs[j][half * HH + d], dv[d]);
        dk[d] = fmaf(ds, Qs[j][half * HH + d], dk[d]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int d = 0; d < HH; ++d) {
      st_act(dqkv + krow * 3 * D + D + h * HD + half * HH + d, dk[d]);  // Qs already carries `scale`
      st_act(dqkv + krow * 3 * D + 2 * D + h * HD + half * HH + d, dv[d]);
    }
  }
}

template <typename T, int HD>
void launch(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta, int n, int tokens, int heads,
            cudaStream_t s) {
  float scale = 1.0f / sqrtf((float)HD);
  dim3 g1((tokens + 127) / 128, heads, n), g2((tokens + 63) / 64, heads, n);
  attn_bwd_dq_kernel<T, HD><<<g1, 128, 0, s>>>((const T*)qkv, (const T*)o, (const T*)dout, lse, (T*)dqkv, delta, tokens, heads, scale);
  attn_bwd_dkv_kernel<T, HD><<<g2, 128, 0, s>>>((const T*)qkv, (const T*)dout, lse, delta, (T*)dqkv, tokens, heads, scale);
}
}  // namespace

// dqkv[M, 3D] <- gradients w.r.t. the (normalised) q, k and v; delta is an [M, H] fp32 scratch.
int mapdit_attn_bwd_simt(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta,
                         int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream) {
  MAPDIT_REQUIRE(qkv && o && dout && lse && dqkv && delta && n_samples > 0 && tokens > 0, "cos_attn_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (head_dim == 64) {
    if (dtype == MAPDIT_F32) launch<float, 64>(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, s);
    else launch<bf16, 64>(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, s);
  } else if (head_dim == 72) {
    if (dtype == MAPDIT_F32) launch<float, 72>(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, s);
    else launch<bf16, 72>(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, s);
  } else {
    mapdit_set_error("cos_attn_bwd: unsupported head_dim %d", head_dim);
    return MAPDIT_ERR_UNSUPPORTED;
  }
  MAPDIT_LAUNCH_CHECK("cos_attn_bwd");
  mapdit_count_launch();
  return MAPDIT_OK;
}
