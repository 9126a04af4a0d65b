// Fused Adam over a flat fp32 span ("next" row N1: train.py:57,96 -> torch.optim.Adam(lr, betas=(0.9, 0.99))).
// 16 B/param of state traffic + 4 B grad read: HBM-bound, 128-bit vectorised.
#include "common.cuh"

// bf16 gradients (the data-parallel path all-reduces a bf16 copy of the gradient span: half the NVLink bytes and half the SM time
// NCCL takes from the backward GEMMs); moments and parameters stay fp32
__global__ void __launch_bounds__(256) adam_g16_kernel(float* __restrict__ p, const bf16* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                                                       float bc2, float gs) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const int cnt = (int)(n - i4 < 4 ? n - i4 : 4);
  float gg[4] = {0.f, 0.f, 0.f, 0.f};
  if (cnt == 4 && (((uintptr_t)g & 7) == 0)) {
    const uint2 u = *reinterpret_cast<const uint2*>(g + i4);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    gg[0] = a.x; gg[1] = a.y; gg[2] = b.x; gg[3] = b.y;
  } else {
    for (int j = 0; j < cnt; ++j) gg[j] = __bfloat162float(g[i4 + j]);
  }
  if (cnt == 4 && ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0)) {
    float4 pv = *reinterpret_cast<float4*>(p + i4), mv = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
    float* pp = &pv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = gg[j] * gs;
      mm[j] = b1 * mm[j] + (1.0f - b1) * gr;
      vq[j] = b2 * vq[j] + (1.0f - b2) * gr * gr;
      pp[j] -= step * mm[j] / (sqrtf(vq[j]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(m + i4) = mv;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int j = 0; j < cnt; ++j) {
      const int64_t i = i4 + j;
      float gr = gg[j] * gs;
      float mi = b1 * m[i] + (1.0f - b1) * gr;
      float vi = b2 * v[i] + (1.0f - b2) * gr * gr;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float bc1,
                                                   float bc2, float gs) {
  const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  if (i4 + 4 <= n && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0)) {
    float4 pv = *reinterpret_cast<float4*>(p + i4), gv = *reinterpret_cast<const float4*>(g + i4);
    float4 mv = *reinterpret_cast<float4*>(m + i4), vv = *reinterpret_cast<float4*>(v + i4);
    float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gr = gg[j] * gs;
      mm[j] = b1 * mm[j] + (1.0f - b1) * gr;
      vq[j] = b2 * vq[j] + (1.0f - b2) * gr * gr;
      pp[j] -= step * mm[j] / (sqrtf(vq[j]) * inv_sqrt_bc2 + eps);
    }
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(m + i4) = mv;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < n && i < i4 + 4; ++i) {
      float gr = g[i] * gs;
      float mi = b1 * m[i] + (1.0f - b1) * gr;
      float vi = b2 * v[i] + (1.0f - b2) * gr * gr;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
  }
}

extern "C" int mapdit_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                                float bias_corr1, float bias_corr2, float grad_scale, void* stream) {
  MAPDIT_REQUIRE(p && g && m && v && n > 0, "adam_step: bad args");
  int64_t threads = (n + 3) / 4;
  adam_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, bias_corr1,
                                                                                    bias_corr2, grad_scale);
  MAPDIT_LAUNCH_CHECK("adam_step");
  return MAPDIT_OK;
}

extern "C" int mapdit_adam_step_g16(float* p, const void* g_bf16, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                    float eps, float bias_corr1, float bias_corr2, float grad_scale, void* stream) {
  MAPDIT_REQUIRE(p && g_bf16 && m && v && n > 0, "adam_step_g16: bad args");
  int64_t threads = (n + 3) / 4;
  adam_g16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, (const bf16*)g_bf16, m, v, n, lr, beta1, beta2, eps,
                                                                                        bias_corr1, bias_corr2, grad_scale);
  MAPDIT_LAUNCH_CHECK("adam_step_g16");
  return MAPDIT_OK;
}

// Multi-tensor EMA update ("next" row N2, src/ema.py:135-140: param.lerp_(model_param, beta) for every parameter of every
// tracked copy): one launch per EMA copy over a device-resident chunk table {dst, src, count}.
struct LerpChunk {
  float* dst;
  const float* src;
  long long n;
};
__global__ void __launch_bounds__(256) multi_lerp_kernel(const LerpChunk* __restrict__ table, float w) {
  const LerpChunk c = table[blockIdx.x];
  for (long long i = threadIdx.x; i < c.n; i += 256) {
    float a = c.dst[i], b = c.src[i];
    float d = b - a;
    c.dst[i] = (w < 0.5f) ? fmaf(w, d, a) : b - d * (1.0f - w);  // torch.lerp's two-branch form
  }
}
extern "C" int mapdit_multi_lerp(const void* chunk_table, int n_chunks, float weight, void* stream) {
  MAPDIT_REQUIRE(chunk_table && n_chunks > 0, "multi_lerp: bad args");
  multi_lerp_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>((const LerpChunk*)chunk_table, weight);
  MAPDIT_LAUNCH_CHECK("multi_lerp");
  return MAPDIT_OK;
}
