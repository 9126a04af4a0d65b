// Shared helpers for the libmapdit kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mapdit.h"

void mapdit_set_error(const char* fmt, ...);
void mapdit_count_launch(int n = 1);
int mapdit_variant();  // host: this thread's MAPDIT_VAR_* word (mapdit_set_variant), passed to kernels as a launch argument

#define MAPDIT_REQUIRE(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      mapdit_set_error(__VA_ARGS__);     \
      return MAPDIT_ERR_ARG;             \
    }                                    \
  } while (0)

#define MAPDIT_LAUNCH_CHECK(name)                                                  \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      mapdit_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return MAPDIT_ERR_CUDA;                                                      \
    }                                                                              \
    mapdit_count_launch();                                                         \
  } while (0)

typedef __nv_bfloat16 bf16;

// mp_sum(a, b, 0.3) denominators etc. as the fp32 values the reference ends up dividing by
// (python double -> fp32 scalar): sqrt(0.7^2+0.3^2), sqrt(0.5).
#define MP_RES_T 0.3f
#define MP_RES_DEN 0.7615773105863908f
#define MP_HALF_DEN 0.7071067811865476f
#define MP_SILU_DIV 0.596f

template <typename T> __device__ __forceinline__ float ld_act(const T* p);
template <> __device__ __forceinline__ float ld_act<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_act<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_act(T* p, float v);
template <> __device__ __forceinline__ void st_act<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_act<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// torch.lerp(a, b, w): w < 0.5 ? a + w (b-a) : b - (b-a)(1-w)   (ATen/native/Lerp.h)
__device__ __forceinline__ float lerp_t(float a, float b, float w) {
  float d = b - a;
  return (w < 0.5f) ? __fmaf_rn(w, d, a) : b - d * (1.0f - w);
}
// 1/sqrt((1-g)^2+g^2) evaluated in double like the reference's python float (src/utils.py:16)
__device__ __forceinline__ float mod_den(float g) {
  double gd = (double)g;
  return (float)sqrt((1.0 - gd) * (1.0 - gd) + gd * gd);
}
__device__ __forceinline__ float modulate_f(float x, float shift, float scale, float g, float den) {
  return lerp_t(x * scale, shift, g) / den;
}
__device__ __forceinline__ float resid_f(float x, float gate, float y) { return lerp_t(x, gate * y, MP_RES_T) / MP_RES_DEN; }
__device__ __forceinline__ float mp_silu_f(float x) { return (x / (1.0f + expf(-x))) / MP_SILU_DIV; }
// README "--use-*" switches turned off (UNPINNED): plain residual / plain SiLU
__device__ __forceinline__ float resid_v(float x, float gate, float y, int var) {
  return (var & MAPDIT_VAR_PLAIN_RESID) ? __fmaf_rn(gate, y, x) : resid_f(x, gate, y);
}
__device__ __forceinline__ float mp_silu_v(float x, int var) { return (var & MAPDIT_VAR_PLAIN_SILU) ? x / (1.0f + expf(-x)) : mp_silu_f(x); }
__device__ __forceinline__ float silu_div_v(int var) { return (var & MAPDIT_VAR_PLAIN_SILU) ? 1.0f : MP_SILU_DIV; }

// 8 consecutive activations <-> registers (16-byte accesses for bf16, 2 x 16 bytes for fp32); p must be 16-byte aligned
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 t = __bfloat1622float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum; every thread gets the result. `red` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}
