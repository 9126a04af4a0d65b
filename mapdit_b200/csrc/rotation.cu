// Rotation modulation (README.md:1,3 of the reference; no reference code exists — SURVEY.md §A.8, parity UNPINNED,
// checked against the self-referential oracle only):
//   h[2i]   = (x[2i] cos t_i - x[2i+1] sin t_i) * scale[2i]
//   h[2i+1] = (x[2i] sin t_i + x[2i+1] cos t_i) * scale[2i+1],   t_i = rot[n, i] * gain   (scale optional)
// A pure per-channel-pair rotation is exactly magnitude preserving.  Vectorised: each lane owns 8 channels (4 pairs).
#include "common.cuh"

template <typename T>
__global__ void __launch_bounds__(256) rotmod_fwd_kernel(const T* __restrict__ x, T* __restrict__ h, const float* __restrict__ rot,
                                                         const float* __restrict__ scale, const float* __restrict__ gain, int64_t ldmod,
                                                         int64_t m_rows, int d, int tokens) {
  const float g = *gain;
  const int per_row = d >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m_rows * per_row; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / per_row;
    const int col = (int)(i - row * per_row) * 8;
    const int64_t n = row / tokens;
    float v[8], o[8];
    load8(x + row * d + col, v);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float sn, cs;
      sincosf(rot[n * ldmod + (col >> 1) + p] * g, &sn, &cs);
      o[2 * p] = v[2 * p] * cs - v[2 * p + 1] * sn;
      o[2 * p + 1] = v[2 * p] * sn + v[2 * p + 1] * cs;
    }
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] *= scale[n * ldmod + col + j];
    }
    store8(h + row * d + col, o);
  }
}

extern "C" int mapdit_rotmod_fwd(const void* x, void* h, const float* rot, const float* scale, const float* gain, int64_t ldmod, int m,
                                 int d, int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && h && rot && gain && m > 0 && d % 8 == 0 && tokens > 0, "rotmod_fwd: bad args");
  int64_t work = (int64_t)m * (d >> 3);
  int grid = (int)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  if (dtype == MAPDIT_F32)
    rotmod_fwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)h, rot, scale, gain, ldmod, m, d, tokens);
  else
    rotmod_fwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)h, rot, scale, gain, ldmod, m, d, tokens);
  MAPDIT_LAUNCH_CHECK("rotmod_fwd");
  return MAPDIT_OK;
}

// (cos, sin) table for the GEMM epilogue MAPDIT_EPI_RESID_ROT: the angle of a channel pair depends on the sample only, so it is
// evaluated once per (sample, pair) here instead of once per token in the epilogue.  blockIdx.y selects the branch triple.
__global__ void __launch_bounds__(256) rot_table_kernel(const float* __restrict__ rot, const float* __restrict__ gain, float* __restrict__ cs,
                                                        const float* __restrict__ rot2, const float* __restrict__ gain2,
                                                        float* __restrict__ cs2, int64_t ldmod, int64_t ldcs, int n_samples, int half_d) {
  const float* r = blockIdx.y ? rot2 : rot;
  float* o = blockIdx.y ? cs2 : cs;
  const float g = blockIdx.y ? *gain2 : *gain;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_samples * half_d) return;
  const int n = i / half_d, p = i - n * half_d;
  float sn, c;
  sincosf(r[n * ldmod + p] * g, &sn, &c);
  *reinterpret_cast<float2*>(o + n * ldcs + 2 * p) = make_float2(c, sn);
}

extern "C" int mapdit_rot_table(const float* rot, const float* gain, float* cs, const float* rot2, const float* gain2, float* cs2,
                                int64_t ldmod, int64_t ldcs, int n_samples, int d, void* stream) {
  MAPDIT_REQUIRE(rot && gain && cs && n_samples > 0 && d > 0 && d % 2 == 0 && ldcs % 2 == 0 && ((uintptr_t)cs & 7) == 0, "rot_table: bad args");
  MAPDIT_REQUIRE(!rot2 || (gain2 && cs2 && ((uintptr_t)cs2 & 7) == 0), "rot_table: second branch needs rot2/gain2/cs2");
  dim3 grid((n_samples * (d / 2) + 255) / 256, rot2 ? 2 : 1);
  rot_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rot, gain, cs, rot2, gain2, cs2, ldmod, ldcs, n_samples, d / 2);
  MAPDIT_LAUNCH_CHECK("rot_table");
  return MAPDIT_OK;
}

// backward: R (+)= R_theta^T (dh * scale);  dscale = sum_t dh * rot(x);  drot[n,i] = gain * sum_t dtheta;  dgain = sum dtheta * rot
// CTA = (sample, 128-column chunk), 8 warps stride over tokens, 4 columns (2 pairs) per lane (same scheme as modulate_bwd);
// dgain partials per CTA.  FUSE: the backward of the residual that precedes this modulation in the block schedule runs on the
// freshly updated R in the same pass (see modulate_bwd_kernel in backward.cu): R'' = ca_r R', dy = cb_r gate R', dgate = sum_t cb_r y R'.
__device__ __forceinline__ void ld4(const float* p, float (&f)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
}
__device__ __forceinline__ void ld4(const bf16* p, float (&f)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void st4(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
__device__ __forceinline__ void st4(bf16* p, const float (&f)[4]) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[0], f[1]);
  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename T, bool FUSE>
__global__ void __launch_bounds__(256, FUSE ? 2 : 3) rotmod_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ x, T* R,
                                                         const float* __restrict__ rot, const float* __restrict__ scale,
                                                         const float* __restrict__ gain, float* __restrict__ drot,
                                                         float* __restrict__ dscale, float* __restrict__ dg_partial, int64_t ldmod, int d,
                                                         int tokens, int accumulate, const T* __restrict__ y, T* __restrict__ dy,
                                                         const float* __restrict__ gate, float* __restrict__ dgate, int var) {
  constexpr int CW = 128;
  __shared__ float red_sc[8][CW];
  __shared__ float red_th[8][CW / 2];
  __shared__ float red_gt[FUSE ? 8 : 1][CW];
  __shared__ float red[32];
  const bool plain_res = var & MAPDIT_VAR_PLAIN_RESID;
  const float ca_r = plain_res ? 1.0f : (1.0f - MP_RES_T) / MP_RES_DEN, cb_r = plain_res ? 1.0f : MP_RES_T / MP_RES_DEN;
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * CW + lane * 4;
  const bool ok = col < d;
  const float g = *gain;
  float sn[2], cs[2], sc[4], a_sc[4], a_th[2], gt[4], a_gt[4];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float th = ok ? rot[n * ldmod + (col >> 1) + p] * g : 0.f;
    sincosf(th, &sn[p], &cs[p]);
    a_th[p] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = (ok && scale) ? scale[n * ldmod + col + j] : 1.0f;
    a_sc[j] = 0.f;
    gt[j] = (FUSE && ok) ? cb_r * gate[n * ldmod + col + j] : 0.f;
    a_gt[j] = 0.f;
  }
  if (ok) {
    constexpr int U = FUSE ? 4 : 2;  // two tokens per trip: every load of both tokens is issued before the first dependent instruction
    for (int t0 = warp; t0 < tokens; t0 += 8 * U) {
      float gh[U][4], xv[U][4], r[U][4], yv[U][4];
      size_t off[U];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + 8 * u;
        live[u] = t < tokens;
        off[u] = ((size_t)n * tokens + (live[u] ? t : t0)) * d + col;
        ld4(dh + off[u], gh[u]);
        ld4(x + off[u], xv[u]);
        if (R && accumulate) ld4(R + off[u], r[u]);
        if (FUSE) ld4(y + off[u], yv[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const float x0 = xv[u][2 * p], x1 = xv[u][2 * p + 1];
          const float r0 = x0 * cs[p] - x1 * sn[p], r1 = x0 * sn[p] + x1 * cs[p];  // rotated x
          a_sc[2 * p] = fmaf(gh[u][2 * p], r0, a_sc[2 * p]);
          a_sc[2 * p + 1] = fmaf(gh[u][2 * p + 1], r1, a_sc[2 * p + 1]);
          const float g0 = gh[u][2 * p] * sc[2 * p], g1 = gh[u][2 * p + 1] * sc[2 * p + 1];  // d/d(rotated x)
          a_th[p] += g0 * (-r1) + g1 * r0;                                                       // d rotated / d theta = (-r1, r0)
          const float dx0 = g0 * cs[p] + g1 * sn[p], dx1 = -g0 * sn[p] + g1 * cs[p];
          r[u][2 * p] = ((R && accumulate) ? r[u][2 * p] : 0.f) + dx0;
          r[u][2 * p + 1] = ((R && accumulate) ? r[u][2 * p + 1] : 0.f) + dx1;
        }
        if (FUSE) {
          float o1[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o1[j] = gt[j] * r[u][j];
            a_gt[j] = fmaf(cb_r * yv[u][j], r[u][j], a_gt[j]);
            r[u][j] *= ca_r;
          }
          st4(dy + off[u], o1);
        }
        if (R) st4(R + off[u], r[u]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red_sc[warp][lane * 4 + j] = a_sc[j];
    if (FUSE) red_gt[warp][lane * 4 + j] = a_gt[j];
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) red_th[warp][lane * 2 + p] = a_th[p];
  __syncthreads();
  float part = 0.f;
  if (threadIdx.x < CW) {
    const int c = blockIdx.x * CW + threadIdx.x;
    if (c < d) {
      if (dscale) {
        float s1 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s1 += red_sc[w][threadIdx.x];
        dscale[n * ldmod + c] = s1;
      }
      if (FUSE) {
        float s3 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s3 += red_gt[w][threadIdx.x];
        dgate[n * ldmod + c] = s3;
      }
    }
  } else if (threadIdx.x < CW + CW / 2) {
    const int q = threadIdx.x - CW;
    const int pi = blockIdx.x * (CW / 2) + q;
    if (pi < (d >> 1)) {
      float s2 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s2 += red_th[w][q];
      drot[n * ldmod + pi] = s2 * g;
      part = s2 * rot[n * ldmod + pi];
    }
  }
  float tot = block_sum(part, red);
  if (threadIdx.x == 0) dg_partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

extern "C" int mapdit_rotmod_bwd(const void* dh, const void* x, void* R, const float* rot, const float* scale, const float* gain,
                                 float* drot, float* dscale, float* dg_partial, int64_t ldmod, int n_samples, int d, int tokens,
                                 int accumulate, int dtype, void* stream) {
  MAPDIT_REQUIRE(dh && x && rot && gain && drot && dg_partial && n_samples > 0 && d % 8 == 0, "rotmod_bwd: bad args");
  dim3 grid((d + 127) / 128, n_samples);
  if (dtype == MAPDIT_F32)
    rotmod_bwd_kernel<float, false><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dh, (const float*)x, (float*)R, rot, scale, gain, drot, dscale, dg_partial, ldmod, d, tokens, accumulate, nullptr, nullptr, nullptr, nullptr, 0);
  else
    rotmod_bwd_kernel<bf16, false><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dh, (const bf16*)x, (bf16*)R, rot, scale, gain, drot, dscale, dg_partial, ldmod, d, tokens, accumulate, nullptr, nullptr, nullptr, nullptr, 0);
  MAPDIT_LAUNCH_CHECK("rotmod_bwd");
  return MAPDIT_OK;
}
extern "C" int mapdit_rotmod_resid_bwd(const void* dh, const void* x, void* R, const float* rot, const float* scale, const float* gain,
                                       float* drot, float* dscale, float* dg_partial, const void* y, void* dy, const float* gate,
                                       float* dgate, int64_t ldmod, int n_samples, int d, int tokens, int accumulate, int dtype,
                                       void* stream) {
  MAPDIT_REQUIRE(dh && x && R && rot && gain && drot && dg_partial && y && dy && gate && dgate && n_samples > 0 && d % 8 == 0,
                 "rotmod_resid_bwd: bad args");
  dim3 grid((d + 127) / 128, n_samples);
  const int var = mapdit_variant();
  if (dtype == MAPDIT_F32)
    rotmod_bwd_kernel<float, true><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dh, (const float*)x, (float*)R, rot, scale, gain, drot, dscale, dg_partial, ldmod, d, tokens, accumulate, (const float*)y, (float*)dy, gate, dgate, var);
  else
    rotmod_bwd_kernel<bf16, true><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dh, (const bf16*)x, (bf16*)R, rot, scale, gain, drot, dscale, dg_partial, ldmod, d, tokens, accumulate, (const bf16*)y, (bf16*)dy, gate, dgate, var);
  MAPDIT_LAUNCH_CHECK("rotmod_resid_bwd");
  return MAPDIT_OK;
}
extern "C" int mapdit_rotmod_bwd_partials(int n_samples, int d) { return ((d + 127) / 128) * n_samples; }
