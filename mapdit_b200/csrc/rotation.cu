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

// backward: R (+)= R_theta^T (dh * scale);  dscale = sum_t dh * rot(x);  drot[n,i] = gain * sum_t dtheta;  dgain = sum dtheta * rot
// CTA = (sample, 256-column chunk), 8 warps stride over tokens (same scheme as modulate_bwd); dgain partials per CTA.
template <typename T>
__global__ void __launch_bounds__(256) rotmod_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ x, T* R,
                                                         const float* __restrict__ rot, const float* __restrict__ scale,
                                                         const float* __restrict__ gain, float* __restrict__ drot,
                                                         float* __restrict__ dscale, float* __restrict__ dg_partial, int64_t ldmod, int d,
                                                         int tokens, int accumulate) {
  __shared__ float red_sc[8][256];
  __shared__ float red_th[8][128];
  __shared__ float red[32];
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  const bool ok = col < d;
  const float g = *gain;
  float sn[4], cs[4], sc[8], a_sc[8], a_th[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float th = ok ? rot[n * ldmod + (col >> 1) + p] * g : 0.f;
    sincosf(th, &sn[p], &cs[p]);
    a_th[p] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = (ok && scale) ? scale[n * ldmod + col + j] : 1.0f;
    a_sc[j] = 0.f;
  }
  if (ok) {
    for (int t = warp; t < tokens; t += 8) {
      const size_t off = ((size_t)n * tokens + t) * d + col;
      float gh[8], xv[8], r[8];
      load8(dh + off, gh);
      load8(x + off, xv);
      if (R && accumulate) load8(R + off, r);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float x0 = xv[2 * p], x1 = xv[2 * p + 1];
        const float r0 = x0 * cs[p] - x1 * sn[p], r1 = x0 * sn[p] + x1 * cs[p];  // rotated x
        a_sc[2 * p] = fmaf(gh[2 * p], r0, a_sc[2 * p]);
        a_sc[2 * p + 1] = fmaf(gh[2 * p + 1], r1, a_sc[2 * p + 1]);
        const float g0 = gh[2 * p] * sc[2 * p], g1 = gh[2 * p + 1] * sc[2 * p + 1];  // d/d(rotated x)
        a_th[p] += g0 * (-r1) + g1 * r0;                                               // d rotated / d theta = (-r1, r0)
        const float dx0 = g0 * cs[p] + g1 * sn[p], dx1 = -g0 * sn[p] + g1 * cs[p];
        r[2 * p] = ((R && accumulate) ? r[2 * p] : 0.f) + dx0;
        r[2 * p + 1] = ((R && accumulate) ? r[2 * p + 1] : 0.f) + dx1;
      }
      if (R) store8(R + off, r);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red_sc[warp][lane * 8 + j] = a_sc[j];
#pragma unroll
  for (int p = 0; p < 4; ++p) red_th[warp][lane * 4 + p] = a_th[p];
  __syncthreads();
  float part = 0.f;
  {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < d && dscale) {
      float s1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s1 += red_sc[w][threadIdx.x];
      dscale[n * ldmod + c] = s1;
    }
    if (threadIdx.x < 128) {
      const int pi = blockIdx.x * 128 + threadIdx.x;
      if (pi < (d >> 1)) {
        float s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s2 += red_th[w][threadIdx.x];
        drot[n * ldmod + pi] = s2 * g;
        part = s2 * rot[n * ldmod + pi];
      }
    }
  }
  float tot = block_sum(part, red);
  if (threadIdx.x == 0) dg_partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

extern "C" int mapdit_rotmod_bwd(const void* dh, const void* x, void* R, const float* rot, const float* scale, const float* gain,
                                 float* drot, float* dscale, float* dg_partial, int64_t ldmod, int n_samples, int d, int tokens,
                                 int accumulate, int dtype, void* stream) {
  MAPDIT_REQUIRE(dh && x && rot && gain && drot && dg_partial && n_samples > 0 && d % 8 == 0, "rotmod_bwd: bad args");
  dim3 grid((d + 255) / 256, n_samples);
  if (dtype == MAPDIT_F32)
    rotmod_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dh, (const float*)x, (float*)R, rot, scale, gain, drot, dscale, dg_partial, ldmod, d, tokens, accumulate);
  else
    rotmod_bwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dh, (const bf16*)x, (bf16*)R, rot, scale, gain, drot, dscale, dg_partial, ldmod, d, tokens, accumulate);
  MAPDIT_LAUNCH_CHECK("rotmod_bwd");
  return MAPDIT_OK;
}
extern "C" int mapdit_rotmod_bwd_partials(int n_samples, int d) { return ((d + 255) / 256) * n_samples; }
