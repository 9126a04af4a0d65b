// Backward kernels of the block's elementwise ops (SURVEY.md §A.3 closed forms, verified against autograd):
// mp residual, modulate (with the detached denominator of src/utils.py:16), mp_silu, q/k normalisation,
// the final-layer scale/unpatchify, and the conditioning path.  Per-sample reductions over the T tokens
// (dshift/dscale/dgate) are done by "column strip" threads: one thread owns one (sample, channel) and
// walks its T rows, so loads are coalesced across channels and the sums are deterministic.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// x_out = mp_sum(x, gate*y, 0.3):   R <- 0.7/den * R (in place) ; dy = 0.3/den * gate * R_in ; dgate = sum_t 0.3/den * y * R_in
// ------------------------------------------------------------------------------------------------
// CTA = (sample, 256-column chunk); 8 warps stride over the T rows, each lane owns 8 consecutive columns (16-byte
// accesses); per-column sums are combined across the warps in shared memory in a fixed order (deterministic).
template <typename T>
__global__ void __launch_bounds__(256) resid_bwd_kernel(T* __restrict__ R, const T* __restrict__ y, T* __restrict__ dy,
                                                        const float* __restrict__ gate, float* __restrict__ dgate, int64_t ldmod,
                                                        int d, int tokens, int var) {
  __shared__ float red[8][256];
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * 256 + lane * 8;
  const bool ok = col < d;
  const bool plain = var & MAPDIT_VAR_PLAIN_RESID;  // x + gate*y instead of mp_sum(x, gate*y, 0.3)
  const float ca = plain ? 1.0f : (1.0f - MP_RES_T) / MP_RES_DEN, cb = plain ? 1.0f : MP_RES_T / MP_RES_DEN;
  float g[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    g[j] = ok ? cb * gate[n * ldmod + col + j] : 0.f;
    acc[j] = 0.f;
  }
  if (ok) {
    for (int t = warp; t < tokens; t += 8) {
      const size_t off = ((size_t)n * tokens + t) * d + col;
      float r[8], yv[8], o1[8], o2[8];
      load8(R + off, r);
      load8(y + off, yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o1[j] = g[j] * r[j];
        acc[j] = fmaf(cb * yv[j], r[j], acc[j]);
        o2[j] = ca * r[j];
      }
      store8(dy + off, o1);
      store8(R + off, o2);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < d) {
    float sacc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) sacc += red[w][threadIdx.x];
    dgate[n * ldmod + c] = sacc;
  }
}

extern "C" int mapdit_resid_bwd(void* R, const void* y, void* dy, const float* gate, float* dgate, int64_t ldmod, int n_samples,
                                int d, int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(R && y && dy && gate && dgate && n_samples > 0 && d > 0 && tokens > 0 && d % 8 == 0, "resid_bwd: bad args");
  dim3 grid((d + 255) / 256, n_samples);
  if (dtype == MAPDIT_F32)
    resid_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)R, (const float*)y, (float*)dy, gate, dgate, ldmod, d, tokens, mapdit_variant());
  else
    resid_bwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((bf16*)R, (const bf16*)y, (bf16*)dy, gate, dgate, ldmod, d, tokens, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("resid_bwd");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// h = lerp(x*scale, shift, g)/den(g), den detached:
//   R += dh*(1-g)/den*scale ; dscale = sum_t dh*(1-g)/den*x ; dshift = sum_t dh*g/den ; dg = sum dh*(shift - x*scale)/den
// dg partials: one float per CTA (deterministic two-stage reduction, finished by mapdit_sum_partials).
// R may be null (first block: the input has no gradient); accumulate==0 overwrites R.
// ------------------------------------------------------------------------------------------------
// FUSE: the residual backward that always follows in the block schedule (mapdit_resid_bwd on the freshly updated R, with the
// branch output y and gate of the PREVIOUS residual) runs in the same pass: R'' = ca_r R', dy = cb_r gate R', dgate = sum_t cb_r y R'
// — 6 passes over [M, D] instead of 8, and R' is never rounded to bf16 in between.
// V = columns per lane.  V = 4 (CTA = 128 columns) keeps the fused kernel under 80 registers so four CTAs fit per SM: the
// first fused version (V = 8, 114 registers, two CTAs per SM) ran at 48 % of the HBM peak, below the two kernels it replaced.
template <int V> struct VecIO;
template <> struct VecIO<8> {
  template <typename T> static __device__ __forceinline__ void ld(const T* p, float (&f)[8]) { load8(p, f); }
  template <typename T> static __device__ __forceinline__ void st(T* p, const float (&f)[8]) { store8(p, f); }
};
template <> struct VecIO<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&f)[4]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  }
  static __device__ __forceinline__ void ld(const bf16* p, float (&f)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static __device__ __forceinline__ void st(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
  static __device__ __forceinline__ void st(bf16* p, const float (&f)[4]) {
    uint2 u;
    *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(f[0], f[1]);
    *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

template <typename T, bool FUSE, int V>
__global__ void __launch_bounds__(256, FUSE ? 2 : 3) modulate_bwd_kernel(const T* __restrict__ dh, const T* __restrict__ x, T* R,
                                                           const float* __restrict__ shift, const float* __restrict__ scale,
                                                           const float* __restrict__ gain, float* __restrict__ dshift,
                                                           float* __restrict__ dscale, float* __restrict__ dg_partial, int64_t ldmod,
                                                           int d, int tokens, int accumulate, const T* __restrict__ y,
                                                           T* __restrict__ dy, const float* __restrict__ gate,
                                                           float* __restrict__ dgate, int var) {
  constexpr int CW = 32 * V;  // columns per CTA
  __shared__ float red_sc[8][CW];
  __shared__ float red_sh[8][CW];
  __shared__ float red_gt[FUSE ? 8 : 1][CW];
  __shared__ float red[32];
  const bool plain_res = var & MAPDIT_VAR_PLAIN_RESID;
  const float ca_r = plain_res ? 1.0f : (1.0f - MP_RES_T) / MP_RES_DEN, cb_r = plain_res ? 1.0f : MP_RES_T / MP_RES_DEN;
  float gt[V], a_gt[V];
  const int n = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = blockIdx.x * CW + lane * V;
  const bool ok = col < d;
  const float g = *gain;
  const float den = mod_den(g);
  const float ca = (1.0f - g) / den, cb = g / den, cd = 1.0f / den;
  float sc[V], sh[V], a_sc[V], a_sh[V], a_g = 0.f;
#pragma unroll
  for (int j = 0; j < V; ++j) {
    sc[j] = ok ? scale[n * ldmod + col + j] : 0.f;
    sh[j] = ok ? shift[n * ldmod + col + j] : 0.f;
    a_sc[j] = 0.f;
    a_sh[j] = 0.f;
    if (FUSE) {
      gt[j] = ok ? cb_r * gate[n * ldmod + col + j] : 0.f;
      a_gt[j] = 0.f;
    }
  }
  if (ok) {
    // two tokens per trip (U = 2): all loads of both tokens are issued before the first dependent instruction, which doubles
    // the bytes each warp keeps in flight (8-byte loads: the single-token loop reached ~55 % of the HBM peak)
    constexpr int U = FUSE ? 4 : 2;
    for (int t0 = warp; t0 < tokens; t0 += 8 * U) {
      float gh[U][V], xv[U][V], r[U][V], yv[U][V];
      size_t off[U];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + 8 * u;
        live[u] = t < tokens;
        off[u] = ((size_t)n * tokens + (live[u] ? t : t0)) * d + col;
        VecIO<V>::ld(dh + off[u], gh[u]);
        VecIO<V>::ld(x + off[u], xv[u]);
        if (R && accumulate) VecIO<V>::ld(R + off[u], r[u]);
        if (FUSE) VecIO<V>::ld(y + off[u], yv[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
        if (R) {
#pragma unroll
          for (int j = 0; j < V; ++j) r[u][j] = (accumulate ? r[u][j] : 0.f) + ca * sc[j] * gh[u][j];
          if (FUSE) {
            float o1[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
              o1[j] = gt[j] * r[u][j];
              a_gt[j] = fmaf(cb_r * yv[u][j], r[u][j], a_gt[j]);
              r[u][j] *= ca_r;
            }
            VecIO<V>::st(dy + off[u], o1);
          }
          VecIO<V>::st(R + off[u], r[u]);
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
          a_sc[j] = fmaf(ca * xv[u][j], gh[u][j], a_sc[j]);
          a_sh[j] += gh[u][j];
          a_g = fmaf(gh[u][j], sh[j] - xv[u][j] * sc[j], a_g);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {
    red_sc[warp][lane * V + j] = a_sc[j];
    red_sh[warp][lane * V + j] = a_sh[j];
    if (FUSE) red_gt[warp][lane * V + j] = a_gt[j];
  }
  float tot = block_sum(a_g * cd, red);  // contains the __syncthreads that publishes red_sc / red_sh
  const int c = blockIdx.x * CW + threadIdx.x;
  if (threadIdx.x < CW && c < d) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      s1 += red_sc[w][threadIdx.x];
      s2 += red_sh[w][threadIdx.x];
    }
    dscale[n * ldmod + c] = s1;
    dshift[n * ldmod + c] = cb * s2;
    if (FUSE) {
      float s3 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s3 += red_gt[w][threadIdx.x];
      dgate[n * ldmod + c] = s3;
    }
  }
  if (threadIdx.x == 0) dg_partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
}

extern "C" int mapdit_modulate_bwd(const void* dh, const void* x, void* R, const float* shift, const float* scale, const float* gain,
                                   float* dshift, float* dscale, float* dg_partial, int64_t ldmod, int n_samples, int d, int tokens,
                                   int accumulate, int dtype, void* stream) {
  MAPDIT_REQUIRE(dh && x && shift && scale && gain && dshift && dscale && dg_partial && n_samples > 0 && d % 8 == 0, "modulate_bwd: bad args");
  dim3 grid((d + 127) / 128, n_samples);
  if (dtype == MAPDIT_F32)
    modulate_bwd_kernel<float, false, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dh, (const float*)x, (float*)R, shift, scale, gain, dshift, dscale, dg_partial, ldmod, d, tokens, accumulate, nullptr, nullptr, nullptr, nullptr, 0);
  else
    modulate_bwd_kernel<bf16, false, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dh, (const bf16*)x, (bf16*)R, shift, scale, gain, dshift, dscale, dg_partial, ldmod, d, tokens, accumulate, nullptr, nullptr, nullptr, nullptr, 0);
  MAPDIT_LAUNCH_CHECK("modulate_bwd");
  return MAPDIT_OK;
}
extern "C" int mapdit_modulate_resid_bwd(const void* dh, const void* x, void* R, const float* shift, const float* scale,
                                         const float* gain, float* dshift, float* dscale, float* dg_partial, const void* y, void* dy,
                                         const float* gate, float* dgate, int64_t ldmod, int n_samples, int d, int tokens,
                                         int accumulate, int dtype, void* stream) {
  MAPDIT_REQUIRE(dh && x && R && shift && scale && gain && dshift && dscale && dg_partial && y && dy && gate && dgate && n_samples > 0 &&
                     d % 8 == 0,
                 "modulate_resid_bwd: bad args");
  dim3 grid((d + 127) / 128, n_samples);
  const int var = mapdit_variant();
  if (dtype == MAPDIT_F32)
    modulate_bwd_kernel<float, true, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)dh, (const float*)x, (float*)R, shift, scale, gain, dshift, dscale, dg_partial, ldmod, d, tokens, accumulate, (const float*)y, (float*)dy, gate, dgate, var);
  else
    modulate_bwd_kernel<bf16, true, 4><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dh, (const bf16*)x, (bf16*)R, shift, scale, gain, dshift, dscale, dg_partial, ldmod, d, tokens, accumulate, (const bf16*)y, (bf16*)dy, gate, dgate, var);
  MAPDIT_LAUNCH_CHECK("modulate_resid_bwd");
  return MAPDIT_OK;
}
extern "C" int mapdit_modulate_bwd_partials(int n_samples, int d) { return ((d + 127) / 128) * n_samples; }

// out[0] (+)= sum(partials[0..n))   -- single CTA, fixed order
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ p, int n, float* __restrict__ out, int accumulate) {
  __shared__ float red[32];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) a += p[i];
  a = block_sum(a, red);
  if (threadIdx.x == 0) *out = accumulate ? *out + a : a;
}
extern "C" int mapdit_sum_partials(const float* partials, int n, float* out, int accumulate, void* stream) {
  MAPDIT_REQUIRE(partials && out && n > 0, "sum_partials: bad args");
  sum_partials_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, n, out, accumulate);
  MAPDIT_LAUNCH_CHECK("sum_partials");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// u = silu(z)/0.596 :  dz = du * sigma(z) (1 + z (1 - sigma(z))) / 0.596
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void mp_silu_bwd_kernel(const T* du, const T* __restrict__ z, T* dz, int64_t n, int var) {
  const float div = silu_div_v(var);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float zv = ld_act(z + i);
    float s = 1.0f / (1.0f + expf(-zv));
    st_act(dz + i, ld_act(du + i) * (s * (1.0f + zv * (1.0f - s))) / div);
  }
}
extern "C" int mapdit_mp_silu_bwd(const void* du, const void* z, void* dz, int64_t n, int dtype, void* stream) {
  MAPDIT_REQUIRE(du && z && dz && n > 0, "mp_silu_bwd: bad args");
  int64_t b = (n + 255) / 256;
  int grid = (int)(b < 148 * 16 ? b : 148 * 16);
  if (dtype == MAPDIT_F32) mp_silu_bwd_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)du, (const float*)z, (float*)dz, n, mapdit_variant());
  else mp_silu_bwd_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)du, (const bf16*)z, (bf16*)dz, n, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("mp_silu_bwd");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// y = g v/(r+eps), g = sqrt(hd):  dv = (g/(r+eps)) (G - y (y.G) (r+eps)/(g^2 r)),  with sc = g/(r+eps) saved by the forward.
// In place on the q|k thirds of dqkv[M, 3D]; one warp per (row, head).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void qk_norm_bwd_kernel(T* __restrict__ dqkv, const T* __restrict__ qkv, const float* __restrict__ sc, int64_t total,
                                   int heads2, int d, int hd, float eps) {
  int lane = threadIdx.x & 31;
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= total) return;
  int64_t row = wid / heads2;
  int hh = (int)(wid - row * heads2);
  const size_t off = (size_t)row * 3 * d + (size_t)hh * hd;
  float yv[4], gv[4], dot = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    yv[j] = (c < hd) ? ld_act(qkv + off + c) : 0.f;
    gv[j] = (c < hd) ? ld_act(dqkv + off + c) : 0.f;
    dot = fmaf(yv[j], gv[j], dot);
  }
  dot = warp_sum(dot);
  const float s = sc[wid];
  const float gsq = (float)hd;
  const float rpe = sqrtf(gsq) / s;  // r + eps
  const float r = fmaxf(rpe - eps, 1e-30f);
  const float coef = dot * rpe / (gsq * r);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    if (c < hd) st_act(dqkv + off + c, s * (gv[j] - yv[j] * coef));
  }
}
// head_dim 64: 8 lanes x 8 elements per head, four heads per warp, 16-byte accesses
template <typename T>
__global__ void qk_norm_bwd_vec64_kernel(T* dqkv, const T* __restrict__ qkv, const float* __restrict__ sc, int64_t total, int heads2,
                                         int d, float eps) {
  const int sub = threadIdx.x & 7;
  int64_t hid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool ok = hid < total;
  if (!ok) hid = total - 1;
  int64_t row = hid / heads2;
  int hh = (int)(hid - row * heads2);
  const size_t off = (size_t)row * 3 * d + (size_t)hh * 64 + sub * 8;
  float yv[8], gv[8], dot = 0.f;
  load8(qkv + off, yv);
  load8(dqkv + off, gv);
#pragma unroll
  for (int j = 0; j < 8; ++j) dot = fmaf(yv[j], gv[j], dot);
  dot += __shfl_xor_sync(0xffffffffu, dot, 1);
  dot += __shfl_xor_sync(0xffffffffu, dot, 2);
  dot += __shfl_xor_sync(0xffffffffu, dot, 4);
  const float s = sc[hid];
  const float rpe = 8.0f / s;
  const float r = fmaxf(rpe - eps, 1e-30f);
  const float coef = dot * rpe / (64.0f * r);
#pragma unroll
  for (int j = 0; j < 8; ++j) gv[j] = s * (gv[j] - yv[j] * coef);
  if (ok) store8(dqkv + off, gv);
}

extern "C" int mapdit_qk_norm_bwd(void* dqkv, const void* qkv, const float* sc, int m, int d, int head_dim, float eps, int dtype,
                                  void* stream) {
  MAPDIT_REQUIRE(dqkv && qkv && sc && m > 0 && d % head_dim == 0 && head_dim <= 128, "qk_norm_bwd: bad args");
  int heads2 = 2 * (d / head_dim);
  int64_t total = (int64_t)m * heads2;
  if (head_dim == 64) {
    unsigned vb = (unsigned)((total * 8 + 255) / 256);
    if (dtype == MAPDIT_F32) qk_norm_bwd_vec64_kernel<float><<<vb, 256, 0, (cudaStream_t)stream>>>((float*)dqkv, (const float*)qkv, sc, total, heads2, d, eps);
    else qk_norm_bwd_vec64_kernel<bf16><<<vb, 256, 0, (cudaStream_t)stream>>>((bf16*)dqkv, (const bf16*)qkv, sc, total, heads2, d, eps);
    MAPDIT_LAUNCH_CHECK("qk_norm_bwd");
    return MAPDIT_OK;
  }
  unsigned blocks = (unsigned)((total * 32 + 255) / 256);
  if (dtype == MAPDIT_F32) qk_norm_bwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)dqkv, (const float*)qkv, sc, total, heads2, d, head_dim, eps);
  else qk_norm_bwd_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)dqkv, (const bf16*)qkv, sc, total, heads2, d, head_dim, eps);
  MAPDIT_LAUNCH_CHECK("qk_norm_bwd");
  return MAPDIT_OK;
}

// forward companion: normalise q|k in place and record sc = sqrt(hd)/(||v||+eps) per (row, head)
template <typename T>
__global__ void qk_normalize_save_kernel(T* __restrict__ qkv, float* __restrict__ sc, int64_t total, int heads2, int d, int hd, float eps) {
  int lane = threadIdx.x & 31;
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= total) return;
  int64_t row = wid / heads2;
  int hh = (int)(wid - row * heads2);
  T* p = qkv + (size_t)row * 3 * d + (size_t)hh * hd;
  float v[4], ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    v[j] = (c < hd) ? ld_act(p + c) : 0.f;
    ss = fmaf(v[j], v[j], ss);
  }
  float nrm = sqrtf(warp_sum(ss));
  float sq = sqrtf((float)hd), den = nrm + eps;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    if (c < hd) st_act(p + c, (v[j] * sq) / den);
  }
  if (lane == 0) sc[wid] = sq / den;
}
// bf16, head_dim % 8 == 0 (72 of DiT-XL: the q/k normalisation is not fused into the qkv GEMM epilogue there): sixteen lanes per
// (row, head), lane l < head_dim / 8 moves one 16-byte chunk — the scalar kernel above issues 2-byte loads and stores and ran at
// 1.3 TB/s (6.5 ms of a 162 ms DiT-XL/2 @ 64x64 training step).  sc may be null (eval forward).
__global__ void __launch_bounds__(256) qk_normalize_vec_kernel(bf16* __restrict__ qkv, float* __restrict__ sc, int64_t total, int heads2, int d,
                                                               int hd, float eps) {
  const int lane = threadIdx.x & 31, sub = lane & 15;
  const int64_t gid = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 2 + (lane >> 4);
  const bool live = gid < total;
  const int64_t row = live ? gid / heads2 : 0;
  const int hh = live ? (int)(gid - row * heads2) : 0;
  uint4* p = reinterpret_cast<uint4*>(qkv + (size_t)row * 3 * d + (size_t)hh * hd);
  const bool mine = live && sub * 8 < hd;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (mine) u = p[sub];
  float f[8];
  {
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = __bfloat1622float2(h2[e]);
      f[2 * e] = t.x;
      f[2 * e + 1] = t.y;
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);  // within the 16-lane group
  const float sq = sqrtf((float)hd), den = sqrtf(ss) + eps;
  if (mine) {
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn((f[2 * e] * sq) / den, (f[2 * e + 1] * sq) / den);
    p[sub] = u;
  }
  if (live && sub == 0 && sc) sc[gid] = sq / den;
}
static bool qk_vec_ok(const void* qkv, int d, int head_dim, int dtype) {
  return dtype == MAPDIT_BF16 && head_dim % 8 == 0 && head_dim <= 128 && d % 8 == 0 && ((uintptr_t)qkv & 15) == 0;
}
int mapdit_qk_normalize_vec(void* qkv, float* sc, int m, int d, int head_dim, float eps, void* stream) {
  const int heads2 = 2 * (d / head_dim);
  const int64_t total = (int64_t)m * heads2;
  const unsigned blocks = (unsigned)((total + 15) / 16);  // 8 warps x 2 (row, head) groups per CTA
  qk_normalize_vec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)qkv, sc, total, heads2, d, head_dim, eps);
  return MAPDIT_OK;
}
extern "C" int mapdit_qk_normalize_save(void* qkv, float* sc, int m, int d, int head_dim, float eps, int dtype, void* stream) {
  MAPDIT_REQUIRE(qkv && sc && m > 0 && d % head_dim == 0 && head_dim <= 128, "qk_normalize_save: bad args");
  if (qk_vec_ok(qkv, d, head_dim, dtype)) {
    mapdit_qk_normalize_vec(qkv, sc, m, d, head_dim, eps, stream);
    MAPDIT_LAUNCH_CHECK("qk_normalize_save(vec)");
    return MAPDIT_OK;
  }
  int heads2 = 2 * (d / head_dim);
  int64_t total = (int64_t)m * heads2;
  unsigned blocks = (unsigned)((total * 32 + 255) / 256);
  if (dtype == MAPDIT_F32) qk_normalize_save_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)qkv, sc, total, heads2, d, head_dim, eps);
  else qk_normalize_save_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)qkv, sc, total, heads2, d, head_dim, eps);
  MAPDIT_LAUNCH_CHECK("qk_normalize_save");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// final layer: out[n, part*C+c, y, x] = lin[tok, j] * s_part[n]
//   dlin[tok, j] = dout * s_part[n] ;  ds_part[n] = sum dout * lin      (one CTA per sample)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) final_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ lin,
                                                        const float* __restrict__ s_mu, const float* __restrict__ s_sg,
                                                        T* __restrict__ dlin, float* __restrict__ ds_mu, float* __restrict__ ds_sg, int C,
                                                        int S, int p) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const int g = S / p, ppc = p * p * C, per = 2 * C * S * S;
  float a_mu = 0.f, a_sg = 0.f;
  const float smu = s_mu[n], ssg = s_sg[n];
  for (int i = threadIdx.x; i < per; i += blockDim.x) {
    int xx = i % S, r = i / S;
    int yy = r % S;
    int ch = r / S;
    int part = ch / C, c = ch - part * C;
    int hh = yy / p, p1 = yy - hh * p, ww = xx / p, p2 = xx - ww * p;
    size_t li = ((size_t)n * (g * g) + hh * g + ww) * (2 * ppc) + part * ppc + (p1 * p + p2) * C + c;
    float go = dout[(size_t)n * per + i];
    float lv = ld_act(lin + li);
    st_act(dlin + li, go * (part ? ssg : smu));
    if (part) a_sg = fmaf(go, lv, a_sg);
    else a_mu = fmaf(go, lv, a_mu);
  }
  a_mu = block_sum(a_mu, red);
  a_sg = block_sum(a_sg, red);
  if (threadIdx.x == 0) {
    ds_mu[n] = a_mu;
    ds_sg[n] = a_sg;
  }
}
extern "C" int mapdit_final_bwd(const float* dout, const void* lin, const float* s_mu, const float* s_sigma, void* dlin, float* ds_mu,
                                float* ds_sigma, int n_samples, int channels, int input_size, int patch, int dtype, void* stream) {
  MAPDIT_REQUIRE(dout && lin && s_mu && s_sigma && dlin && ds_mu && ds_sigma && n_samples > 0, "final_bwd: bad args");
  if (dtype == MAPDIT_F32)
    final_bwd_kernel<float><<<n_samples, 256, 0, (cudaStream_t)stream>>>(dout, (const float*)lin, s_mu, s_sigma, (float*)dlin, ds_mu, ds_sigma, channels, input_size, patch);
  else
    final_bwd_kernel<bf16><<<n_samples, 256, 0, (cudaStream_t)stream>>>(dout, (const bf16*)lin, s_mu, s_sigma, (bf16*)dlin, ds_mu, ds_sigma, channels, input_size, patch);
  MAPDIT_LAUNCH_CHECK("final_bwd");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// MPScale: s = sigmoid(a), a = sum_j l[n,j] ref[j] / sqrt(A), l = c W_eff^T  (src/blocks/final_layer.py:20-22)
//   da = ds s (1-s);  dl[n,j] = da ref[j]/sqrt(A);  dref[j] = sum_n da l[n,j]/sqrt(A)   (single CTA)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mp_scale_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ s,
                                                           const float* __restrict__ l, const float* __restrict__ ref,
                                                           float* __restrict__ dl, float* __restrict__ dref, int n, int adim, int accumulate) {
  __shared__ float red[32];
  const float isq = rsqrtf((float)adim);
  for (int j = 0; j < adim; ++j) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float da = ds[i] * s[i] * (1.0f - s[i]);
      dl[(size_t)i * adim + j] = da * ref[j] * isq;
      acc = fmaf(da * isq, l[(size_t)i * adim + j], acc);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) dref[j] = accumulate ? dref[j] + acc : acc;
  }
}
extern "C" int mapdit_mp_scale_bwd(const float* ds, const float* s, const float* l, const float* ref, float* dl, float* dref, int n,
                                   int adim, int accumulate, void* stream) {
  MAPDIT_REQUIRE(ds && s && l && ref && dl && dref && n > 0 && adim > 0, "mp_scale_bwd: bad args");
  mp_scale_bwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ds, s, l, ref, dl, dref, n, adim, accumulate);
  MAPDIT_LAUNCH_CHECK("mp_scale_bwd");
  return MAPDIT_OK;
}
// forward companion: s[n] = sigmoid(sum_j l[n,j] ref[j]/sqrt(A)) from the precomputed l = c W_eff^T
__global__ void mp_scale_from_lin_kernel(const float* __restrict__ l, const float* __restrict__ ref, float* __restrict__ s, int n, int adim) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int j = 0; j < adim; ++j) a = fmaf(l[(size_t)i * adim + j], ref[j], a);
  a /= sqrtf((float)adim);
  s[i] = 1.0f / (1.0f + expf(-a));
}
extern "C" int mapdit_mp_scale_from_lin(const float* l, const float* ref, float* s, int n, int adim, void* stream) {
  MAPDIT_REQUIRE(l && ref && s && n > 0, "mp_scale_from_lin: bad args");
  mp_scale_from_lin_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(l, ref, s, n, adim);
  MAPDIT_LAUNCH_CHECK("mp_scale_from_lin");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// conditioning path
// ------------------------------------------------------------------------------------------------
// c = (a+b) * 0.5 / sqrt(.5), cs = silu(c)/0.596 :  dc_total = dc + dcs * silu'(c)/0.596 ; da = db = dc_total * 0.5/sqrt(.5)
__global__ void cond_combine_bwd_kernel(const float* __restrict__ c, const float* __restrict__ dc, const float* __restrict__ dcs,
                                        float* __restrict__ dab, int64_t n, int var) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float cv = c[i];
  float s = 1.0f / (1.0f + expf(-cv));
  float tot = (dc ? dc[i] : 0.f) + dcs[i] * (s * (1.0f + cv * (1.0f - s))) / silu_div_v(var);
  dab[i] = (var & MAPDIT_VAR_PLAIN_EMBED) ? tot : tot * (0.5f / MP_HALF_DEN);
}
extern "C" int mapdit_cond_combine_bwd(const float* c, const float* dc, const float* dcs, float* dab, int64_t n, void* stream) {
  MAPDIT_REQUIRE(c && dcs && dab && n > 0, "cond_combine_bwd: bad args");
  cond_combine_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, dc, dcs, dab, n, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("cond_combine_bwd");
  return MAPDIT_OK;
}

// label embedding: out[n] = normalize(E[id]) -> dE[id] += sqrt(d)/(r+eps) (G - v (v.G)/(r (r+eps)))  (rows may repeat: atomics)
__global__ void __launch_bounds__(256) embed_rows_bwd_kernel(const int64_t* __restrict__ idx, const uint8_t* __restrict__ drop,
                                                             int64_t null_idx, const float* __restrict__ table,
                                                             const float* __restrict__ g, float* __restrict__ dtable, int d, float eps,
                                                             int var) {
  __shared__ float red[32];
  int n = blockIdx.x;
  int64_t id = idx[n];
  if (drop && drop[n]) id = null_idx;
  const float* row = table + id * d;
  if (var & MAPDIT_VAR_PLAIN_EMBED) {
    for (int i = threadIdx.x; i < d; i += blockDim.x) atomicAdd(dtable + id * d + i, g[(size_t)n * d + i]);
    return;
  }
  float ss = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    ss = fmaf(row[i], row[i], ss);
    dot = fmaf(row[i], g[(size_t)n * d + i], dot);
  }
  float r = sqrtf(block_sum(ss, red));
  dot = block_sum(dot, red);
  float k = sqrtf((float)d) / (r + eps);
  float coef = dot / (fmaxf(r, 1e-30f) * (r + eps));
  for (int i = threadIdx.x; i < d; i += blockDim.x) atomicAdd(dtable + id * d + i, k * (g[(size_t)n * d + i] - row[i] * coef));
}
extern "C" int mapdit_embed_rows_bwd(const int64_t* idx, const uint8_t* drop_mask, int64_t null_idx, const float* table,
                                     const float* g, float* dtable, int n, int d, float eps, void* stream) {
  MAPDIT_REQUIRE(idx && table && g && dtable && n > 0 && d > 0, "embed_rows_bwd: bad args");
  embed_rows_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(idx, drop_mask, null_idx, table, g, dtable, d, eps, mapdit_variant());
  MAPDIT_LAUNCH_CHECK("embed_rows_bwd");
  return MAPDIT_OK;
}

// patchify(x)|1 as an explicit [M, K+1] fp32 matrix (A operand of the x_embedder weight gradient)
__global__ void patchify_kernel(const float* __restrict__ x, float* __restrict__ P, int64_t total, int C, int S, int p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int g = S / p, K = p * p * C, K1 = K + 1;
  int f = (int)(i % K1);
  int64_t tok = i / K1;
  float v = 1.0f;
  if (f < K) {
    int64_t n = tok / (g * g);
    int tt = (int)(tok - n * (g * g));
    int hh = tt / g, ww = tt - hh * g;
    int c = f % C, pp = f / C, p1 = pp / p, p2 = pp - p1 * p;
    v = x[((n * C + c) * S + (hh * p + p1)) * (int64_t)S + (ww * p + p2)];
  }
  P[i] = v;
}
extern "C" int mapdit_patchify(const float* x, float* P, int n_samples, int channels, int input_size, int patch, void* stream) {
  MAPDIT_REQUIRE(x && P && n_samples > 0, "patchify: bad args");
  int g = input_size / patch;
  int64_t total = (int64_t)n_samples * g * g * (patch * patch * channels + 1);
  patchify_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, P, total, channels, input_size, patch);
  MAPDIT_LAUNCH_CHECK("patchify");
  return MAPDIT_OK;
}

// x_embedder weight gradient: dW[dch, k] = scale * sum_tok R[tok, dch] * P[tok, k]  (K+1 <= 32 columns per pass).
// CTA = 256 tokens x 256 channels: patches staged in smem, each thread owns one channel and keeps K+1 accumulators,
// partial sums over the token chunks are combined with fp32 atomics into a zeroed dW.
// KP = columns per pass, a multiple of 4 (20 covers the 17 columns of the patch-2 models in one pass): a token's patch row is
// read with KP/4 broadcast 16-byte shared-memory loads (the first version issued 32 scalar loads per token and was bound by
// shared-memory wavefronts: 0.35 ms for 1.6 GFMA).
template <typename T, int KP>
__global__ void __launch_bounds__(256) patch_embed_wgrad_kernel(const T* __restrict__ R, const float* __restrict__ x,
                                                                float* __restrict__ dW, int64_t m_total, int C, int S, int p, int d,
                                                                int k0, float scale) {
  __shared__ __align__(16) float sp[64][KP];
  const int g = S / p, T_ = g * g, K = p * p * C, K1 = K + 1;
  const int kn = min(KP, K1 - k0);
  const int col = blockIdx.y * 256 + threadIdx.x;
  float acc[KP];
#pragma unroll
  for (int j = 0; j < KP; ++j) acc[j] = 0.f;
  const int64_t tok_begin = (int64_t)blockIdx.x * 256;
  for (int64_t t0 = tok_begin; t0 < min(m_total, tok_begin + 256); t0 += 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * KP; i += 256) {
      const int tl = i / KP, kk = i - tl * KP;
      const int64_t tok = t0 + tl;
      float v = 0.f;
      if (tok < m_total && kk < kn) {
        const int f = k0 + kk;
        if (f == K) v = 1.0f;
        else {
          const unsigned tk = (unsigned)tok;
          const unsigned n = tk / (unsigned)T_, tt = tk - n * (unsigned)T_;
          const unsigned hh = tt / (unsigned)g, ww = tt - hh * (unsigned)g;
          const unsigned pp = (unsigned)f / (unsigned)C, c = (unsigned)f - pp * (unsigned)C, p1 = pp / (unsigned)p, p2 = pp - p1 * (unsigned)p;
          v = x[(((size_t)n * C + c) * S + (hh * p + p1)) * (size_t)S + (ww * p + p2)];
        }
      }
      sp[tl][kk] = v;
    }
    __syncthreads();
    if (col < d) {
      const int tmax = (int)min((int64_t)64, m_total - t0);
#pragma unroll 4
      for (int tl = 0; tl < tmax; ++tl) {
        const float r = ld_act(R + (t0 + tl) * d + col);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&sp[tl][4 * q]);
          acc[4 * q] = fmaf(r, v.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(r, v.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(r, v.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(r, v.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (col < d) {
#pragma unroll
    for (int j = 0; j < KP; ++j)
      if (j < kn) atomicAdd(dW + (size_t)col * K1 + k0 + j, acc[j] * scale);
  }
}
extern "C" int mapdit_patch_embed_wgrad(const void* R, const float* x, float* dW, int n_samples, int channels, int input_size,
                                        int patch, int d, float scale, int dtype, void* stream) {
  MAPDIT_REQUIRE(R && x && dW && n_samples > 0 && input_size % patch == 0, "patch_embed_wgrad: bad args");
  const int g = input_size / patch, K1 = patch * patch * channels + 1;
  const int64_t m_total = (int64_t)n_samples * g * g;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(dW, 0, (size_t)d * K1 * sizeof(float), s);
  dim3 grid((unsigned)((m_total + 255) / 256), (unsigned)((d + 255) / 256));
  MAPDIT_REQUIRE(m_total < (1LL << 31), "patch_embed_wgrad: more than 2^31 tokens");
  if (K1 <= 20) {
    if (dtype == MAPDIT_F32)
      patch_embed_wgrad_kernel<float, 20><<<grid, 256, 0, s>>>((const float*)R, x, dW, m_total, channels, input_size, patch, d, 0, scale);
    else
      patch_embed_wgrad_kernel<bf16, 20><<<grid, 256, 0, s>>>((const bf16*)R, x, dW, m_total, channels, input_size, patch, d, 0, scale);
    MAPDIT_LAUNCH_CHECK("patch_embed_wgrad");
    return MAPDIT_OK;
  }
  for (int k0 = 0; k0 < K1; k0 += 32) {
    if (dtype == MAPDIT_F32)
      patch_embed_wgrad_kernel<float, 32><<<grid, 256, 0, s>>>((const float*)R, x, dW, m_total, channels, input_size, patch, d, k0, scale);
    else
      patch_embed_wgrad_kernel<bf16, 32><<<grid, 256, 0, s>>>((const bf16*)R, x, dW, m_total, channels, input_size, patch, d, k0, scale);
    MAPDIT_LAUNCH_CHECK("patch_embed_wgrad");
  }
  return MAPDIT_OK;
}

// generic: y = a*x (+ y)  fp32, used for tiny conditioning-path scalings
__global__ void axpby_kernel(const float* x, float* y, float a, int accumulate, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = accumulate ? fmaf(a, x[i], y[i]) : a * x[i];
}
extern "C" int mapdit_axpby(const float* x, float* y, float a, int accumulate, int64_t n, void* stream) {
  MAPDIT_REQUIRE(x && y && n > 0, "axpby: bad args");
  axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, a, accumulate, n);
  MAPDIT_LAUNCH_CHECK("axpby");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// adaLN with LayerNorm (use_no_layernorm=False, UNPINNED): h = xh (1 + scale) + shift, xh = (x - mean) rstd.
//   dshift = sum_t dh ; dscale = sum_t dh xh ; dxh = dh (1 + scale) ; dx = rstd (dxh - mean_c(dxh) - xh mean_c(dxh xh)) ; R (+)= dx
// Kernel 1: one warp per row (dx into R).  Kernel 2: column strips per sample (dshift/dscale), deterministic.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ln_modulate_bwd_rows_kernel(const T* __restrict__ dh, const T* __restrict__ x, T* R,
                                                                   const float2* __restrict__ stats, const float* __restrict__ scale,
                                                                   int64_t ldmod, int m, int d, int tokens, int accumulate) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  const float2 st = stats[row];
  const int64_t n = row / tokens;
  const size_t off = (size_t)row * d;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float g = ld_act(dh + off + i) * (1.0f + scale[n * ldmod + i]);
    const float xh = (ld_act(x + off + i) - st.x) * st.y;
    s1 += g;
    s2 = __fmaf_rn(g, xh, s2);
  }
  s1 = warp_sum(s1) / (float)d;
  s2 = warp_sum(s2) / (float)d;
  for (int i = lane; i < d; i += 32) {
    const float g = ld_act(dh + off + i) * (1.0f + scale[n * ldmod + i]);
    const float xh = (ld_act(x + off + i) - st.x) * st.y;
    float dx = st.y * (g - s1 - xh * s2);
    if (accumulate) dx += ld_act(R + off + i);
    st_act(R + off + i, dx);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) ln_modulate_bwd_cols_kernel(const T* __restrict__ dh, const T* __restrict__ x,
                                                                   const float2* __restrict__ stats, float* __restrict__ dshift,
                                                                   float* __restrict__ dscale, int64_t ldmod, int d, int tokens) {
  const int n = blockIdx.y, c = blockIdx.x * 256 + threadIdx.x;
  if (c >= d) return;
  float a = 0.f, b = 0.f;
  for (int t = 0; t < tokens; ++t) {
    const size_t row = (size_t)n * tokens + t;
    const float2 st = stats[row];
    const float g = ld_act(dh + row * d + c);
    a += g;
    b = __fmaf_rn(g, (ld_act(x + row * d + c) - st.x) * st.y, b);
  }
  dshift[n * ldmod + c] = a;
  dscale[n * ldmod + c] = b;
}
extern "C" int mapdit_ln_modulate_bwd(const void* dh, const void* x, void* R, const float* stats, const float* scale, float* dshift,
                                      float* dscale, int64_t ldmod, int n_samples, int d, int tokens, int accumulate, int dtype,
                                      void* stream) {
  MAPDIT_REQUIRE(dh && x && stats && scale && dshift && dscale && n_samples > 0 && d > 0 && tokens > 0, "ln_modulate_bwd: bad args");
  const int m = n_samples * tokens;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 gc((d + 255) / 256, n_samples);
  if (dtype == MAPDIT_F32) {
    if (R) ln_modulate_bwd_rows_kernel<float><<<(m + 7) / 8, 256, 0, s>>>((const float*)dh, (const float*)x, (float*)R, (const float2*)stats, scale, ldmod, m, d, tokens, accumulate);
    ln_modulate_bwd_cols_kernel<float><<<gc, 256, 0, s>>>((const float*)dh, (const float*)x, (const float2*)stats, dshift, dscale, ldmod, d, tokens);
  } else {
    if (R) ln_modulate_bwd_rows_kernel<bf16><<<(m + 7) / 8, 256, 0, s>>>((const bf16*)dh, (const bf16*)x, (bf16*)R, (const float2*)stats, scale, ldmod, m, d, tokens, accumulate);
    ln_modulate_bwd_cols_kernel<bf16><<<gc, 256, 0, s>>>((const bf16*)dh, (const bf16*)x, (const float2*)stats, dshift, dscale, ldmod, d, tokens);
  }
  MAPDIT_LAUNCH_CHECK("ln_modulate_bwd");
  mapdit_count_launch(R ? 1 : 0);
  return MAPDIT_OK;
}
