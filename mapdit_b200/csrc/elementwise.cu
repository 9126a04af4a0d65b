// Elementwise / row-wise kernels of the MaP-DiT hot path (HBM-bound; vectorised where it matters).
// Each entry point cites the reference code it replaces in include/mapdit.h.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// error / bookkeeping
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void mapdit_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void mapdit_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* mapdit_last_error(void) { return g_err; }
extern "C" int mapdit_abi_version(void) { return 1; }
extern "C" int mapdit_sizeof_gemm_args(void) { return (int)sizeof(mapdit_gemm_args); }
extern "C" int64_t mapdit_launch_count(void) { return (int64_t)g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// K1 weight normalisation.  One CTA per row.  The row is re-read from L1/L2 in each pass
// (rows are <= 18 KB), so DRAM traffic stays at 4 B read + outputs per element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weight_norm_fwd_kernel(float* __restrict__ w, int cols, float eps, int force,
                                                              float* __restrict__ eff_f32, bf16* __restrict__ eff_bf16,
                                                              bf16* __restrict__ eff_bf16_t, float* __restrict__ inv_norm,
                                                              int rows, int64_t ld_t) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  float* row = w + (size_t)r * cols;
  float ss = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float v = row[i];
    ss = __fmaf_rn(v, v, ss);
  }
  float nrm = sqrtf(block_sum(ss, red));
  float inv;
  if (force) {
    // w <- w*sqrt(n)/(||w||+eps)  (src/utils.py:19-23), then eff is computed from the forced row
    const float sq = sqrtf((float)cols);
    const float den = nrm + eps;
    float ss2 = 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float v = (row[i] * sq) / den;
      ss2 = __fmaf_rn(v, v, ss2);
    }
    float nrm2 = sqrtf(block_sum(ss2, red));
    inv = 1.0f / (nrm2 + eps);
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float v = (row[i] * sq) / den;
      row[i] = v;
      float e = v * inv;
      if (eff_f32) eff_f32[(size_t)r * cols + i] = e;
      if (eff_bf16) eff_bf16[(size_t)r * cols + i] = __float2bfloat16_rn(e);
      if (eff_bf16_t) eff_bf16_t[(size_t)i * ld_t + r] = __float2bfloat16_rn(e);
    }
  } else {
    inv = 1.0f / (nrm + eps);
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float e = row[i] * inv;
      if (eff_f32) eff_f32[(size_t)r * cols + i] = e;
      if (eff_bf16) eff_bf16[(size_t)r * cols + i] = __float2bfloat16_rn(e);
      if (eff_bf16_t) eff_bf16_t[(size_t)i * ld_t + r] = __float2bfloat16_rn(e);
    }
  }
  if (inv_norm && threadIdx.x == 0) inv_norm[r] = inv;
}

extern "C" int mapdit_weight_norm_fwd(float* w, int rows, int cols, float eps, int force, float* eff_f32, void* eff_bf16,
                                      void* eff_bf16_t, int64_t ld_t, float* inv_norm, void* stream) {
  MAPDIT_REQUIRE(w && rows > 0 && cols > 0, "weight_norm_fwd: bad args");
  weight_norm_fwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(w, cols, eps, force, eff_f32, (bf16*)eff_bf16,
                                                                  (bf16*)eff_bf16_t, inv_norm, rows, ld_t > 0 ? ld_t : rows);
  MAPDIT_LAUNCH_CHECK("weight_norm_fwd");
  return MAPDIT_OK;
}

// grad_v = (G - v (v.G)/(r (r+eps)))/(r+eps)     (autograd of src/utils.py:19-23 / sqrt(n))
__global__ void __launch_bounds__(256) weight_norm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                              float* __restrict__ gv, int cols, float eps, int accumulate) {
  __shared__ float red[32];
  const size_t off = (size_t)blockIdx.x * cols;
  float ss = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float a = v[off + i];
    ss = __fmaf_rn(a, a, ss);
    dot = __fmaf_rn(a, g[off + i], dot);
  }
  float r = sqrtf(block_sum(ss, red));
  dot = block_sum(dot, red);
  float inv = 1.0f / (r + eps);
  float coef = dot / (fmaxf(r, 1e-30f) * (r + eps));
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float val = (g[off + i] - v[off + i] * coef) * inv;
    gv[off + i] = accumulate ? gv[off + i] + val : val;
  }
}

extern "C" int mapdit_weight_norm_bwd(const float* v, const float* g_eff, float* grad_v, int rows, int cols, float eps,
                                      int accumulate, void* stream) {
  MAPDIT_REQUIRE(v && g_eff && grad_v && rows > 0 && cols > 0, "weight_norm_bwd: bad args");
  weight_norm_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(v, g_eff, grad_v, cols, eps, accumulate);
  MAPDIT_LAUNCH_CHECK("weight_norm_bwd");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// K3 standalone: modulate, residual, mp_silu, qk-normalise, casts
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void modulate_kernel(const T* __restrict__ x, T* __restrict__ h, const float* __restrict__ shift,
                                const float* __restrict__ scale, const float* __restrict__ gain, int64_t ldmod, int64_t total,
                                int d, int tokens) {
  const float g = *gain;
  const float den = mod_den(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d;
    int col = (int)(i - row * d);
    int64_t n = row / tokens;
    st_act(h + i, modulate_f(ld_act(x + i), shift[n * ldmod + col], scale[n * ldmod + col], g, den));
  }
}

extern "C" int mapdit_modulate_fwd(const void* x, void* h, const float* shift, const float* scale, const float* gain,
                                   int64_t ldmod, int m, int d, int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && h && shift && scale && gain && m > 0 && d > 0 && tokens > 0, "modulate_fwd: bad args");
  int64_t total = (int64_t)m * d;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == MAPDIT_F32)
    modulate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)h, shift, scale, gain, ldmod, total, d, tokens);
  else
    modulate_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)h, shift, scale, gain, ldmod, total, d, tokens);
  MAPDIT_LAUNCH_CHECK("modulate_fwd");
  return MAPDIT_OK;
}

template <typename T>
__global__ void resid_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ xo, const float* __restrict__ gate,
                             int64_t ldmod, int64_t total, int d, int tokens) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / d;
    int col = (int)(i - row * d);
    int64_t n = row / tokens;
    st_act(xo + i, resid_f(ld_act(x + i), gate[n * ldmod + col], ld_act(y + i)));
  }
}

extern "C" int mapdit_resid_fwd(const void* x, const void* y, void* xout, const float* gate, int64_t ldmod, int m, int d,
                                int tokens, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && y && xout && gate && m > 0 && d > 0 && tokens > 0, "resid_fwd: bad args");
  int64_t total = (int64_t)m * d;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (dtype == MAPDIT_F32)
    resid_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)y, (float*)xout, gate, ldmod, total, d, tokens);
  else
    resid_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)y, (bf16*)xout, gate, ldmod, total, d, tokens);
  MAPDIT_LAUNCH_CHECK("resid_fwd");
  return MAPDIT_OK;
}

template <typename TI, typename TO>
__global__ void mp_silu_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_act(y + i, mp_silu_f(ld_act(x + i)));
}
template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_act(y + i, ld_act(x + i));
}

static inline int ew_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

extern "C" int mapdit_mp_silu_fwd(const void* x, void* y, int64_t n, int in_dtype, int out_dtype, void* stream) {
  MAPDIT_REQUIRE(x && y && n > 0, "mp_silu_fwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int grid = ew_grid(n);
  if (in_dtype == MAPDIT_F32 && out_dtype == MAPDIT_F32) mp_silu_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, n);
  else if (in_dtype == MAPDIT_F32) mp_silu_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)x, (bf16*)y, n);
  else if (out_dtype == MAPDIT_F32) mp_silu_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)x, (float*)y, n);
  else mp_silu_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)y, n);
  MAPDIT_LAUNCH_CHECK("mp_silu_fwd");
  return MAPDIT_OK;
}

extern "C" int mapdit_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  MAPDIT_REQUIRE(src && dst && n > 0, "cast: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  int grid = ew_grid(n);
  if (src_dtype == MAPDIT_F32 && dst_dtype == MAPDIT_BF16) cast_kernel<float, bf16><<<grid, 256, 0, s>>>((const float*)src, (bf16*)dst, n);
  else if (src_dtype == MAPDIT_BF16 && dst_dtype == MAPDIT_F32) cast_kernel<bf16, float><<<grid, 256, 0, s>>>((const bf16*)src, (float*)dst, n);
  else if (src_dtype == MAPDIT_F32) cast_kernel<float, float><<<grid, 256, 0, s>>>((const float*)src, (float*)dst, n);
  else cast_kernel<bf16, bf16><<<grid, 256, 0, s>>>((const bf16*)src, (bf16*)dst, n);
  MAPDIT_LAUNCH_CHECK("cast");
  return MAPDIT_OK;
}

// one warp per (row, head) of the q and k thirds of qkv[M, 3D]
template <typename T>
__global__ void qk_normalize_kernel(T* __restrict__ qkv, int64_t n_heads_total, int heads2, int d, int hd, float eps) {
  int lane = threadIdx.x & 31;
  int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= n_heads_total) return;
  int64_t row = wid / heads2;
  int hh = (int)(wid - row * heads2);  // 0..2H-1 : q heads then k heads (contiguous columns [0, 2D))
  T* p = qkv + row * (int64_t)(3 * d) + (int64_t)hh * hd;
  float v[4];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    v[j] = (c < hd) ? ld_act(p + c) : 0.f;
    ss = __fmaf_rn(v[j], v[j], ss);
  }
  float nrm = sqrtf(warp_sum(ss));
  float sq = sqrtf((float)hd), den = nrm + eps;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int c = lane + 32 * j;
    if (c < hd) st_act(p + c, (v[j] * sq) / den);
  }
}

extern "C" int mapdit_qk_normalize(void* qkv, int m, int d, int head_dim, float eps, int dtype, void* stream) {
  MAPDIT_REQUIRE(qkv && m > 0 && d > 0 && head_dim > 0 && head_dim <= 128 && d % head_dim == 0, "qk_normalize: bad args");
  int heads2 = 2 * (d / head_dim);
  int64_t total = (int64_t)m * heads2;
  int64_t blocks = (total * 32 + 255) / 256;
  if (dtype == MAPDIT_F32)
    qk_normalize_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)qkv, total, heads2, d, head_dim, eps);
  else
    qk_normalize_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)qkv, total, heads2, d, head_dim, eps);
  MAPDIT_LAUNCH_CHECK("qk_normalize");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// patch embed: x0 = mp_sum(patchify(x)|1 · Wx_eff^T, pos, 0.5)  (+ optional modulate for block 0)
// CTA = 16 tokens x 256 output channels; patches staged in smem.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) patch_embed_kernel(const float* __restrict__ x, const float* __restrict__ wx,
                                                          const float* __restrict__ pos, T* __restrict__ x0, T* __restrict__ h,
                                                          const float* __restrict__ shift, const float* __restrict__ scale,
                                                          const float* __restrict__ gain, int64_t ldmod, int64_t m_total, int C,
                                                          int S, int p, int d) {
  extern __shared__ float sp[];  // [16][K1]
  const int g = S / p, T_ = g * g, K = p * p * C, K1 = K + 1;
  const int64_t tok0 = (int64_t)blockIdx.x * 16;
  for (int i = threadIdx.x; i < 16 * K1; i += blockDim.x) {
    int tl = i / K1, f = i - tl * K1;
    int64_t tok = tok0 + tl;
    float v = 0.f;
    if (tok < m_total) {
      if (f == K) v = 1.0f;  // bias column (src/dit.py:82)
      else {
        int64_t n = tok / T_;
        int tt = (int)(tok - n * T_);
        int hh = tt / g, ww = tt - hh * g;
        int c = f % C, pp = f / C, p1 = pp / p, p2 = pp - p1 * p;
        v = x[((n * C + c) * S + (hh * p + p1)) * (int64_t)S + (ww * p + p2)];
      }
    }
    sp[i] = v;
  }
  __syncthreads();
  const int col = blockIdx.y * 256 + threadIdx.x;
  if (col >= d) return;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const float* wr = wx + (size_t)col * K1;
  for (int k = 0; k < K1; ++k) {
    float w = wr[k];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = __fmaf_rn(sp[j * K1 + k], w, acc[j]);
  }
  float gn = 0.f, den = 1.f;
  if (h) { gn = *gain; den = mod_den(gn); }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    int64_t tok = tok0 + j;
    if (tok >= m_total) break;
    int64_t n = tok / T_;
    int tt = (int)(tok - n * T_);
    float v = lerp_t(acc[j], pos[(size_t)tt * d + col], 0.5f) / MP_HALF_DEN;
    st_act(x0 + tok * d + col, v);
    if (h) {
      // modulate sees the value as stored in the residual stream
      float xs = ld_act(x0 + tok * d + col);
      st_act(h + tok * d + col, modulate_f(xs, shift[n * ldmod + col], scale[n * ldmod + col], gn, den));
    }
  }
}

extern "C" int mapdit_patch_embed(const float* x, const float* wx_eff, const float* pos, void* x0, void* h, const float* shift,
                                  const float* scale, const float* gain, int64_t ldmod, int n_samples, int channels,
                                  int input_size, int patch, int d, int dtype, void* stream) {
  MAPDIT_REQUIRE(x && wx_eff && pos && x0 && n_samples > 0 && input_size % patch == 0, "patch_embed: bad args");
  int g = input_size / patch;
  int64_t m_total = (int64_t)n_samples * g * g;
  int K1 = patch * patch * channels + 1;
  size_t smem = (size_t)16 * K1 * sizeof(float);
  dim3 grid((unsigned)((m_total + 15) / 16), (unsigned)((d + 255) / 256));
  if (dtype == MAPDIT_F32)
    patch_embed_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(x, wx_eff, pos, (float*)x0, (float*)h, shift, scale, gain, ldmod, m_total, channels, input_size, patch, d);
  else
    patch_embed_kernel<bf16><<<grid, 256, smem, (cudaStream_t)stream>>>(x, wx_eff, pos, (bf16*)x0, (bf16*)h, shift, scale, gain, ldmod, m_total, channels, input_size, patch, d);
  MAPDIT_LAUNCH_CHECK("patch_embed");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// conditioning path
// ------------------------------------------------------------------------------------------------
__global__ void fourier_kernel(const int64_t* __restrict__ t, const float* __restrict__ scale, const float* __restrict__ shift,
                               float* __restrict__ e, int n, int ch) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ch) return;
  int r = i / ch, j = i - r * ch;
  // no FMA contraction: fl(fl(t*scale)+shift), accurate cosf
  float arg = __fadd_rn(__fmul_rn((float)t[r], scale[j]), shift[j]);
  e[i] = 1.4142135623730951f * cosf(arg);
}
extern "C" int mapdit_fourier(const int64_t* t, const float* scale, const float* shift, float* e, int n, int channels, void* stream) {
  MAPDIT_REQUIRE(t && scale && shift && e && n > 0 && channels > 0, "fourier: bad args");
  fourier_kernel<<<(n * channels + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, scale, shift, e, n, channels);
  MAPDIT_LAUNCH_CHECK("fourier");
  return MAPDIT_OK;
}

__global__ void __launch_bounds__(256) embed_rows_kernel(const int64_t* __restrict__ idx, const uint8_t* __restrict__ drop,
                                                         int64_t null_idx, const float* __restrict__ table, float* __restrict__ out,
                                                         int d, float eps) {
  __shared__ float red[32];
  int n = blockIdx.x;
  int64_t id = idx[n];
  if (drop && drop[n]) id = null_idx;
  const float* row = table + id * d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) ss = __fmaf_rn(row[i], row[i], ss);
  float nrm = sqrtf(block_sum(ss, red));
  float sq = sqrtf((float)d), den = nrm + eps;
  for (int i = threadIdx.x; i < d; i += blockDim.x) out[(size_t)n * d + i] = (row[i] * sq) / den;
}
extern "C" int mapdit_embed_rows(const int64_t* idx, const uint8_t* drop_mask, int64_t null_idx, const float* table, float* out,
                                 int n, int d, float eps, void* stream) {
  MAPDIT_REQUIRE(idx && table && out && n > 0 && d > 0, "embed_rows: bad args");
  embed_rows_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(idx, drop_mask, null_idx, table, out, d, eps);
  MAPDIT_LAUNCH_CHECK("embed_rows");
  return MAPDIT_OK;
}

__global__ void cond_combine_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c,
                                    float* __restrict__ cs32, bf16* __restrict__ cs16, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = lerp_t(a[i], b[i], 0.5f) / MP_HALF_DEN;
  if (c) c[i] = v;
  float s = mp_silu_f(v);
  if (cs32) cs32[i] = s;
  if (cs16) cs16[i] = __float2bfloat16_rn(s);
}
extern "C" int mapdit_cond_combine(const float* a, const float* b, float* c, float* cs_f32, void* cs_bf16, int64_t n, void* stream) {
  MAPDIT_REQUIRE(a && b && n > 0, "cond_combine: bad args");
  cond_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, c, cs_f32, (bf16*)cs_bf16, n);
  MAPDIT_LAUNCH_CHECK("cond_combine");
  return MAPDIT_OK;
}

// one CTA per sample: adim (<=32) dot products of length d, then sigmoid(sum_j dot_j*ref_j/sqrt(adim))
__global__ void __launch_bounds__(256) mp_scale_kernel(const float* __restrict__ c, const float* __restrict__ w,
                                                       const float* __restrict__ ref, float* __restrict__ s, int d, int adim) {
  __shared__ float red[32];
  int n = blockIdx.x;
  float angle = 0.f;
  for (int j = 0; j < adim; ++j) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) acc = __fmaf_rn(c[(size_t)n * d + i], w[(size_t)j * d + i], acc);
    acc = block_sum(acc, red);
    angle = __fmaf_rn(acc, ref[j], angle);
  }
  if (threadIdx.x == 0) {
    float a = angle / sqrtf((float)adim);
    s[n] = 1.0f / (1.0f + expf(-a));
  }
}
extern "C" int mapdit_mp_scale(const float* c, const float* w_eff, const float* ref, float* s, int n, int d, int adim, void* stream) {
  MAPDIT_REQUIRE(c && w_eff && ref && s && n > 0 && d > 0 && adim > 0, "mp_scale: bad args");
  mp_scale_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(c, w_eff, ref, s, d, adim);
  MAPDIT_LAUNCH_CHECK("mp_scale");
  return MAPDIT_OK;
}

// out[n, part*C + c, hh*p+p1, ww*p+p2] = lin[tok, part*p*p*C + (p1*p+p2)*C + c] * s_part[n]
template <typename T>
__global__ void final_unpatchify_kernel(const T* __restrict__ lin, const float* __restrict__ s_mu, const float* __restrict__ s_sg,
                                        float* __restrict__ out, int64_t total, int C, int S, int p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // i indexes the OUTPUT (coalesced writes): [n, 2C, S, S]
  int xx = (int)(i % S);
  int64_t r = i / S;
  int yy = (int)(r % S);
  r /= S;
  int ch = (int)(r % (2 * C));
  int64_t n = r / (2 * C);
  int part = ch / C, c = ch - part * C;
  int g = S / p, hh = yy / p, p1 = yy - hh * p, ww = xx / p, p2 = xx - ww * p;
  int64_t tok = n * (g * g) + hh * g + ww;
  int ppc = p * p * C;
  float v = ld_act(lin + tok * (2 * ppc) + part * ppc + (p1 * p + p2) * C + c);
  out[i] = v * (part ? s_sg[n] : s_mu[n]);
}
extern "C" int mapdit_final_unpatchify(const void* lin, const float* s_mu, const float* s_sigma, float* out, int n_samples,
                                       int channels, int input_size, int patch, int dtype, void* stream) {
  MAPDIT_REQUIRE(lin && s_mu && s_sigma && out && n_samples > 0, "final_unpatchify: bad args");
  int64_t total = (int64_t)n_samples * 2 * channels * input_size * input_size;
  unsigned grid = (unsigned)((total + 255) / 256);
  if (dtype == MAPDIT_F32)
    final_unpatchify_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)lin, s_mu, s_sigma, out, total, channels, input_size, patch);
  else
    final_unpatchify_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)lin, s_mu, s_sigma, out, total, channels, input_size, patch);
  MAPDIT_LAUNCH_CHECK("final_unpatchify");
  return MAPDIT_OK;
}

// eps channels of both halves <- uncond + s*(cond - uncond)   (src/dit.py:113-118)
__global__ void cfg_combine_kernel(float* __restrict__ out, int n_half, int C, int hw, float s) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per = (int64_t)C * hw;
  if (i >= (int64_t)n_half * per) return;
  int64_t n = i / per, r = i - n * per;
  int64_t a = n * (2 * per) + r, b = (n + n_half) * (2 * per) + r;
  float cond = out[a], unc = out[b];
  float v = unc + s * (cond - unc);
  out[a] = v;
  out[b] = v;
}
extern "C" int mapdit_cfg_combine(float* out, int n_half, int channels, int hw, float cfg_scale, void* stream) {
  MAPDIT_REQUIRE(out && n_half > 0, "cfg_combine: bad args");
  int64_t total = (int64_t)n_half * channels * hw;
  cfg_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out, n_half, channels, hw, cfg_scale);
  MAPDIT_LAUNCH_CHECK("cfg_combine");
  return MAPDIT_OK;
}
