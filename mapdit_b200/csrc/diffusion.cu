// K5 / K6: fused diffusion-step update and fused q_sample + loss (forward and gradient).
// Coefficient table layout: include/mapdit.h (MAPDIT_DIFF_ROWS x steps, fp32).
// Arithmetic is written with explicit round-to-nearest mul/add (no FMA contraction) so the fp32
// results follow the reference's op-by-op PyTorch evaluation as closely as possible.
#include "common.cuh"

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))

// ------------------------------------------------------------------------------------------------
// K5  (diffusion/gaussian_diffusion.py:285-293 log-variance, :334-339 x0 from eps, :310-315 clip,
//      :238-241 posterior mean, :410-416 noise add)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) diffusion_step_kernel(const float* __restrict__ mo, const float* __restrict__ x,
                                                             const float* __restrict__ noise, const int64_t* __restrict__ t,
                                                             const float* __restrict__ tab, int steps, float* __restrict__ sample,
                                                             float* __restrict__ x0out, int64_t total, int chw, int clip) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / chw;
  int r = (int)(i - n * chw);
  int ti = (int)t[n];
  const float srac = tab[2 * steps + ti], srm1 = tab[3 * steps + ti], c1 = tab[4 * steps + ti], c2 = tab[5 * steps + ti];
  const float minlog = tab[6 * steps + ti], maxlog = tab[7 * steps + ti];
  const float eps = mo[n * 2 * chw + r], v = mo[n * 2 * chw + chw + r];
  const float xv = x[i];
  float frac = FA(v, 1.0f) / 2.0f;
  float logvar = FA(FM(frac, maxlog), FM(FS(1.0f, frac), minlog));
  float x0 = FS(FM(srac, xv), FM(srm1, eps));
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  float mean = FA(FM(c1, x0), FM(c2, xv));
  float mask = (ti != 0) ? 1.0f : 0.0f;
  float sd = expf(FM(0.5f, logvar));
  sample[i] = FA(mean, FM(FM(mask, sd), noise[i]));
  if (x0out) x0out[i] = x0;
}

extern "C" int mapdit_diffusion_step(const float* model_out, const float* x, const float* noise, const int64_t* t,
                                     const float* tables, int steps, float* sample, float* pred_xstart, int n_samples,
                                     int channels, int hw, int clip_denoised, void* stream) {
  MAPDIT_REQUIRE(model_out && x && noise && t && tables && sample && n_samples > 0 && steps > 0, "diffusion_step: bad args");
  int chw = channels * hw;
  int64_t total = (int64_t)n_samples * chw;
  diffusion_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      model_out, x, noise, t, tables, steps, sample, pred_xstart, total, chw, clip_denoised);
  MAPDIT_LAUNCH_CHECK("diffusion_step");
  return MAPDIT_OK;
}

// DDIM step (diffusion/gaussian_diffusion.py:513-560), "next" row N4: x0 from eps (clip), eps re-derived from the clipped x0,
// sigma = eta sqrt((1-abar_prev)/(1-abar)) sqrt(1-abar/abar_prev), sample = x0 sqrt(abar_prev) + sqrt(1-abar_prev-sigma^2) eps + [t!=0] sigma noise
__global__ void __launch_bounds__(256) ddim_step_kernel(const float* __restrict__ mo, const float* __restrict__ x,
                                                        const float* __restrict__ noise, const int64_t* __restrict__ t,
                                                        const float* __restrict__ tab, int steps, float* __restrict__ sample,
                                                        float* __restrict__ x0out, int64_t total, int chw, int clip, float eta) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / chw;
  int r = (int)(i - n * chw);
  int ti = (int)t[n];
  const float srac = tab[2 * steps + ti], srm1 = tab[3 * steps + ti], ab = tab[8 * steps + ti], abp = tab[9 * steps + ti];
  const float eps_m = mo[n * 2 * chw + r];
  const float xv = x[i];
  float x0 = FS(FM(srac, xv), FM(srm1, eps_m));
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  const float eps = FS(FM(srac, xv), x0) / srm1;
  const float sigma = FM(FM(eta, sqrtf(FS(1.0f, abp) / FS(1.0f, ab))), sqrtf(FS(1.0f, ab / abp)));
  const float mean = FA(FM(x0, sqrtf(abp)), FM(sqrtf(FS(FS(1.0f, abp), FM(sigma, sigma))), eps));
  const float mask = (ti != 0) ? 1.0f : 0.0f;
  sample[i] = FA(mean, FM(FM(mask, sigma), noise ? noise[i] : 0.0f));
  if (x0out) x0out[i] = x0;
}
extern "C" int mapdit_ddim_step(const float* model_out, const float* x, const float* noise, const int64_t* t, const float* tables,
                                int steps, float* sample, float* pred_xstart, int n_samples, int channels, int hw, int clip_denoised,
                                float eta, void* stream) {
  MAPDIT_REQUIRE(model_out && x && t && tables && sample && n_samples > 0 && steps > 0, "ddim_step: bad args");
  int chw = channels * hw;
  int64_t total = (int64_t)n_samples * chw;
  ddim_step_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(model_out, x, noise, t, tables, steps, sample,
                                                                                      pred_xstart, total, chw, clip_denoised, eta);
  MAPDIT_LAUNCH_CHECK("ddim_step");
  return MAPDIT_OK;
}

// p_mean_variance as separate tensors (gaussian_diffusion.py:254-332), for callers that want the dict
__global__ void __launch_bounds__(256) p_mean_variance_kernel(const float* __restrict__ mo, const float* __restrict__ x,
                                                              const int64_t* __restrict__ t, const float* __restrict__ tab, int steps,
                                                              float* __restrict__ mean_o, float* __restrict__ var_o,
                                                              float* __restrict__ logvar_o, float* __restrict__ x0_o, int64_t total,
                                                              int chw, int clip) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t n = i / chw;
  int r = (int)(i - n * chw);
  int ti = (int)t[n];
  const float srac = tab[2 * steps + ti], srm1 = tab[3 * steps + ti], c1 = tab[4 * steps + ti], c2 = tab[5 * steps + ti];
  const float minlog = tab[6 * steps + ti], maxlog = tab[7 * steps + ti];
  const float eps = mo[n * 2 * chw + r], v = mo[n * 2 * chw + chw + r];
  const float xv = x[i];
  float frac = FA(v, 1.0f) / 2.0f;
  float logvar = FA(FM(frac, maxlog), FM(FS(1.0f, frac), minlog));
  float x0 = FS(FM(srac, xv), FM(srm1, eps));
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  mean_o[i] = FA(FM(c1, x0), FM(c2, xv));
  var_o[i] = expf(logvar);
  logvar_o[i] = logvar;
  x0_o[i] = x0;
}
extern "C" int mapdit_p_mean_variance(const float* model_out, const float* x, const int64_t* t, const float* tables, int steps,
                                      float* mean, float* variance, float* log_variance, float* pred_xstart, int n_samples,
                                      int channels, int hw, int clip_denoised, void* stream) {
  MAPDIT_REQUIRE(model_out && x && t && tables && mean && variance && log_variance && pred_xstart && n_samples > 0,
                 "p_mean_variance: bad args");
  int chw = channels * hw;
  int64_t total = (int64_t)n_samples * chw;
  p_mean_variance_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      model_out, x, t, tables, steps, mean, variance, log_variance, pred_xstart, total, chw, clip_denoised);
  MAPDIT_LAUNCH_CHECK("p_mean_variance");
  return MAPDIT_OK;
}

// posterior mean c1*x0 + c2*x_t (gaussian_diffusion.py:238-241) and the noise add of p_sample (:410-416)
__global__ void posterior_mean_kernel(const float* __restrict__ x0, const float* __restrict__ x, const int64_t* __restrict__ t,
                                      const float* __restrict__ tab, int steps, float* __restrict__ mean, int64_t total, int chw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ti = (int)t[i / chw];
  mean[i] = FA(FM(tab[4 * steps + ti], x0[i]), FM(tab[5 * steps + ti], x[i]));
}
extern "C" int mapdit_posterior_mean(const float* x0, const float* x, const int64_t* t, const float* tables, int steps, float* mean,
                                     int n_samples, int chw, void* stream) {
  MAPDIT_REQUIRE(x0 && x && t && tables && mean && n_samples > 0, "posterior_mean: bad args");
  int64_t total = (int64_t)n_samples * chw;
  posterior_mean_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x0, x, t, tables, steps, mean, total, chw);
  MAPDIT_LAUNCH_CHECK("posterior_mean");
  return MAPDIT_OK;
}
__global__ void noise_add_kernel(const float* __restrict__ mean, const float* __restrict__ logvar, const float* __restrict__ noise,
                                 const int64_t* __restrict__ t, float* __restrict__ sample, int64_t total, int chw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float mask = (t[i / chw] != 0) ? 1.0f : 0.0f;
  sample[i] = FA(mean[i], FM(FM(mask, expf(FM(0.5f, logvar[i]))), noise[i]));
}
extern "C" int mapdit_noise_add(const float* mean, const float* log_variance, const float* noise, const int64_t* t, float* sample,
                                int n_samples, int chw, void* stream) {
  MAPDIT_REQUIRE(mean && log_variance && noise && t && sample && n_samples > 0, "noise_add: bad args");
  int64_t total = (int64_t)n_samples * chw;
  noise_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(mean, log_variance, noise, t, sample, total, chw);
  MAPDIT_LAUNCH_CHECK("noise_add");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// q_sample (diffusion/gaussian_diffusion.py:215-230)
// ------------------------------------------------------------------------------------------------
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const int64_t* __restrict__ t,
                                const float* __restrict__ tab, int steps, float* __restrict__ xt, int64_t total, int chw) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int ti = (int)t[i / chw];
  xt[i] = FA(FM(tab[ti], x0[i]), FM(tab[steps + ti], noise[i]));
}
extern "C" int mapdit_q_sample(const float* x0, const float* noise, const int64_t* t, const float* tables, int steps, float* x_t,
                               int n_samples, int chw, void* stream) {
  MAPDIT_REQUIRE(x0 && noise && t && tables && x_t && n_samples > 0, "q_sample: bad args");
  int64_t total = (int64_t)n_samples * chw;
  q_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x0, noise, t, tables, steps, x_t, total, chw);
  MAPDIT_LAUNCH_CHECK("q_sample");
  return MAPDIT_OK;
}

// ------------------------------------------------------------------------------------------------
// K6 loss: one CTA per sample.
//   mse = mean (noise - eps)^2                                   (gaussian_diffusion.py:771-779)
//   vb  = where(t==0, nll, kl)/ln2 with eps detached             (:682-713, :753-765)
//   grad: d mse/d eps, d vb/d v (through the interpolated log-variance only)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float approx_cdf(float u, float* dcdf) {
  const float c = 0.7978845608028654f;  // sqrt(2/pi)
  float u3 = u * u * u;
  float w = c * (u + 0.044715f * u3);
  float th = tanhf(w);
  *dcdf = 0.5f * (1.0f - th * th) * c * (1.0f + 3.0f * 0.044715f * u * u);
  return 0.5f * (1.0f + th);
}

__global__ void __launch_bounds__(256) loss_kernel(const float* __restrict__ mo, const float* __restrict__ x0p,
                                                   const float* __restrict__ xtp, const float* __restrict__ noisep,
                                                   const int64_t* __restrict__ t, const float* __restrict__ tab, int steps,
                                                   float* __restrict__ loss, float* __restrict__ mse_o, float* __restrict__ vb_o,
                                                   float* __restrict__ gout, const float* __restrict__ gs_eps,
                                                   const float* __restrict__ gs_var, int chw) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const int ti = (int)t[n];
  const float srac = tab[2 * steps + ti], srm1 = tab[3 * steps + ti], c1 = tab[4 * steps + ti], c2 = tab[5 * steps + ti];
  const float minlog = tab[6 * steps + ti], maxlog = tab[7 * steps + ti];
  const float inv_cnt = 1.0f / (float)chw;
  const float LN2 = 0.6931471805599453f;
  float s_mse = 0.f, s_vb = 0.f;
  const float ge = (gout && gs_eps) ? gs_eps[n] : 1.0f, gv = (gout && gs_var) ? gs_var[n] : 1.0f;
  for (int r = threadIdx.x; r < chw; r += blockDim.x) {
    const size_t i = (size_t)n * chw + r;
    const float eps = mo[(size_t)n * 2 * chw + r], v = mo[(size_t)n * 2 * chw + chw + r];
    const float x0 = x0p[i], xt = xtp[i], nz = noisep[i];
    float d = FS(nz, eps);
    s_mse = FA(s_mse, FM(d, d));
    // model distribution (no clipping in the loss path, gaussian_diffusion.py:764)
    float frac = FA(v, 1.0f) / 2.0f;
    float lvp = FA(FM(frac, maxlog), FM(FS(1.0f, frac), minlog));
    float px0 = FS(FM(srac, xt), FM(srm1, eps));
    float mp = FA(FM(c1, px0), FM(c2, xt));
    float mq = FA(FM(c1, x0), FM(c2, xt));
    float term, dterm_dlv;
    if (ti != 0) {
      float dm = FS(mq, mp);
      float e1 = expf(FS(minlog, lvp));          // exp(lv_q - lv_p), lv_q = posterior_log_variance_clipped
      float e2 = FM(FM(dm, dm), expf(-lvp));
      term = 0.5f * (-1.0f + lvp - minlog + e1 + e2);   // diffusion_utils.py:30-36
      dterm_dlv = 0.5f * (1.0f - e1 - e2);
    } else {
      // discretised Gaussian NLL (diffusion_utils.py:62-88), log_scales = 0.5*lvp
      float cx = FS(x0, mp);
      float inv = expf(-0.5f * lvp);
      float pin = inv * (cx + 1.0f / 255.0f), min_ = inv * (cx - 1.0f / 255.0f);
      float dcp, dcm;
      float cp = approx_cdf(pin, &dcp), cm = approx_cdf(min_, &dcm);
      // d pin / d lvp = -0.5 pin
      float gp = dcp * (-0.5f * pin), gm = dcm * (-0.5f * min_);
      float lp, dlp;
      if (x0 < -0.999f) {
        lp = logf(fmaxf(cp, 1e-12f));
        dlp = (cp >= 1e-12f) ? gp / cp : 0.f;
      } else if (x0 > 0.999f) {
        float om = 1.0f - cm;
        lp = logf(fmaxf(om, 1e-12f));
        dlp = (om >= 1e-12f) ? -gm / om : 0.f;
      } else {
        float dl = cp - cm;
        lp = logf(fmaxf(dl, 1e-12f));
        dlp = (dl >= 1e-12f) ? (gp - gm) / dl : 0.f;
      }
      term = -lp;
      dterm_dlv = -dlp;
    }
    s_vb += term;
    if (gout) {
      gout[(size_t)n * 2 * chw + r] = ge * (-2.0f * d * inv_cnt);
      gout[(size_t)n * 2 * chw + chw + r] = gv * dterm_dlv * 0.5f * (maxlog - minlog) * inv_cnt / LN2;
    }
  }
  s_mse = block_sum(s_mse, red);
  s_vb = block_sum(s_vb, red);
  if (threadIdx.x == 0) {
    float m = s_mse * inv_cnt, vb = (s_vb * inv_cnt) / LN2;
    if (mse_o) mse_o[n] = m;
    if (vb_o) vb_o[n] = vb;
    if (loss) loss[n] = m + vb;
  }
}

extern "C" int mapdit_loss_fwd_bwd(const float* model_out, const float* x0, const float* x_t, const float* noise, const int64_t* t,
                                   const float* tables, int steps, float* loss, float* mse, float* vb, float* grad_out,
                                   const float* gs_eps, const float* gs_var, int n_samples, int channels, int hw, void* stream) {
  MAPDIT_REQUIRE(model_out && x0 && x_t && noise && t && tables && n_samples > 0, "loss_fwd_bwd: bad args");
  loss_kernel<<<n_samples, 256, 0, (cudaStream_t)stream>>>(model_out, x0, x_t, noise, t, tables, steps, loss, mse, vb, grad_out,
                                                           gs_eps, gs_var, channels * hw);
  MAPDIT_LAUNCH_CHECK("loss_fwd_bwd");
  return MAPDIT_OK;
}
