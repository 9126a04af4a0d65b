/*
 * mapdit.h — C ABI of libmapdit.so: the B200 (sm_100a) kernels behind the MaP-DiT hot path.
 *
 * The reference (ericbill21/map-dit) is pure Python/PyTorch and has no FFI of its own; its
 * boundary is the Python API (src/models.py:50-56 DIT_MODELS, src/dit.py:70-118 DiT.forward /
 * forward_with_cfg, diffusion/respace.py + diffusion/gaussian_diffusion.py training_losses /
 * p_sample_loop).  The Python package mapdit_b200 mirrors that API and calls the entry points
 * below through ctypes with raw device pointers.  Each entry point names the reference code it
 * replaces.  Conventions:
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - row-major tensors; `ld*` are leading dimensions in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises or allocates;
 *   - return 0 on success, negative on error (text via mapdit_last_error()).
 *   - dtype codes: MAPDIT_F32 = 0, MAPDIT_BF16 = 1.
 */
#ifndef MAPDIT_H_
#define MAPDIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAPDIT_F32 0
#define MAPDIT_BF16 1

#define MAPDIT_OK 0
#define MAPDIT_ERR_ARG (-1)
#define MAPDIT_ERR_CUDA (-2)
#define MAPDIT_ERR_UNSUPPORTED (-3)

const char* mapdit_last_error(void);
int mapdit_abi_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t mapdit_launch_count(void);

/* ---- README "--use-*" switches turned OFF (README.md:59-66).  The reference snapshot hard-codes every switch on and
 * ships no "off" branch, so these variants are UNPINNED: their oracle is this repo's own
 * restatement of the vanilla DiT ops.  The word is per host thread and is read at launch time by the entry points
 * that have a variant (resid, mp_silu, cond_combine, patch_embed, embed_rows, the fused GEMM epilogues, attention
 * dispatch, and their backward kernels); returns the previous word.                                              */
#define MAPDIT_VAR_PLAIN_RESID 1  /* use_mp_residual=False : x + gate*y instead of mp_sum(x, gate*y, 0.3)            */
#define MAPDIT_VAR_PLAIN_SILU 2   /* use_mp_silu=False     : silu(x) instead of silu(x)/0.596                        */
#define MAPDIT_VAR_PLAIN_POS 4    /* use_mp_pos_enc=False  : x + pos instead of mp_sum(x, pos, 0.5)                  */
#define MAPDIT_VAR_PLAIN_EMBED 8  /* use_mp_embedding=False: plain table gather, c = t_emb + y_emb                   */
#define MAPDIT_VAR_DOT_ATTN 16    /* use_cosine_attention=False: unbounded logits -> running-max softmax kernels     */
int mapdit_set_variant(int flags);

/* ---- K1: weight normalisation --------------------------------------------------------------
 * Replaces normalize()/chunk_normalize() + the forced in-place normalisation
 * (src/utils.py:19-34, src/basic/mp_linear.py:37-46,67-75, src/basic/mp_embedding.py:16-22).
 * For each row r of w[rows, cols]:
 *   force > 0  : w[r] <- w[r]*sqrt(cols)/(||w[r]||+eps)   (written back in place, train mode)
 *   force < 0  : no normalisation at all, eff = w (use_weight_normalization=False, UNPINNED)
 *   eff = w[r]/(||w[r]||+eps) (= normalize(w)/sqrt(cols)), computed from the (forced) row and
 *   written to any of eff_f32 [rows, cols], eff_bf16 [rows, cols], eff_bf16_t [cols, ld_t >= rows]
 *   (transposed copy for dgrad).  inv_norm [rows] (optional) receives 1/(||w||+eps) of the
 *   row that eff was computed from (needed by the backward).                                  */
int mapdit_weight_norm_fwd(float* w, int rows, int cols, float eps, int force, float* eff_f32,
                           void* eff_bf16, void* eff_bf16_t, int64_t ld_t /* 0 = rows */, float* inv_norm, void* stream);
/* The same for many weights in two launches.  descs: device table of n_desc rows of 10 x int64
 * {w, eff_f32, eff_bf16, eff_bf16_t (pointers, nullable), ld_t, rows, cols (cols % 4 == 0, 16-byte aligned pointers), group0, tile0,
 * row0}: group0 = sum over earlier tensors of ceil(rows/8), tile0 = sum of ceil(rows/64)*ceil(cols/64), row0 = sum of rows;
 * scratch: float2[total rows]. */
int mapdit_weight_norm_fwd_multi(const void* descs, int n_desc, int total_groups, int total_tiles, float eps, int force,
                                 void* scratch, void* stream);
/* Backward of eff = v/(||v||+eps) per row: given G = dL/d eff [rows, cols]
 * and the (forced) weights v, grad_v = (G - v (v·G)/(r (r+eps)))/(r+eps); accumulate==0 overwrites. */
int mapdit_weight_norm_bwd(const float* v, const float* g_eff, float* grad_v, int rows, int cols,
                           float eps, int accumulate, void* stream);
/* the same for several tensors in ONE launch, in place (g holds G on entry and grad_v on return; cols % 4 == 0, 16-byte aligned).
 * table: int64[n_items][5] = {v ptr, g ptr, rows, cols, group0}; group0 = sum over earlier items of ceil(rows/8); n_groups = total */
int mapdit_weight_norm_bwd_multi(const void* table, int n_items, int n_groups, float eps, void* stream);

/* ---- fp32 GEMM (mode a: CUDA-core FFMA, strided) --------------------------------------------
 * C[m,n] (+)= sum_k A(m,k) * B(n,k), A element (m,k) at a[m*sam + k*sak], B element (n,k) at
 * b[n*sbn + k*sbk]; C row-major with leading dimension ldc.  Replaces F.linear
 * (src/basic/mp_linear.py:46,75) and its autograd transposes in the fp32 parity mode.         */
int mapdit_gemm_f32(const float* a, int64_t sam, int64_t sak, const float* b, int64_t sbn, int64_t sbk,
                    float* c, int64_t ldc, int m, int n, int k, int accumulate, void* stream);

/* ---- K2/K3: bf16 tcgen05 GEMM with fused epilogues ------------------------------------------
 * D = A[M,K] · B[N,K]^T, bf16 operands (K-major, row-major with lda/ldb), fp32 accumulation in
 * TMEM, TMA-fed, persistent.  Epilogues fuse the reference's elementwise neighbours:          */
#define MAPDIT_EPI_STORE 0      /* out = acc                                   (F.linear)                         */
#define MAPDIT_EPI_QKNORM 1     /* q,k heads L2-normalised in registers        (src/layers/attention.py:43-45)    */
#define MAPDIT_EPI_MPSILU 2     /* out = silu(acc)/0.596 (+ optional pre-act)  (src/basic/mp_silu.py:7)           */
#define MAPDIT_EPI_RESID_MOD 3  /* x' = mp_sum(x, gate*acc, .3); h = modulate(x', shift, scale, g)
                                   (src/blocks/dit_block.py:35-36, src/utils.py:11-16)                            */
#define MAPDIT_EPI_RESID 4      /* x' = mp_sum(x, gate*acc, .3) only                                              */
#define MAPDIT_EPI_SILU_BWD 5   /* out = acc * d/dz[silu(z)/0.596], z = `resid` (dgrad of fc2 fused with MPSiLU's backward) */
#define MAPDIT_EPI_RESID_ROT 6  /* x' = mp_sum(x, gate*acc, .3); h = R(theta) x' (* scale): rotation modulation (README.md:1,3 of the
                                   reference; no reference code, SURVEY.md §A.8 — UNPINNED).  `shift` points at the per-sample
                                   (cos, sin) table of mapdit_rot_table (leading dimension `ldrot`), `scale` may be null  */
#define MAPDIT_EPI_STORE_DELTA 7 /* out = acc (bf16) and aux[row, head] (fp32, [M, N/64]) = sum over the head's 64 columns of
                                  * bf16(acc) * resid: the out-proj dgrad that also emits delta = dO.O for the attention backward
                                  * (autograd of F.scaled_dot_product_attention, src/layers/attention.py:47); N % 64 == 0 */

typedef struct mapdit_gemm_args {
  const void* a;   /* bf16 [M, K] */
  const void* b;   /* bf16 [N, K] */
  void* out;       /* bf16 or f32 [M, N] (EPI_RESID*: the new residual stream x') */
  void* out2;      /* EPI_MPSILU: optional bf16 pre-activation [M,N]; EPI_RESID_MOD: h [M,N] bf16 */
  const void* resid;  /* EPI_RESID*: bf16 x [M, N] (may alias out) */
  const float* gate;  /* EPI_RESID*: per-sample fp32 [n_samples, ldmod] column slice start */
  const float* shift; /* EPI_RESID_MOD */
  const float* scale; /* EPI_RESID_MOD */
  const float* gain;  /* EPI_RESID_MOD: device scalar g (blocks.i.gain_*) */
  void* aux;          /* optional, for the backward: EPI_RESID*: bf16 [M,N] raw branch output (acc);
                         EPI_QKNORM: fp32 [M, qk_cols/head_dim] per-head scale sqrt(hd)/(||v||+eps) */
  int64_t lda, ldb, ldo, ldmod;
  int m, n, k;
  int tokens;      /* rows per sample (sample index = row / tokens) */
  int head_dim;    /* EPI_QKNORM */
  int qk_cols;     /* EPI_QKNORM: columns [0, qk_cols) are q|k heads, the rest (v) is stored as is */
  int epilogue;
  int out_dtype;   /* MAPDIT_BF16 or MAPDIT_F32 (EPI_STORE only) */
  float eps;
  int64_t ldrot;   /* EPI_RESID_ROT: leading dimension of the (cos, sin) table `shift` points into (0 = ldmod) */
} mapdit_gemm_args;

int mapdit_gemm_bf16(const mapdit_gemm_args* args, void* stream);
int mapdit_sizeof_gemm_args(void); /* lets a binding check its struct mirror */
/* runtime switches (A/B and developer use; defaults are the fast paths): "gemm_2cta" (0/1) selects the cta_group::2 256xBN kernel
 * for large-M GEMMs; "gemm_2cta_bn" (0 = auto, 128 / 192 / 256 = force that pair-tile width where it applies); "gemm_fused_resid" (0 = first-generation
 * residual epilogues, 1 = second generation where the main loop is short (default), 2 = always); "attn_v2" (0/1) selects the one-CTA-per-SM
 * ping-pong attention forward for tokens % 256 == 0; "attn_bwd_fused" (tokens == 256: 0 = dq + dkv kernel pair, 1 = single fused
 * kernel, 2 = fused kernel with a dedicated read-out warpgroup, 3 = that kernel with 64-query half-iterations and P^T / dS^T kept in
 * TMEM as MMA operands, the default) */
int mapdit_set_option(const char* name, int value);
/* developer hook: device buffer of >= 2048 int64 that one CTA of the attn_v2 / attn_bwd_fused / attention backward pair / 2-CTA GEMM
 * kernels fills with clock64 stamps (null = off); read by tools/attn_timeline.py, tools/attn_bwd_timeline.py,
 * tools/attn_bwd_xl_timeline.py, tools/gemm_timeline.py */
int mapdit_attn_debug_buffer(void* buf);
/* weight gradient C[N_out, K_in] (fp32) = dY[M, N_out]^T · X[M, K_in] on tcgen05, operands read MN-major in place
 * (autograd of F.linear, src/basic/mp_linear.py:46,75); split-K with fp32 vector reductions when N_out*K_in is small */
int mapdit_gemm_bf16_tn(const void* dy, int64_t ldy, const void* x, int64_t ldx, float* c, int64_t ldc, int m_tokens,
                        int n_out, int k_in, void* stream);
/* fused Adam step over one flat fp32 parameter span (torch.optim.Adam semantics, train.py:57,96):
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr * (m/bc1) / (sqrt(v/bc2) + eps) */
int mapdit_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float bias_corr1, float bias_corr2, float grad_scale, void* stream);
/* the same with bf16 gradients: the data-parallel path all-reduces a bf16 copy of the gradient span (train.py:91-96 under DDP) */
int mapdit_adam_step_g16(float* p, const void* g_bf16, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                         float bias_corr1, float bias_corr2, float grad_scale, void* stream);

/* N3, device-side CustomDataset.__getitem__ (train.py:168-176): out[n] = ((means[idx[n]] + eps[n]*stds[idx[n]]) - ch_mean[c]) / ch_std[c];
 * means/stds [items, C, H*W] resident in HBM, idx int64 [n], eps/out [n, C, H*W] */
int mapdit_latent_sample(const float* means, const float* stds, const int64_t* idx, const float* eps, const float* ch_mean,
                         const float* ch_std, float* out, int n, int channels, int hw, void* stream);

/* multi-tensor EMA update (src/ema.py:135-140): for each chunk {float* dst; const float* src; int64 n} of the device
 * table, dst = lerp(dst, src, weight) with torch.lerp's rounding */
int mapdit_multi_lerp(const void* chunk_table, int n_chunks, float weight, void* stream);

/* ---- K3 standalone elementwise ops (fp32 mode and fallbacks); dtype = activation dtype ------ */
/* h = modulate(x, shift, scale, g) = lerp(x*scale, shift, g)/sqrt((1-g)^2+g^2)  (src/utils.py:11-16) */
int mapdit_modulate_fwd(const void* x, void* h, const float* shift, const float* scale, const float* gain,
                        int64_t ldmod, int m, int d, int tokens, int dtype, void* stream);
/* xout = mp_sum(x, gate*y, 0.3) (src/blocks/dit_block.py:35-36) */
int mapdit_resid_fwd(const void* x, const void* y, void* xout, const float* gate, int64_t ldmod, int m, int d,
                     int tokens, int dtype, void* stream);
/* y = silu(x)/0.596 (src/basic/mp_silu.py:7); in/out dtypes independent */
int mapdit_mp_silu_fwd(const void* x, void* y, int64_t n, int in_dtype, int out_dtype, void* stream);
/* in-place L2 normalisation of the q and k heads of qkv[M, 3D] (src/layers/attention.py:43-45) */
int mapdit_qk_normalize(void* qkv, int m, int d, int head_dim, float eps, int dtype, void* stream);
/* dtype casts */
int mapdit_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream);
/* the same for a [rows, cols] window of row-major matrices with leading dimensions ld_src / ld_dst */
int mapdit_cast_2d(const void* src, int64_t ld_src, void* dst, int64_t ld_dst, int rows, int cols, int src_dtype, int dst_dtype,
                   void* stream);

/* ---- K4: cosine attention --------------------------------------------------------------------
 * o[M, D] = merge_heads(softmax(q^ k^T / sqrt(hd)) v) for qkv[M, 3D] whose q,k heads are already
 * normalised (src/layers/attention.py:37-49).  f32: CUDA-core flash kernel (mode a);
 * bf16: tcgen05/TMEM kernel.                                                                   */
int mapdit_cos_attn_fwd(const void* qkv, void* o, float* lse /* nullable: [M, H] log-sum-exp for the backward */,
                        int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream);
/* dqkv[M, 3D] <- d/d(q^, k^, v) from dout[M, D]; delta: [M, H] fp32 scratch (src/layers/attention.py:47 autograd).
 * o may be NULL on the fused bf16 path (head_dim 64, tokens == 256) when `delta` already holds dO.O per (row, head), as the
 * MAPDIT_EPI_STORE_DELTA epilogue of the out-proj dgrad GEMM leaves it */
int mapdit_cos_attn_bwd(const void* qkv, const void* o, const void* dout, const float* lse, void* dqkv, float* delta,
                        int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream);

/* the same followed by the backward of the q/k L2 normalisation (sc [M, 2H] = sqrt(hd)/(||v||+eps) from the forward): dqkv receives
 * d/d(raw q, raw k, v); fused into the dq/dk epilogues on the tcgen05 path */
int mapdit_cos_attn_bwd_qknorm(const void* qkv, const void* o, const void* dout, const float* lse, const float* sc, float eps,
                               void* dqkv, float* delta, int n_samples, int tokens, int heads, int head_dim, int dtype, void* stream);

/* ---- embedders / final layer ------------------------------------------------------------------ */
/* x0 = mp_sum(patchify(x)|1 · Wx^T, pos, .5) (src/dit.py:81-84); optional h = modulate(x0,...).   */
int mapdit_patch_embed(const float* x, const float* wx_eff, const float* pos, void* x0, void* h,
                       const float* shift, const float* scale, const float* gain, int64_t ldmod,
                       int n_samples, int channels, int input_size, int patch, int d, int dtype, void* stream);
/* use_mp_embedding=False (UNPINNED): vanilla-DiT sinusoidal features e[n, :dim/2] = cos(t f), e[n, dim/2:] = sin(t f) */
int mapdit_timestep_sincos(const int64_t* t, float* e, int n, int dim, float max_period, void* stream);
/* use_no_layernorm=False (UNPINNED): h = LayerNorm(x; eps 1e-6, no affine) * (1 + scale) + shift; stats [M, 2] = {mean, rstd} (nullable) */
int mapdit_ln_modulate_fwd(const void* x, void* h, const float* shift, const float* scale, float* stats, int64_t ldmod, int m,
                           int d, int tokens, int dtype, void* stream);
int mapdit_ln_modulate_bwd(const void* dh, const void* x, void* R /* nullable */, const float* stats, const float* scale,
                           float* dshift, float* dscale, int64_t ldmod, int n_samples, int d, int tokens, int accumulate,
                           int dtype, void* stream);
/* e[n, j] = sqrt(2) cos(fl(fl(t*scale_j)+shift_j)) (src/blocks/timestep_embedder.py:18-21) */
int mapdit_fourier(const int64_t* t, const float* scale, const float* shift, float* e, int n, int channels, void* stream);
/* out[n,:] = normalize(table[idx[n],:]) (src/basic/mp_embedding.py:21-24); drop: idx -> null_idx where mask */
int mapdit_embed_rows(const int64_t* idx, const uint8_t* drop_mask, int64_t null_idx, const float* table,
                      float* out, int n, int d, float eps, void* stream);
/* c = mp_sum(a, b, 0.5) (src/dit.py:88) and cs = silu(c)/0.596 in f32 and/or bf16 (nullable) */
int mapdit_cond_combine(const float* a, const float* b, float* c, float* cs_f32, void* cs_bf16, int64_t n, void* stream);
/* s[n] = sigmoid((c[n,:]·W_eff^T)·ref/sqrt(adim)) (src/blocks/final_layer.py:12-22) */
int mapdit_mp_scale(const float* c, const float* w_eff, const float* ref, float* s, int n, int d, int adim, void* stream);
/* out[N, 2C, H, W] = cat(unpatchify(mean*s_mu), unpatchify(sigma*s_sigma)) from lin[M, 2 p^2 C]
 * (src/blocks/final_layer.py:57-59, src/dit.py:95-100) */
int mapdit_final_unpatchify(const void* lin, const float* s_mu, const float* s_sigma, float* out, int n_samples,
                            int channels, int input_size, int patch, int dtype, void* stream);
/* classifier-free guidance combine (src/dit.py:113-118), in place on out[2n, 2C, H, W] */
int mapdit_cfg_combine(float* out, int n_half, int channels, int hw, float cfg_scale, void* stream);

/* ---- diffusion coefficient table (fp32 [8, steps], built on the host from the float64 tables of
 * diffusion/gaussian_diffusion.py:166-201 and cast to fp32 exactly where the reference casts,
 * :870): rows 0 sqrt_alphas_cumprod, 1 sqrt_one_minus_alphas_cumprod, 2 sqrt_recip_alphas_cumprod,
 * 3 sqrt_recipm1_alphas_cumprod, 4 posterior_mean_coef1, 5 posterior_mean_coef2,
 * 6 posterior_log_variance_clipped, 7 log(betas).                                              */
#define MAPDIT_DIFF_ROWS 10 /* rows 8, 9: alphas_cumprod, alphas_cumprod_prev (DDIM) */

/* ---- K5: fused diffusion step (diffusion/gaussian_diffusion.py:285-293,320-323,334-339,410-416)
 * t: int64 [N] respaced step index per sample.  sample may alias x.  pred_xstart nullable.     */
int mapdit_diffusion_step(const float* model_out, const float* x, const float* noise, const int64_t* t,
                          const float* tables, int steps, float* sample, float* pred_xstart, int n_samples,
                          int channels, int hw, int clip_denoised, void* stream);
/* DDIM step (diffusion/gaussian_diffusion.py:513-560); noise may be null when eta == 0 */
int mapdit_ddim_step(const float* model_out, const float* x, const float* noise, const int64_t* t, const float* tables,
                     int steps, float* sample, float* pred_xstart, int n_samples, int channels, int hw, int clip_denoised,
                     float eta, void* stream);
/* ---- K6: fused q_sample + loss (diffusion/gaussian_diffusion.py:215-230,682-713,747-783;
 * diffusion/diffusion_utils.py:10-36,62-88).                                                    */
int mapdit_q_sample(const float* x0, const float* noise, const int64_t* t, const float* tables, int steps,
                    float* x_t, int n_samples, int chw, void* stream);
/* loss[n] = mse[n] + vb[n] (outputs nullable).  grad_out (nullable) [N, 2C, H, W]: eps channels receive
 * gs_eps[n] * d mse[n]/d eps, variance channels gs_var[n] * d vb[n]/d v (gs_* nullable = 1).       */
int mapdit_loss_fwd_bwd(const float* model_out, const float* x0, const float* x_t, const float* noise,
                        const int64_t* t, const float* tables, int steps, float* loss, float* mse, float* vb,
                        float* grad_out, const float* gs_eps, const float* gs_var, int n_samples, int channels,
                        int hw, void* stream);
/* p_mean_variance as separate tensors (diffusion/gaussian_diffusion.py:254-332) */
int mapdit_p_mean_variance(const float* model_out, const float* x, const int64_t* t, const float* tables, int steps,
                           float* mean, float* variance, float* log_variance, float* pred_xstart, int n_samples,
                           int channels, int hw, int clip_denoised, void* stream);
/* posterior mean c1*x0 + c2*x_t (:238-241) and sample = mean + [t!=0] exp(.5 logvar) noise (:410-416) */
int mapdit_posterior_mean(const float* x0, const float* x, const int64_t* t, const float* tables, int steps,
                          float* mean, int n_samples, int chw, void* stream);
int mapdit_noise_add(const float* mean, const float* log_variance, const float* noise, const int64_t* t,
                     float* sample, int n_samples, int chw, void* stream);

/* ---- backward of the elementwise ops (closed forms: SURVEY.md §A.3) ------------------------------ */
/* mp residual: R <- 0.7/den R (in place); dy = 0.3/den gate R; dgate[n, :] = sum_t 0.3/den y R */
int mapdit_resid_bwd(void* R, const void* y, void* dy, const float* gate, float* dgate, int64_t ldmod, int n_samples,
                     int d, int tokens, int dtype, void* stream);
/* modulate: R (+)= dh (1-g)/den scale; dscale, dshift per sample; dg partial sums (one float per CTA,
 * mapdit_modulate_bwd_partials() of them, finished by mapdit_sum_partials) */
int mapdit_modulate_bwd(const void* dh, const void* x, void* R, const float* shift, const float* scale, const float* gain,
                        float* dshift, float* dscale, float* dg_partial, int64_t ldmod, int n_samples, int d, int tokens,
                        int accumulate, int dtype, void* stream);
/* mapdit_modulate_bwd followed, in the same pass, by mapdit_resid_bwd of the residual that precedes it in the block
 * (y, gate, dgate: that residual's branch output and gate; dy: its output gradient): 6 passes over [M, D] instead of 8 */
int mapdit_modulate_resid_bwd(const void* dh, const void* x, void* R, const float* shift, const float* scale, const float* gain,
                              float* dshift, float* dscale, float* dg_partial, const void* y, void* dy, const float* gate,
                              float* dgate, int64_t ldmod, int n_samples, int d, int tokens, int accumulate, int dtype, void* stream);
int mapdit_modulate_bwd_partials(int n_samples, int d);
int mapdit_sum_partials(const float* partials, int n, float* out, int accumulate, void* stream);
int mapdit_mp_silu_bwd(const void* du, const void* z, void* dz, int64_t n, int dtype, void* stream);
/* q/k normalisation: forward variant that records sc = sqrt(hd)/(||v||+eps) [M, 2H], and its backward (in place on dqkv) */
int mapdit_qk_normalize_save(void* qkv, float* sc, int m, int d, int head_dim, float eps, int dtype, void* stream);
int mapdit_qk_norm_bwd(void* dqkv, const void* qkv, const float* sc, int m, int d, int head_dim, float eps, int dtype,
                       void* stream);
int mapdit_final_bwd(const float* dout, const void* lin, const float* s_mu, const float* s_sigma, void* dlin, float* ds_mu,
                     float* ds_sigma, int n_samples, int channels, int input_size, int patch, int dtype, void* stream);
int mapdit_mp_scale_from_lin(const float* l, const float* ref, float* s, int n, int adim, void* stream);
int mapdit_mp_scale_bwd(const float* ds, const float* s, const float* l, const float* ref, float* dl, float* dref, int n,
                        int adim, int accumulate, void* stream);
int mapdit_cond_combine_bwd(const float* c, const float* dc, const float* dcs, float* dab, int64_t n, void* stream);
int mapdit_embed_rows_bwd(const int64_t* idx, const uint8_t* drop_mask, int64_t null_idx, const float* table,
                          const float* g, float* dtable, int n, int d, float eps, void* stream);
int mapdit_patchify(const float* x, float* P, int n_samples, int channels, int input_size, int patch, void* stream);
int mapdit_axpby(const float* x, float* y, float a, int accumulate, int64_t n, void* stream);
/* rotation modulation (UNPINNED: no reference code): h = R(rot[n,:] * gain) x (* scale), pairs (2i, 2i+1);
 * backward: R (+)= R^T (dh*scale), dscale, drot, dgain partials (mapdit_modulate_bwd_partials() of them) */
int mapdit_rotmod_fwd(const void* x, void* h, const float* rot, const float* scale, const float* gain, int64_t ldmod, int m,
                      int d, int tokens, int dtype, void* stream);
int mapdit_rotmod_bwd(const void* dh, const void* x, void* R, const float* rot, const float* scale, const float* gain,
                      float* drot, float* dscale, float* dg_partial, int64_t ldmod, int n_samples, int d, int tokens,
                      int accumulate, int dtype, void* stream);
int mapdit_rotmod_bwd_partials(int n_samples, int d); /* number of dgain partials mapdit_rotmod_bwd writes */
/* the same followed, in the same pass, by the backward of the residual that precedes the modulation in the block schedule
 * (mapdit_resid_bwd on the updated R with branch output y / gate): R'' = .7/den R', dy = .3/den gate R', dgate = sum_t .3/den y R' */
int mapdit_rotmod_resid_bwd(const void* dh, const void* x, void* R, const float* rot, const float* scale, const float* gain,
                            float* drot, float* dscale, float* dg_partial, const void* y, void* dy, const float* gate,
                            float* dgate, int64_t ldmod, int n_samples, int d, int tokens, int accumulate, int dtype,
                            void* stream);
/* per-sample rotation table for MAPDIT_EPI_RESID_ROT: cs[n, 2i] = cos(rot[n, i] * gain), cs[n, 2i+1] = sin(rot[n, i] * gain),
 * i < d/2; the second {rot2, gain2, cs2} triple is optional (both branches of a block in one launch) */
int mapdit_rot_table(const float* rot, const float* gain, float* cs, const float* rot2, const float* gain2, float* cs2,
                     int64_t ldmod, int64_t ldcs, int n_samples, int d, void* stream);
/* x_embedder weight gradient dW[D, p*p*C+1] = scale * R[M, D]^T · (patchify(x)|1), patches gathered on the fly (src/dit.py:81-84) */
int mapdit_patch_embed_wgrad(const void* R, const float* x, float* dW, int n_samples, int channels, int input_size,
                             int patch, int d, float scale, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAPDIT_H_ */
