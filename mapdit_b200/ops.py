"""Torch-tensor wrappers over the C ABI (device pointers + current CUDA stream).

Each wrapper validates device/dtype/contiguity and then calls the kernel; none of them has a
PyTorch fallback.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BF16, F32, GemmArgs, check, lib

EPS = 1e-4  # src/utils.py:19 normalize eps


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, "mapdit_b200 ops need CUDA tensors (there is no CPU path)"
    return C.c_void_p(t.data_ptr())


def weight_norm_fwd(w, force=False, eff_f32=None, eff_bf16=None, eff_bf16_t=None, inv_norm=None, ld_t=0):
    assert w.dtype == torch.float32 and w.is_contiguous() and w.dim() == 2
    rows, cols = w.shape
    for e in (eff_f32, eff_bf16):
        assert e is None or (e.is_contiguous() and e.numel() == w.numel())
    # force: True/1 = forced write-back + normalise, False/0 = normalise only, -1 = plain copy (no weight normalisation)
    check(lib().mapdit_weight_norm_fwd(_ptr(w), rows, cols, EPS, int(force), _ptr(eff_f32), _ptr(eff_bf16),
                                       _ptr(eff_bf16_t), ld_t, _ptr(inv_norm), _stream()), "weight_norm_fwd")


class WeightNormBatch:
    """descriptor table for mapdit_weight_norm_fwd_multi: `items` = [(w, eff_f32|None, eff_bf16|None, eff_bf16_t|None, ld_t)]"""

    def __init__(self, items, device):
        rows_tab, g0, t0, r0 = [], 0, 0, 0
        for w, e32, e16, e16t, ld_t in items:
            rows, cols = w.shape
            assert w.dtype == torch.float32 and w.is_contiguous() and cols % 4 == 0 and w.data_ptr() % 16 == 0
            rows_tab.append([w.data_ptr(), e32.data_ptr() if e32 is not None else 0, e16.data_ptr() if e16 is not None else 0,
                             e16t.data_ptr() if e16t is not None else 0, ld_t if ld_t else rows, rows, cols, g0, t0, r0])
            g0 += (rows + 7) // 8
            t0 += ((rows + 63) // 64) * ((cols + 63) // 64)
            r0 += rows
        self.table = torch.tensor(rows_tab, dtype=torch.int64).to(device)
        self.n, self.groups, self.tiles = len(rows_tab), g0, t0
        self.scratch = torch.empty(r0, 2, device=device, dtype=torch.float32)
        self.signature = tuple(r[0] for r in rows_tab)

    def run(self, force):
        check(lib().mapdit_weight_norm_fwd_multi(_ptr(self.table), self.n, self.groups, self.tiles, EPS, int(force), _ptr(self.scratch),
                                                 _stream()), "weight_norm_fwd_multi")


def weight_norm_bwd(v, g_eff, grad_v, accumulate=False):
    rows, cols = v.shape
    assert v.is_contiguous() and g_eff.is_contiguous() and grad_v.is_contiguous()
    check(lib().mapdit_weight_norm_bwd(_ptr(v), _ptr(g_eff), _ptr(grad_v), rows, cols, EPS, int(accumulate), _stream()),
          "weight_norm_bwd")


class WeightNormBwdBatch:
    """descriptor table for mapdit_weight_norm_bwd_multi: `items` = [(v, g)], g = d(effective weight) on entry, d(v) on return"""

    def __init__(self, items, device):
        rows_tab, g0 = [], 0
        for v, g in items:
            rows, cols = v.shape
            assert v.dtype == g.dtype == torch.float32 and v.is_contiguous() and g.is_contiguous() and g.shape == v.shape
            assert cols % 4 == 0 and v.data_ptr() % 16 == 0 and g.data_ptr() % 16 == 0
            rows_tab.append([v.data_ptr(), g.data_ptr(), rows, cols, g0])
            g0 += (rows + 7) // 8
        self.table = torch.tensor(rows_tab, dtype=torch.int64).to(device)
        self.n, self.groups = len(rows_tab), g0

    @staticmethod
    def supports(v, g):
        return v.dim() == 2 and v.shape[1] % 4 == 0 and v.is_contiguous() and g.is_contiguous() and v.data_ptr() % 16 == 0 and g.data_ptr() % 16 == 0

    def run(self):
        check(lib().mapdit_weight_norm_bwd_multi(_ptr(self.table), self.n, self.groups, EPS, _stream()), "weight_norm_bwd_multi")


def gemm_f32(a, b, out=None, trans_a=False, trans_b=False, accumulate=False):
    """out[m,n] (+)= sum_k A(m,k) B(n,k).  a is [m,k] (or [k,m] with trans_a), b is [n,k] (or [k,n] with
    trans_b); 2-D fp32 with arbitrary strides."""
    assert a.dtype == torch.float32 and b.dtype == torch.float32 and a.dim() == 2 and b.dim() == 2
    if trans_a:
        a = a.t()
    if trans_b:
        b = b.t()
    m, k = a.shape
    n, k2 = b.shape
    assert k == k2, (a.shape, b.shape)
    if out is None:
        out = torch.empty(m, n, device=a.device, dtype=torch.float32)
    assert out.stride(1) == 1
    check(lib().mapdit_gemm_f32(_ptr(a), a.stride(0), a.stride(1), _ptr(b), b.stride(0), b.stride(1), _ptr(out),
                                out.stride(0), m, n, k, int(accumulate), _stream()), "gemm_f32")
    return out


def gemm_bf16(a, b, out, epilogue=_lib.EPI_STORE, out2=None, resid=None, gate=None, shift=None, scale=None, gain=None,
              ldmod=0, tokens=1, head_dim=0, qk_cols=0, aux=None, ldrot=0):
    """out = epilogue(a[M,K] @ b[N,K]^T) on the tcgen05 path (bf16 operands, fp32 accumulate)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    assert a.stride(1) == 1 and b.stride(1) == 1 and out.stride(1) == 1
    m, k = a.shape
    n, k2 = b.shape
    assert k == k2 and out.shape[0] == m and out.shape[1] == n
    args = GemmArgs(a=a.data_ptr(), b=b.data_ptr(), out=out.data_ptr(),
                    out2=out2.data_ptr() if out2 is not None else None,
                    resid=resid.data_ptr() if resid is not None else None,
                    gate=gate.data_ptr() if gate is not None else None,
                    shift=shift.data_ptr() if shift is not None else None,
                    scale=scale.data_ptr() if scale is not None else None,
                    gain=gain.data_ptr() if gain is not None else None,
                    aux=aux.data_ptr() if aux is not None else None,
                    lda=a.stride(0), ldb=b.stride(0), ldo=out.stride(0), ldmod=ldmod, m=m, n=n, k=k, tokens=tokens,
                    head_dim=head_dim, qk_cols=qk_cols, epilogue=epilogue, out_dtype=_dt(out), eps=EPS, ldrot=ldrot)
    check(lib().mapdit_gemm_bf16(C.byref(args), _stream()), "gemm_bf16")
    return out


def gemm_bf16_tn(dy, x, out):
    """out[N_out, K_in] (fp32) = dy[M, N_out]^T @ x[M, K_in] on the tcgen05 path (weight gradient)."""
    assert dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and out.dtype == torch.float32
    assert dy.stride(1) == 1 and x.stride(1) == 1 and out.stride(1) == 1 and dy.shape[0] == x.shape[0]
    m, n_out = dy.shape
    k_in = x.shape[1]
    assert out.shape == (n_out, k_in)
    check(lib().mapdit_gemm_bf16_tn(_ptr(dy), dy.stride(0), _ptr(x), x.stride(0), _ptr(out), out.stride(0), m, n_out, k_in, _stream()),
          "gemm_bf16_tn")
    return out


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """in-place Adam on flat fp32 tensors; `step` is the 1-based step count (bias corrections as torch.optim.Adam)."""
    n = p.numel()
    assert p.is_contiguous() and g.is_contiguous() and m.is_contiguous() and v.is_contiguous()
    check(lib().mapdit_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), n, float(lr), float(beta1), float(beta2), float(eps),
                                 1.0 - beta1 ** step, 1.0 - beta2 ** step, float(grad_scale), _stream()), "adam_step")


def adam_step_g16(p, g16, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """adam_step with the gradient span in bf16 (the all-reduced copy of the data-parallel path)"""
    n = p.numel()
    assert g16.dtype == torch.bfloat16 and g16.numel() == n and p.is_contiguous() and g16.is_contiguous()
    check(lib().mapdit_adam_step_g16(_ptr(p), _ptr(g16), _ptr(m), _ptr(v), n, float(lr), float(beta1), float(beta2), float(eps),
                                     1.0 - beta1 ** step, 1.0 - beta2 ** step, float(grad_scale), _stream()), "adam_step_g16")


def multi_lerp(chunk_table, n_chunks, weight):
    check(lib().mapdit_multi_lerp(_ptr(chunk_table), n_chunks, float(weight), _stream()), "multi_lerp")


def latent_sample(means, stds, idx, eps, ch_mean, ch_std, out):
    n, c = out.shape[0], out.shape[1]
    hw = out[0, 0].numel()
    check(lib().mapdit_latent_sample(_ptr(means), _ptr(stds), _ptr(idx), _ptr(eps), _ptr(ch_mean), _ptr(ch_std), _ptr(out), n, c, hw,
                                     _stream()), "latent_sample")


def set_variant(flags: int) -> int:
    """select the README --use-* "off" variants for subsequent launches of this thread; returns the previous word"""
    return lib().mapdit_set_variant(int(flags))


def timestep_sincos(t, e, max_period=10000.0):
    check(lib().mapdit_timestep_sincos(_ptr(t), _ptr(e), t.shape[0], e.shape[1], float(max_period), _stream()), "timestep_sincos")


def ln_modulate(x, h, shift, scale, stats, ldmod, tokens):
    m, d = x.shape
    check(lib().mapdit_ln_modulate_fwd(_ptr(x), _ptr(h), _ptr(shift), _ptr(scale), _ptr(stats), ldmod, m, d, tokens, _dt(x), _stream()),
          "ln_modulate_fwd")


def ln_modulate_bwd(dh, x, R, stats, scale, dshift, dscale, ldmod, n_samples, tokens, accumulate):
    d = dh.shape[1]
    check(lib().mapdit_ln_modulate_bwd(_ptr(dh), _ptr(x), _ptr(R), _ptr(stats), _ptr(scale), _ptr(dshift), _ptr(dscale), ldmod,
                                       n_samples, d, tokens, int(accumulate), _dt(dh), _stream()), "ln_modulate_bwd")


def modulate(x, h, shift, scale, gain, ldmod, tokens):
    m, d = x.shape
    check(lib().mapdit_modulate_fwd(_ptr(x), _ptr(h), _ptr(shift), _ptr(scale), _ptr(gain), ldmod, m, d, tokens, _dt(x),
                                    _stream()), "modulate_fwd")


def resid(x, y, xout, gate, ldmod, tokens):
    m, d = x.shape
    check(lib().mapdit_resid_fwd(_ptr(x), _ptr(y), _ptr(xout), _ptr(gate), ldmod, m, d, tokens, _dt(x), _stream()), "resid_fwd")


def mp_silu(x, y):
    check(lib().mapdit_mp_silu_fwd(_ptr(x), _ptr(y), x.numel(), _dt(x), _dt(y), _stream()), "mp_silu_fwd")


def qk_normalize(qkv, d, head_dim):
    check(lib().mapdit_qk_normalize(_ptr(qkv), qkv.shape[0], d, head_dim, EPS, _dt(qkv), _stream()), "qk_normalize")


def cast(src, dst):
    check(lib().mapdit_cast(_ptr(src), _ptr(dst), src.numel(), _dt(src), _dt(dst), _stream()), "cast")


def cast_2d(src, dst):
    """dst[r, c] = src[r, c] for 2-D views with unit column stride (column slices of wider matrices)"""
    assert src.dim() == 2 and src.shape == dst.shape and src.stride(1) == 1 and dst.stride(1) == 1
    check(lib().mapdit_cast_2d(_ptr(src), src.stride(0), _ptr(dst), dst.stride(0), src.shape[0], src.shape[1], _dt(src), _dt(dst),
                               _stream()), "cast_2d")


def cos_attn(qkv, o, n_samples, tokens, heads, head_dim, lse=None):
    check(lib().mapdit_cos_attn_fwd(_ptr(qkv), _ptr(o), _ptr(lse), n_samples, tokens, heads, head_dim, _dt(qkv), _stream()),
          "cos_attn_fwd")


def cos_attn_bwd(qkv, o, dout, lse, dqkv, delta, n_samples, tokens, heads, head_dim):
    check(lib().mapdit_cos_attn_bwd(_ptr(qkv), _ptr(o), _ptr(dout), _ptr(lse), _ptr(dqkv), _ptr(delta), n_samples, tokens, heads,
                                    head_dim, _dt(qkv), _stream()), "cos_attn_bwd")


def cos_attn_bwd_qknorm(qkv, o, dout, lse, sc, dqkv, delta, n_samples, tokens, heads, head_dim):
    """attention backward + q/k normalisation backward: dqkv = d/d(raw q, raw k, v)"""
    check(lib().mapdit_cos_attn_bwd_qknorm(_ptr(qkv), _ptr(o), _ptr(dout), _ptr(lse), _ptr(sc), EPS, _ptr(dqkv), _ptr(delta), n_samples,
                                           tokens, heads, head_dim, _dt(qkv), _stream()), "cos_attn_bwd_qknorm")


def resid_bwd(R, y, dy, gate, dgate, ldmod, n_samples, tokens):
    d = R.shape[1]
    check(lib().mapdit_resid_bwd(_ptr(R), _ptr(y), _ptr(dy), _ptr(gate), _ptr(dgate), ldmod, n_samples, d, tokens, _dt(R), _stream()),
          "resid_bwd")


def modulate_bwd_partials(n_samples, d):
    return lib().mapdit_modulate_bwd_partials(n_samples, d)


def rotmod_bwd_partials(n_samples, d):
    return lib().mapdit_rotmod_bwd_partials(n_samples, d)


def modulate_bwd(dh, x, R, shift, scale, gain, dshift, dscale, dg_partial, ldmod, n_samples, tokens, accumulate):
    d = dh.shape[1]
    check(lib().mapdit_modulate_bwd(_ptr(dh), _ptr(x), _ptr(R), _ptr(shift), _ptr(scale), _ptr(gain), _ptr(dshift), _ptr(dscale),
                                    _ptr(dg_partial), ldmod, n_samples, d, tokens, int(accumulate), _dt(dh), _stream()), "modulate_bwd")


def modulate_resid_bwd(dh, x, R, shift, scale, gain, dshift, dscale, dg_partial, y, dy, gate, dgate, ldmod, n_samples, tokens, accumulate):
    d = dh.shape[1]
    check(lib().mapdit_modulate_resid_bwd(_ptr(dh), _ptr(x), _ptr(R), _ptr(shift), _ptr(scale), _ptr(gain), _ptr(dshift), _ptr(dscale),
                                          _ptr(dg_partial), _ptr(y), _ptr(dy), _ptr(gate), _ptr(dgate), ldmod, n_samples, d, tokens,
                                          int(accumulate), _dt(dh), _stream()), "modulate_resid_bwd")


def sum_partials(partials, n, out, accumulate=False):
    check(lib().mapdit_sum_partials(_ptr(partials), n, _ptr(out), int(accumulate), _stream()), "sum_partials")


def mp_silu_bwd(du, z, dz):
    check(lib().mapdit_mp_silu_bwd(_ptr(du), _ptr(z), _ptr(dz), du.numel(), _dt(du), _stream()), "mp_silu_bwd")


def qk_normalize_save(qkv, sc, d, head_dim):
    check(lib().mapdit_qk_normalize_save(_ptr(qkv), _ptr(sc), qkv.shape[0], d, head_dim, EPS, _dt(qkv), _stream()), "qk_normalize_save")


def qk_norm_bwd(dqkv, qkv, sc, d, head_dim):
    check(lib().mapdit_qk_norm_bwd(_ptr(dqkv), _ptr(qkv), _ptr(sc), qkv.shape[0], d, head_dim, EPS, _dt(qkv), _stream()), "qk_norm_bwd")


def final_bwd(dout, lin, s_mu, s_sigma, dlin, ds_mu, ds_sigma, patch):
    n, c2, s, _ = dout.shape
    check(lib().mapdit_final_bwd(_ptr(dout), _ptr(lin), _ptr(s_mu), _ptr(s_sigma), _ptr(dlin), _ptr(ds_mu), _ptr(ds_sigma), n,
                                 c2 // 2, s, patch, _dt(lin), _stream()), "final_bwd")


def mp_scale_from_lin(l, ref, s):
    check(lib().mapdit_mp_scale_from_lin(_ptr(l), _ptr(ref), _ptr(s), l.shape[0], l.shape[1], _stream()), "mp_scale_from_lin")


def mp_scale_bwd(ds, s, l, ref, dl, dref, accumulate=False):
    check(lib().mapdit_mp_scale_bwd(_ptr(ds), _ptr(s), _ptr(l), _ptr(ref), _ptr(dl), _ptr(dref), l.shape[0], l.shape[1],
                                    int(accumulate), _stream()), "mp_scale_bwd")


def cond_combine_bwd(c, dc, dcs, dab):
    check(lib().mapdit_cond_combine_bwd(_ptr(c), _ptr(dc), _ptr(dcs), _ptr(dab), c.numel(), _stream()), "cond_combine_bwd")


def embed_rows_bwd(idx, drop_mask, null_idx, table, g, dtable):
    check(lib().mapdit_embed_rows_bwd(_ptr(idx), _ptr(drop_mask), null_idx, _ptr(table), _ptr(g), _ptr(dtable), idx.shape[0],
                                      table.shape[1], EPS, _stream()), "embed_rows_bwd")


def patchify(x, P, patch):
    n, c, s, _ = x.shape
    check(lib().mapdit_patchify(_ptr(x), _ptr(P), n, c, s, patch, _stream()), "patchify")


def patch_embed_wgrad(R, x, dW, patch, scale):
    n, c, s, _ = x.shape
    check(lib().mapdit_patch_embed_wgrad(_ptr(R), _ptr(x), _ptr(dW), n, c, s, patch, R.shape[1], float(scale), _dt(R), _stream()),
          "patch_embed_wgrad")


def rotmod(x, h, rot, scale, gain, ldmod, tokens):
    m, d = x.shape
    check(lib().mapdit_rotmod_fwd(_ptr(x), _ptr(h), _ptr(rot), _ptr(scale), _ptr(gain), ldmod, m, d, tokens, _dt(x), _stream()), "rotmod_fwd")


def rotmod_bwd(dh, x, R, rot, scale, gain, drot, dscale, dg_partial, ldmod, n_samples, tokens, accumulate):
    d = dh.shape[1]
    check(lib().mapdit_rotmod_bwd(_ptr(dh), _ptr(x), _ptr(R), _ptr(rot), _ptr(scale), _ptr(gain), _ptr(drot), _ptr(dscale),
                                  _ptr(dg_partial), ldmod, n_samples, d, tokens, int(accumulate), _dt(dh), _stream()), "rotmod_bwd")


def rotmod_resid_bwd(dh, x, R, rot, scale, gain, drot, dscale, dg_partial, y, dy, gate, dgate, ldmod, n_samples, tokens, accumulate):
    d = dh.shape[1]
    check(lib().mapdit_rotmod_resid_bwd(_ptr(dh), _ptr(x), _ptr(R), _ptr(rot), _ptr(scale), _ptr(gain), _ptr(drot), _ptr(dscale),
                                        _ptr(dg_partial), _ptr(y), _ptr(dy), _ptr(gate), _ptr(dgate), ldmod, n_samples, d, tokens,
                                        int(accumulate), _dt(dh), _stream()), "rotmod_resid_bwd")


def rot_table(rot, gain, cs, ldmod, d, rot2=None, gain2=None, cs2=None):
    """cs[n, 2i | 2i+1] = cos | sin(rot[n, i] * gain): the per-sample table the EPI_RESID_ROT GEMM epilogue reads"""
    check(lib().mapdit_rot_table(_ptr(rot), _ptr(gain), _ptr(cs), _ptr(rot2), _ptr(gain2), _ptr(cs2), ldmod, cs.stride(0),
                                 cs.shape[0], d, _stream()), "rot_table")


def axpby(x, y, a, accumulate=False):
    check(lib().mapdit_axpby(_ptr(x), _ptr(y), float(a), int(accumulate), x.numel(), _stream()), "axpby")


def patch_embed(x, wx_eff, pos, x0, h, shift, scale, gain, ldmod, patch):
    n, c, s, _ = x.shape
    d = wx_eff.shape[0]
    check(lib().mapdit_patch_embed(_ptr(x), _ptr(wx_eff), _ptr(pos), _ptr(x0), _ptr(h), _ptr(shift), _ptr(scale), _ptr(gain),
                                   ldmod, n, c, s, patch, d, _dt(x0), _stream()), "patch_embed")


def fourier(t, scale, shift, e):
    check(lib().mapdit_fourier(_ptr(t), _ptr(scale), _ptr(shift), _ptr(e), t.shape[0], scale.shape[0], _stream()), "fourier")


def embed_rows(idx, drop_mask, null_idx, table, out):
    check(lib().mapdit_embed_rows(_ptr(idx), _ptr(drop_mask), null_idx, _ptr(table), _ptr(out), idx.shape[0], table.shape[1],
                                  EPS, _stream()), "embed_rows")


def cond_combine(a, b, c, cs_f32, cs_bf16):
    check(lib().mapdit_cond_combine(_ptr(a), _ptr(b), _ptr(c), _ptr(cs_f32), _ptr(cs_bf16), a.numel(), _stream()), "cond_combine")


def mp_scale(c, w_eff, ref, s):
    check(lib().mapdit_mp_scale(_ptr(c), _ptr(w_eff), _ptr(ref), _ptr(s), c.shape[0], c.shape[1], ref.shape[0], _stream()), "mp_scale")


def final_unpatchify(lin, s_mu, s_sigma, out, patch):
    n, c2, s, _ = out.shape
    check(lib().mapdit_final_unpatchify(_ptr(lin), _ptr(s_mu), _ptr(s_sigma), _ptr(out), n, c2 // 2, s, patch, _dt(lin), _stream()),
          "final_unpatchify")


def cfg_combine(out, in_channels, cfg_scale):
    n2, c2, h, w = out.shape
    assert c2 == 2 * in_channels and out.is_contiguous()
    check(lib().mapdit_cfg_combine(_ptr(out), n2 // 2, in_channels, h * w, float(cfg_scale), _stream()), "cfg_combine")


def diffusion_step(model_out, x, noise, t, tables, sample, pred_xstart, clip_denoised):
    n, c, h, w = x.shape
    check(lib().mapdit_diffusion_step(_ptr(model_out), _ptr(x), _ptr(noise), _ptr(t), _ptr(tables), tables.shape[1], _ptr(sample),
                                      _ptr(pred_xstart), n, c, h * w, int(clip_denoised), _stream()), "diffusion_step")


def ddim_step(model_out, x, noise, t, tables, sample, pred_xstart, clip_denoised, eta):
    n, c, h, w = x.shape
    check(lib().mapdit_ddim_step(_ptr(model_out), _ptr(x), _ptr(noise), _ptr(t), _ptr(tables), tables.shape[1], _ptr(sample),
                                 _ptr(pred_xstart), n, c, h * w, int(clip_denoised), float(eta), _stream()), "ddim_step")


def q_sample(x0, noise, t, tables, x_t):
    n = x0.shape[0]
    check(lib().mapdit_q_sample(_ptr(x0), _ptr(noise), _ptr(t), _ptr(tables), tables.shape[1], _ptr(x_t), n, x0[0].numel(), _stream()),
          "q_sample")


def loss_fwd_bwd(model_out, x0, x_t, noise, t, tables, loss, mse, vb, grad_out, gs_eps, gs_var):
    n, c, h, w = x0.shape
    check(lib().mapdit_loss_fwd_bwd(_ptr(model_out), _ptr(x0), _ptr(x_t), _ptr(noise), _ptr(t), _ptr(tables), tables.shape[1],
                                    _ptr(loss), _ptr(mse), _ptr(vb), _ptr(grad_out), _ptr(gs_eps), _ptr(gs_var), n, c, h * w,
                                    _stream()), "loss_fwd_bwd")


def p_mean_variance(model_out, x, t, tables, mean, var, logvar, x0, clip_denoised):
    n, c, h, w = x.shape
    check(lib().mapdit_p_mean_variance(_ptr(model_out), _ptr(x), _ptr(t), _ptr(tables), tables.shape[1], _ptr(mean), _ptr(var),
                                       _ptr(logvar), _ptr(x0), n, c, h * w, int(clip_denoised), _stream()), "p_mean_variance")


def posterior_mean(x0, x, t, tables, mean):
    check(lib().mapdit_posterior_mean(_ptr(x0), _ptr(x), _ptr(t), _ptr(tables), tables.shape[1], _ptr(mean), x.shape[0],
                                      x[0].numel(), _stream()), "posterior_mean")


def noise_add(mean, logvar, noise, t, sample):
    check(lib().mapdit_noise_add(_ptr(mean), _ptr(logvar), _ptr(noise), _ptr(t), _ptr(sample), mean.shape[0], mean[0].numel(),
                                 _stream()), "noise_add")
