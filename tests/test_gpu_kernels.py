"""Per-kernel parity tests (run on the B200 with -m gpu).  Every call goes through the C ABI
(mapdit_b200.ops -> ctypes -> libmapdit.so); the checker is the CPU oracle or the same formula
evaluated with PyTorch in fp32/fp64."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import mapdit_oracle as O

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5  # north_star mode (a): rel-L2 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def rnd(*shape, seed=0, dev="cuda:0"):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dev)


def test_weight_norm_fwd_bwd(dev):
    from mapdit_b200 import ops
    for rows, cols in [(8, 384), (1152, 384), (97, 17), (384, 1536), (33, 4608)]:
        w = rnd(rows, cols, seed=rows)
        ref_forced = O.normalize(w.cpu())
        ref_eff = O.normalize(ref_forced) / math.sqrt(cols)
        wf = w.clone()
        e32 = torch.empty_like(w)
        e16 = torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)
        e16t = torch.empty(cols, rows, device=dev, dtype=torch.bfloat16)
        inv = torch.empty(rows, device=dev)
        ops.weight_norm_fwd(wf, force=True, eff_f32=e32, eff_bf16=e16, eff_bf16_t=e16t, inv_norm=inv)
        assert rel_l2(wf.cpu(), ref_forced) < 1e-6
        assert rel_l2(e32.cpu(), ref_eff) < 1e-6
        assert rel_l2(e16.float().cpu(), ref_eff) < 5e-3
        assert torch.equal(e16t, e16.t().contiguous())
        # eval flavour: no write-back
        w2 = w.clone()
        ops.weight_norm_fwd(w2, force=False, eff_f32=e32)
        assert torch.equal(w2, w)
        assert rel_l2(e32.cpu(), O.normalize(w.cpu()) / math.sqrt(cols)) < 1e-6
        # backward vs autograd (fp64)
        v = w.double().cpu().requires_grad_(True)
        g = rnd(rows, cols, seed=rows + 1)
        eff = v / (v.norm(dim=1, keepdim=True) + 1e-4)
        (eff * g.double().cpu()).sum().backward()
        gv = torch.empty_like(w)
        ops.weight_norm_bwd(w, g, gv)
        assert rel_l2(gv.cpu(), v.grad) < 1e-5


@pytest.mark.parametrize("m,n,k", [(64, 384, 384), (130, 70, 17), (512, 1152, 384), (8, 8, 384), (257, 129, 100), (33, 1536, 2304)])
def test_gemm_f32(dev, m, n, k):
    from mapdit_b200 import ops
    a, b = rnd(m, k, seed=1), rnd(n, k, seed=2)
    ref = (a.double() @ b.double().t())
    assert rel_l2(ops.gemm_f32(a, b), ref) < 2e-6
    # dgrad form: A[m,k] @ B[k,n]  (B given as [k, n], so trans_b)
    bt = b.t().contiguous()
    assert rel_l2(ops.gemm_f32(a, bt, trans_b=True), ref) < 2e-6
    # wgrad form: A^T given as [k, m]
    at = a.t().contiguous()
    assert rel_l2(ops.gemm_f32(at, b, trans_a=True), ref) < 2e-6
    out = torch.ones(m, n, device=dev)
    ops.gemm_f32(a, b, out=out, accumulate=True)
    assert rel_l2(out, ref + 1) < 2e-6


def test_elementwise_modulate_resid_silu(dev):
    from mapdit_b200 import ops
    N, T, D = 3, 64, 384
    x = rnd(N * T, D, seed=3)
    mods = rnd(N, 6 * D, seed=4)
    gain = torch.tensor(0.37, device=dev)
    h = torch.empty_like(x)
    ops.modulate(x, h, mods[:, D:], mods[:, 2 * D:], gain, mods.shape[1], T)
    ref = O.modulate(x.view(N, T, D).cpu(), mods[:, D:2 * D].cpu(), mods[:, 2 * D:3 * D].cpu(), gain.cpu()).reshape(N * T, D)
    assert rel_l2(h.cpu(), ref) < 1e-6
    gain0 = torch.tensor(0.0, device=dev)
    ops.modulate(x, h, mods[:, D:], mods[:, 2 * D:], gain0, mods.shape[1], T)
    ref0 = O.modulate(x.view(N, T, D).cpu(), mods[:, D:2 * D].cpu(), mods[:, 2 * D:3 * D].cpu(), 0.0).reshape(N * T, D)
    assert rel_l2(h.cpu(), ref0) < 1e-6
    y = rnd(N * T, D, seed=5)
    xo = torch.empty_like(x)
    ops.resid(x, y, xo, mods[:, 3 * D:], mods.shape[1], T)
    refr = O.mp_sum(x.view(N, T, D).cpu(), mods[:, 3 * D:4 * D].cpu().unsqueeze(1) * y.view(N, T, D).cpu(), 0.3).reshape(N * T, D)
    assert rel_l2(xo.cpu(), refr) < 1e-6
    s = torch.empty_like(x)
    ops.mp_silu(x * 3, s)
    assert rel_l2(s.cpu(), O.mp_silu(x.cpu() * 3)) < 1e-6
    # bf16 flavours
    xb = x.bfloat16()
    hb = torch.empty_like(xb)
    ops.modulate(xb, hb, mods[:, D:], mods[:, 2 * D:], gain, mods.shape[1], T)
    refb = O.modulate(xb.float().view(N, T, D).cpu(), mods[:, D:2 * D].cpu(), mods[:, 2 * D:3 * D].cpu(), gain.cpu()).reshape(N * T, D)
    assert rel_l2(hb.float().cpu(), refb) < 4e-3


@pytest.mark.parametrize("N,T,H,hd", [(2, 64, 6, 64), (1, 256, 4, 64), (2, 16, 4, 64), (1, 200, 2, 72), (1, 1024, 2, 64)])
def test_attention_f32(dev, N, T, H, hd):
    from mapdit_b200 import ops
    D = H * hd
    qkv = rnd(N * T, 3 * D, seed=6)
    ops.qk_normalize(qkv, D, hd)
    q, k, v = qkv.view(N, T, 3, H, hd).permute(2, 0, 3, 1, 4).cpu()
    raw = rnd(N * T, 3 * D, seed=6).view(N, T, 3, H, hd).permute(2, 0, 3, 1, 4).cpu()
    assert rel_l2(q, O.normalize(raw[0])) < 1e-6 and rel_l2(k, O.normalize(raw[1])) < 1e-6
    assert torch.equal(v, raw[2])
    o = torch.empty(N * T, D, device=dev)
    ops.cos_attn(qkv, o, N, T, H, hd)
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), scale=1 / math.sqrt(hd)).transpose(1, 2).reshape(N * T, D)
    assert rel_l2(o.cpu(), ref) < 2e-6


@pytest.mark.parametrize("name", ["DiT-XS/8", "DiT-S/4", "DiT-XS/2"])
def test_patch_embed_and_cond(dev, name):
    from mapdit_b200 import ops
    cfg = O.config_for(name)
    sd = O.init_state_dict(cfg, seed=9)
    N, D, T = 3, cfg.hidden_size, cfg.tokens
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, 4, 32, 32, generator=g)
    w = sd["x_embedder.weight"]
    weff = (O.normalize(w) / math.sqrt(w.shape[1])).contiguous()
    P = O.patchify(x, cfg.patch_size)
    P = torch.cat([P, torch.ones_like(P[:, :, :1])], -1)
    ref = O.mp_sum(F.linear(P, weff), sd["pos_embed"], 0.5).reshape(N * T, D)
    x0 = torch.empty(N * T, D, device=dev)
    ops.patch_embed(x.to(dev), weff.to(dev), sd["pos_embed"].to(dev).contiguous(), x0, None, None, None, None, 0, cfg.patch_size)
    assert rel_l2(x0.cpu(), ref) < 2e-6
    # fourier + label rows + combine
    t = torch.tensor([0, 999, 517])
    e = torch.empty(N, 256, device=dev)
    ops.fourier(t.to(dev), sd["t_embedder.embedding.scale"].to(dev), sd["t_embedder.embedding.shift"].to(dev), e)
    eref = O.fourier_features(t, sd["t_embedder.embedding.scale"], sd["t_embedder.embedding.shift"])
    assert (e.cpu() - eref).abs().max() < 2e-6  # no FMA contraction on the argument
    y = torch.tensor([5, 1000, 999])
    table = sd["y_embedder.embedding.weight"].to(dev)
    out = torch.empty(N, D, device=dev)
    ops.embed_rows(y.to(dev), None, 1000, table, out)
    assert rel_l2(out.cpu(), O.normalize(sd["y_embedder.embedding.weight"])[y]) < 1e-6
    mask = torch.tensor([1, 0, 0], dtype=torch.uint8, device=dev)
    ops.embed_rows(y.to(dev), mask, 1000, table, out)
    assert rel_l2(out.cpu(), O.normalize(sd["y_embedder.embedding.weight"])[torch.tensor([1000, 1000, 999])]) < 1e-6
    c, cs = torch.empty(N, D, device=dev), torch.empty(N, D, device=dev)
    cs16 = torch.empty(N, D, device=dev, dtype=torch.bfloat16)
    ops.cond_combine(e[:, :D].contiguous() if D <= 256 else out, out, c, cs, cs16)
    a = (e[:, :D].contiguous() if D <= 256 else out).cpu()
    cref = O.mp_sum(a, out.cpu(), 0.5)
    assert rel_l2(c.cpu(), cref) < 1e-6 and rel_l2(cs.cpu(), O.mp_silu(cref)) < 1e-6
    # mp_scale
    wmu = rnd(8, D, seed=12)
    ref8 = rnd(8, seed=13)
    s = torch.empty(N, device=dev)
    ops.mp_scale(c, wmu, ref8, s)
    sref = torch.sigmoid((c.cpu() @ wmu.cpu().t()) @ ref8.cpu() / math.sqrt(8))
    assert rel_l2(s.cpu(), sref) < 1e-5


def test_final_unpatchify_and_cfg(dev):
    from mapdit_b200 import ops
    for p in (2, 4, 8):
        N, C, S = 2, 4, 32
        T = (S // p) ** 2
        lin = rnd(N * T, 2 * p * p * C, seed=p)
        smu, ssg = rnd(N, seed=20).abs(), rnd(N, seed=21).abs()
        out = torch.empty(N, 2 * C, S, S, device=dev)
        ops.final_unpatchify(lin, smu, ssg, out, p)
        mean, sig = lin.view(N, T, -1).cpu().chunk(2, dim=-1)
        ref = torch.cat([O.unpatchify(mean * smu.cpu().view(-1, 1, 1), S, p), O.unpatchify(sig * ssg.cpu().view(-1, 1, 1), S, p)], 1)
        assert rel_l2(out.cpu(), ref) < 1e-6
    o = rnd(6, 8, 32, 32, seed=30)
    ref = o.clone().cpu()
    eps, rest = ref[:, :4], ref[:, 4:]
    cond, unc = eps[:3], eps[3:]
    half = unc + 2.5 * (cond - unc)
    ref = torch.cat([torch.cat([half, half], 0), rest], 1)
    ops.cfg_combine(o, 4, 2.5)
    assert rel_l2(o.cpu(), ref) < 1e-6


def test_diffusion_step_and_loss(dev):
    from mapdit_b200 import create_diffusion
    from mapdit_b200.diffusion import gaussian_diffusion as gd
    g = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "diffusion.npz"))
    x0, noise, mo = (torch.from_numpy(g[k]).to(dev) for k in ("tl_x0", "tl_noise", "tl_mo"))
    # training_losses against the reference's golden output and the oracle's gradient
    d = create_diffusion("")
    mo_req = mo.clone().requires_grad_(True)
    terms = d.training_losses(lambda *a, **k: mo_req, x0, torch.from_numpy(g["tl_t"]).to(dev), noise=noise)
    for k in ("loss", "mse", "vb"):
        assert rel_l2(terms[k].detach().cpu(), g["tl_" + k]) < 2e-6, k
    wts = torch.linspace(0.5, 1.5, x0.shape[0], device=dev)
    (terms["loss"] * wts).sum().backward()
    mo_cpu = mo.cpu().double().requires_grad_(True)
    T = O.make_tables("")
    tr = O.training_losses(T, lambda *a: mo_cpu.float(), x0.cpu(), torch.from_numpy(g["tl_t"]), noise.cpu())
    (tr["loss"].double() * wts.cpu().double()).sum().backward()
    assert rel_l2(mo_req.grad.cpu(), mo_cpu.grad) < 2e-5
    # p_sample against the reference's golden output (noise injected)
    d50 = create_diffusion("50")
    t50 = torch.from_numpy(g["ps_t"]).to(dev)
    real = gd._randn_like
    gd._randn_like = lambda x: noise
    try:
        for clip in (1, 0):
            out = d50.p_sample(lambda *a, **k: mo, x0, t50, clip_denoised=bool(clip))
            assert rel_l2(out["sample"].cpu(), g[f"ps_sample_clip{clip}"]) < 1e-6
            assert rel_l2(out["pred_xstart"].cpu(), g[f"ps_x0_clip{clip}"]) < 1e-6
            pm = d50.p_mean_variance(lambda *a, **k: mo, x0, t50, clip_denoised=bool(clip))
            ref = O.p_mean_variance(O.make_tables("50"), mo.cpu(), x0.cpu(), t50.cpu(), bool(clip))
            for k in ("mean", "variance", "log_variance", "pred_xstart"):
                assert rel_l2(pm[k].cpu(), ref[k]) < 1e-6, k
            # generic path with a python callback gives the same numbers
            out2 = d50.p_sample(lambda *a, **k: mo, x0, t50, clip_denoised=bool(clip), denoised_fn=lambda v: v)
            assert rel_l2(out2["sample"].cpu(), g[f"ps_sample_clip{clip}"]) < 1e-6
    finally:
        gd._randn_like = real
    # DDIM step against the reference's golden output (eta 0 and 0.7, noise injected)
    d25 = create_diffusion("ddim25")
    t25 = torch.from_numpy(g["dd_t"]).to(dev)
    gd._randn_like = lambda x: noise
    try:
        for eta in (0.0, 0.7):
            for clip in (1, 0):
                out = d25.ddim_sample(lambda *a, **k: mo, x0, t25, clip_denoised=bool(clip), eta=eta)
                assert rel_l2(out["sample"].cpu(), g[f"dd_sample_eta{eta}_clip{clip}"]) < 2e-6
                assert rel_l2(out["pred_xstart"].cpu(), g[f"dd_x0_eta{eta}_clip{clip}"]) < 1e-6
    finally:
        gd._randn_like = real
    xt = d.q_sample(x0, torch.from_numpy(g["tl_t"]).to(dev), noise=noise)
    assert rel_l2(xt.cpu(), O.q_sample(T, x0.cpu(), torch.from_numpy(g["tl_t"]), noise.cpu())) < 1e-6


def test_calc_bpd_loop_and_ddim_reverse_sample_vs_reference_golden(dev):
    """gaussian_diffusion.py:806-858 / :562-598 of the unmodified reference on a synthetic model output (oracle/make_golden.py
    eval_helpers_case): host passthroughs over the p_mean_variance / q_sample kernels"""
    import os
    from conftest import GOLDEN
    from mapdit_b200 import create_diffusion
    from mapdit_b200.diffusion import gaussian_diffusion as gd
    g = np.load(os.path.join(GOLDEN, "eval_helpers.npz"))
    x0, mo = torch.from_numpy(g["x0"]).to(dev), torch.from_numpy(g["mo"]).to(dev)
    d10 = create_diffusion("10")
    real = gd._randn_like
    try:
        for clip in (1, 0):
            it = iter(torch.from_numpy(g["noises"]).to(dev))
            gd._randn_like = lambda x: next(it)
            out = d10.calc_bpd_loop(lambda *a, **k: mo, x0, clip_denoised=bool(clip))
            for k in ("total_bpd", "prior_bpd", "vb", "xstart_mse", "mse"):
                e = rel_l2(out[k].cpu(), g[f"bpd_{k}_clip{clip}"])
                assert e < 2e-5, (k, clip, e)
    finally:
        gd._randn_like = real
    d25 = create_diffusion("ddim25")
    t25 = torch.from_numpy(g["rev_t"]).to(dev)
    for clip in (1, 0):
        out = d25.ddim_reverse_sample(lambda *a, **k: mo, x0, t25, clip_denoised=bool(clip))
        assert rel_l2(out["sample"].cpu(), g[f"rev_sample_clip{clip}"]) < 2e-6
        assert rel_l2(out["pred_xstart"].cpu(), g[f"rev_x0_clip{clip}"]) < 1e-6


def test_multi_tensor_weight_norm_bwd_cast2d_and_bf16_grad_adam(dev):
    from mapdit_b200 import ops
    # in-place multi-tensor weight-norm backward == the single-tensor kernel, tensor by tensor
    shapes = [(2304, 768), (768, 768), (3072, 768), (768, 3072), (4608, 768), (13, 64)]
    vs = [rnd(r, c, seed=10 + i) for i, (r, c) in enumerate(shapes)]
    gs = [rnd(r, c, seed=30 + i) for i, (r, c) in enumerate(shapes)]
    want = []
    for v, g in zip(vs, gs):
        o = torch.empty_like(v)
        ops.weight_norm_bwd(v, g, o)
        want.append(o)
    inplace = [g.clone() for g in gs]
    ops.WeightNormBwdBatch(list(zip(vs, inplace)), dev).run()
    for a, b in zip(inplace, want):
        assert rel_l2(a, b) < 1e-6
    # strided 2-D cast of a column window
    src = rnd(37, 500, seed=3)
    dst = torch.zeros(37, 640, device=dev, dtype=torch.bfloat16)
    ops.cast_2d(src[:, 100:356], dst[:, 8:264])
    assert torch.equal(dst[:, 8:264], src[:, 100:356].bfloat16()) and float(dst[:, :8].abs().sum()) == 0 and float(dst[:, 264:].abs().sum()) == 0
    # Adam with bf16 gradients == Adam with the same gradients widened to fp32 (odd length: scalar tail)
    n = 4099
    p0, g32 = rnd(n, seed=5), rnd(n, seed=6)
    g16 = g32.bfloat16()
    pa, ma, va = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    pb, mb, vb = p0.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in (1, 2, 3):
        ops.adam_step(pa, g16.float(), ma, va, 1e-2, 0.9, 0.99, 1e-8, step, grad_scale=0.5)
        ops.adam_step_g16(pb, g16, mb, vb, 1e-2, 0.9, 0.99, 1e-8, step, grad_scale=0.5)
    assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb)


@pytest.mark.parametrize("M,H,hd", [(300, 16, 72), (257, 6, 64), (64, 3, 32), (1000, 2, 128)])
def test_qk_normalize_bf16_vectorised(dev, M, H, hd):
    """bf16 q/k L2 normalisation with 16-byte accesses (src/layers/attention.py:43-45 for head dims the GEMM epilogue does not fuse):
    q, k heads scaled to norm sqrt(hd), v untouched, sc = sqrt(hd)/(||v|| + eps) saved for the backward"""
    from mapdit_b200 import ops
    D = H * hd
    raw = (rnd(M, 3 * D, seed=41) * 1.7).bfloat16()
    ref = raw.float().view(M, 3, H, hd).clone()
    nrm = ref[:, :2].norm(dim=-1, keepdim=True)
    want_sc = (math.sqrt(hd) / (nrm + 1e-4)).reshape(M, 2 * H)
    ref[:, :2] = ref[:, :2] * math.sqrt(hd) / (nrm + 1e-4)
    a, b = raw.clone(), raw.clone()
    sc = torch.full((M, 2 * H), float("nan"), device=dev)
    ops.qk_normalize_save(a, sc, D, hd)
    ops.qk_normalize(b, D, hd)
    assert torch.equal(a, b)
    assert torch.equal(a[:, 2 * D:], raw[:, 2 * D:])
    assert rel_l2(a.float(), ref.reshape(M, 3 * D)) < 3e-3  # bf16 output rounding
    assert rel_l2(sc, want_sc) < 1e-6
