"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on seeded inputs.  Run in the build container only:

    python oracle/make_golden.py

The reference cannot travel to the GPU box, so its outputs are committed as small
fixtures; tests/test_oracle_golden.py pins oracle/mapdit_oracle.py against them.
Weights come from oracle.init_state_dict (numpy PCG64 — platform stable) and are loaded
into the reference modules with load_state_dict, so no weights need to be stored; a
float64 checksum of the weights is stored to detect generator drift.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("MAPDIT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import mapdit_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def ref_model(name, cfg, sd):
    from src.models import DIT_MODELS
    m = DIT_MODELS[name](in_channels=cfg.in_channels, input_size=cfg.input_size, num_classes=cfg.num_classes)
    missing = m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m


def checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def proj_vec(shape, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal(shape)).double()


inputs = O.seeded_inputs  # (x, t, y) from numpy PCG64: platform stable, so large cases store the seed instead of the tensors


def eval_case(tag, name, B, seed):
    cfg = O.config_for(name)
    sd = O.init_state_dict(cfg, seed=seed)
    m = ref_model(name, cfg, sd).eval()
    x, t, y = inputs(cfg, B, seed + 100)
    with torch.no_grad():
        out = m(x, t, y)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), name=name, seed=seed, x=x.numpy(), t=t.numpy(), y=y.numpy(),
                        out=out.numpy(), wsum=checksum(sd))
    print(tag, out.shape, float(out.abs().mean()))


def cfg_case(tag, name, B, seed, scale):
    cfg = O.config_for(name)
    sd = O.init_state_dict(cfg, seed=seed)
    m = ref_model(name, cfg, sd).eval()
    x, t, y = inputs(cfg, B, seed + 100)
    x = torch.cat([x, x], 0)
    t = torch.cat([t, t], 0)
    y = torch.cat([y, torch.full_like(y, cfg.num_classes)], 0)
    with torch.no_grad():
        out = m.forward_with_cfg(x, t, y, scale)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), name=name, seed=seed, x=x.numpy(), t=t.numpy(), y=y.numpy(),
                        out=out.numpy(), cfg_scale=scale, wsum=checksum(sd))
    print(tag, out.shape)


def train_case(tag, name, B, seed, store_inputs=True):
    """One reference training forward/backward (train.py:86-95 without the optimiser).  `store_inputs=False` (large batches)
    keeps x / noise out of the fixture: tests regenerate them with O.seeded_inputs / O.seeded_noise from the stored seed."""
    from diffusion import create_diffusion
    cfg = O.config_for(name)
    sd = O.init_state_dict(cfg, seed=seed)
    m = ref_model(name, cfg, sd).train()
    x, t, y = inputs(cfg, B, seed + 100)
    t[0] = 0  # exercise the decoder-NLL branch
    noise = O.seeded_noise(x.shape, seed + 200)
    torch.manual_seed(seed + 300)
    drop = torch.rand(B) < cfg.class_dropout_prob
    drop[1] = True
    # make the reference draw exactly this mask
    import src.blocks.label_embedder as le
    real_rand = torch.rand
    le.torch.rand = lambda *a, **k: torch.where(drop, 0.0, 1.0)
    try:
        diff = create_diffusion(timestep_respacing="")
        terms = diff.training_losses(m, x, t, dict(y=y), noise=noise)
        terms["loss"].mean().backward()
    finally:
        le.torch.rand = real_rand
    rec = dict(name=name, seed=seed, batch=B, t=t.numpy(), y=y.numpy(), drop=drop.numpy(),
               loss=terms["loss"].detach().numpy(), mse=terms["mse"].detach().numpy(), vb=terms["vb"].detach().numpy(),
               wsum=checksum(sd), xsum=float(x.double().abs().sum()), nsum=float(noise.double().abs().sum()))
    if store_inputs:
        rec.update(x=x.numpy(), noise=noise.numpy())
    names, stats = [], []
    for i, (k, p) in enumerate(m.named_parameters()):
        g = p.grad.double()
        names.append(k)
        stats.append([g.norm().item(), g.sum().item(), (g * proj_vec(g.shape, 7000 + i)).sum().item(),
                      p.detach().double().norm().item()])  # last = weight norm after forced WN
        if g.numel() <= 8 * cfg.hidden_size:
            rec["grad::" + k] = p.grad.numpy()
    rec["grad_names"] = np.array(names)
    rec["grad_stats"] = np.array(stats, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **rec)
    print(tag, terms["loss"].detach())


def diffusion_case(tag):
    from diffusion import create_diffusion
    rec = {}
    for rs in ["", "50", "250", "10", "ddim25"]:
        d = create_diffusion(timestep_respacing=rs)
        key = rs or "full"
        rec[f"{key}::timestep_map"] = np.array(d.timestep_map)
        rec[f"{key}::betas"] = d.betas
        for k in ["sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"]:
            rec[f"{key}::{k}"] = getattr(d, k)
    # training_losses / p_sample on a synthetic model output
    g = np.random.default_rng(11)
    B = 6
    x0 = torch.from_numpy(g.standard_normal((B, 4, 8, 8), dtype=np.float32))
    x0[0] = x0[0].clamp(-1.2, 1.2)
    x0[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.9995, -0.9995])
    noise = torch.from_numpy(g.standard_normal((B, 4, 8, 8), dtype=np.float32))
    mo = torch.from_numpy(g.standard_normal((B, 8, 8, 8), dtype=np.float32))
    t = torch.tensor([0, 1, 17, 500, 998, 999])
    d = create_diffusion(timestep_respacing="")
    terms = d.training_losses(lambda *a, **k: mo, x0, t, noise=noise)
    rec.update(tl_x0=x0.numpy(), tl_noise=noise.numpy(), tl_mo=mo.numpy(), tl_t=t.numpy(),
               tl_loss=terms["loss"].numpy(), tl_mse=terms["mse"].numpy(), tl_vb=terms["vb"].numpy())
    import diffusion.gaussian_diffusion as gd
    d50 = create_diffusion(timestep_respacing="50")
    t50 = torch.tensor([0, 1, 7, 25, 48, 49])
    for clip in (True, False):
        real = gd.th.randn_like
        gd.th.randn_like = lambda x: noise
        try:
            out = d50.p_sample(lambda *a, **k: mo, x0, t50, clip_denoised=clip)
        finally:
            gd.th.randn_like = real
        rec[f"ps_sample_clip{int(clip)}"] = out["sample"].numpy()
        rec[f"ps_x0_clip{int(clip)}"] = out["pred_xstart"].numpy()
    rec["ps_t"] = t50.numpy()
    # DDIM steps (gaussian_diffusion.py:513-560) on the ddim25 spacing, eta 0 and 0.7
    d25 = create_diffusion(timestep_respacing="ddim25")
    t25 = torch.tensor([0, 1, 7, 12, 23, 24])
    rec["dd_t"] = t25.numpy()
    for eta in (0.0, 0.7):
        for clip in (True, False):
            real = gd.th.randn_like
            gd.th.randn_like = lambda x: noise
            try:
                out = d25.ddim_sample(lambda *a, **k: mo, x0, t25, clip_denoised=clip, eta=eta)
            finally:
                gd.th.randn_like = real
            rec[f"dd_sample_eta{eta}_clip{int(clip)}"] = out["sample"].numpy()
            rec[f"dd_x0_eta{eta}_clip{int(clip)}"] = out["pred_xstart"].numpy()
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **rec)
    print(tag, "ok")


def eval_helpers_case(tag):
    """Reference calc_bpd_loop (gaussian_diffusion.py:806-858) and ddim_reverse_sample (:562-598) on a synthetic model output."""
    from diffusion import create_diffusion
    import diffusion.gaussian_diffusion as gd
    g = np.random.default_rng(23)
    B = 5
    x0 = torch.from_numpy(g.standard_normal((B, 4, 8, 8), dtype=np.float32)).clamp(-1.3, 1.3)
    x0[0, 0, 0, :4] = torch.tensor([-1.0, 1.0, 0.9995, -0.9995])
    mo = torch.from_numpy(g.standard_normal((B, 8, 8, 8), dtype=np.float32))
    d10 = create_diffusion(timestep_respacing="10")
    noises = [torch.from_numpy(g.standard_normal((B, 4, 8, 8), dtype=np.float32)) for _ in range(10)]
    rec = dict(x0=x0.numpy(), mo=mo.numpy(), noises=np.stack([n.numpy() for n in noises]))
    for clip in (True, False):
        it = iter(noises)
        real = gd.th.randn_like
        gd.th.randn_like = lambda x: next(it)
        try:
            out = d10.calc_bpd_loop(lambda *a, **k: mo, x0, clip_denoised=clip)
        finally:
            gd.th.randn_like = real
        for k, v in out.items():
            rec[f"bpd_{k}_clip{int(clip)}"] = v.numpy()
    d25 = create_diffusion(timestep_respacing="ddim25")
    t25 = torch.tensor([0, 1, 7, 12, 24])
    rec["rev_t"] = t25.numpy()
    for clip in (True, False):
        out = d25.ddim_reverse_sample(lambda *a, **k: mo, x0, t25, clip_denoised=clip)
        rec[f"rev_sample_clip{int(clip)}"] = out["sample"].numpy()
        rec[f"rev_x0_clip{int(clip)}"] = out["pred_xstart"].numpy()
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **rec)
    print(tag, "ok", rec["bpd_total_bpd_clip1"])


def loop_case(tag, name, B, seed, steps, use_cfg):
    """Reference p_sample_loop_progressive with injected noise (sample.py:52-61 call pattern)."""
    from diffusion import create_diffusion
    import diffusion.gaussian_diffusion as gd
    cfg = O.config_for(name)
    sd = O.init_state_dict(cfg, seed=seed)
    m = ref_model(name, cfg, sd).eval()
    g = np.random.default_rng(seed + 400)
    z = torch.from_numpy(g.standard_normal((B, 4, 32, 32), dtype=np.float32))
    y = torch.from_numpy(g.integers(0, 1000, size=(B,))).long()
    noises = [torch.from_numpy(g.standard_normal((2 * B if use_cfg else B, 4, 32, 32), dtype=np.float32)) for _ in range(steps)]
    if use_cfg:
        zz = torch.cat([z, z], 0)
        noises = [torch.cat([n[:B], n[:B]], 0) for n in noises]
        kw = dict(y=torch.cat([y, torch.full_like(y, 1000)], 0), cfg_scale=1.5)
        fn = m.forward_with_cfg
    else:
        zz, kw, fn = z, dict(y=y), m.forward
    rec = dict(name=name, seed=seed, steps=steps, use_cfg=use_cfg, z=zz.numpy(), y=kw["y"].numpy(),
               noises=np.stack([n.numpy() for n in noises]), wsum=checksum(sd))
    d = create_diffusion(str(steps))
    for clip in (True, False):
        it = iter(noises)
        real = gd.th.randn_like
        gd.th.randn_like = lambda x: next(it)
        try:
            outs = [o["sample"].numpy() for o in d.p_sample_loop_progressive(fn, zz.shape, zz, clip_denoised=clip,
                                                                               model_kwargs=kw, device="cpu")]
        finally:
            gd.th.randn_like = real
        rec[f"samples_clip{int(clip)}"] = np.stack(outs)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **rec)
    print(tag, "ok", [float(np.abs(rec[f'samples_clip{c}'][-1]).mean()) for c in (1, 0)])


def ema_case(tag):
    """Reference src/ema.py host math + a post-hoc reconstruction from four synthetic fp16 snapshots."""
    import tempfile
    from src import ema as R
    stds = np.array([0.05, 0.1, 0.075])
    rec = dict(stds=stds, gammas=R.std_to_gamma(stds), back=R.gamma_to_std(R.std_to_gamma(stds)))
    rec["beta_ts"] = np.array([1, 2, 100, 40000])
    rec["betas"] = np.array([[R.calc_beta(s, t) for t in rec["beta_ts"]] for s in (0.05, 0.1)])
    snaps = [(0.05, 500), (0.1, 500), (0.05, 1000), (0.1, 1000)]
    rec["snap_stds"] = np.array([s for s, _ in snaps])
    rec["snap_ts"] = np.array([t for _, t in snaps])
    rec["weights"] = R.solve_weights(rec["snap_ts"], R.std_to_gamma(rec["snap_stds"]), 1000, R.std_to_gamma(0.075))
    g = np.random.default_rng(77)
    with tempfile.TemporaryDirectory() as d:
        for i, (s, t) in enumerate(snaps):
            sd = {"a.weight": torch.from_numpy(g.standard_normal((5, 7), dtype=np.float32)).half(),
                  "b": torch.from_numpy(g.standard_normal((3,), dtype=np.float32)).half()}
            rec[f"snap{i}_a"], rec[f"snap{i}_b"] = sd["a.weight"].float().numpy(), sd["b"].float().numpy()
            torch.save({"std": s, "t": t, "state_dict": sd}, os.path.join(d, f"{s:.3f}_{t:07d}.pt"))
        out = R.calculate_posthoc_ema(0.075, d, verbose=False)
        rec["posthoc_a"], rec["posthoc_b"] = out["a.weight"].numpy(), out["b"].numpy()
        hit = R.calculate_posthoc_ema(0.1, d, verbose=False)
        rec["exact_a"] = hit["a.weight"].float().numpy()
    # EMA.update semantics on a tiny module: three updates of two tracked copies (src/ema.py:124-140)
    net = torch.nn.Linear(6, 4)
    with torch.no_grad():
        net.weight.copy_(torch.from_numpy(g.standard_normal((4, 6), dtype=np.float32)))
        net.bias.copy_(torch.from_numpy(g.standard_normal((4,), dtype=np.float32)))
    with tempfile.TemporaryDirectory() as d:
        e = R.EMA(net, d)
        rec["upd_w0"], rec["upd_b0"] = net.weight.detach().numpy().copy(), net.bias.detach().numpy().copy()
        deltas = []
        for t in (1, 2, 3):
            dw = torch.from_numpy(g.standard_normal((4, 6), dtype=np.float32))
            db = torch.from_numpy(g.standard_normal((4,), dtype=np.float32))
            deltas.append((dw.numpy(), db.numpy()))
            with torch.no_grad():
                net.weight.add_(dw)
                net.bias.add_(db)
            e.update(t, net)
        rec["upd_dw"] = np.stack([a for a, _ in deltas])
        rec["upd_db"] = np.stack([b for _, b in deltas])
        for s in (0.05, 0.1):
            rec[f"upd_w_{s}"] = e.emas[s].weight.numpy().copy()
            rec[f"upd_b_{s}"] = e.emas[s].bias.numpy().copy()
    # LR schedule factor (train.py:179-197): the function is lifted out of the reference's train.py source (the module
    # itself imports torchvision/yaml and opens a dataset), default recipe num_lin_warmup=1000 / start_decay=20000 + edge cases
    import ast
    src = open(os.path.join(REF, "train.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "create_lr_lambda")
    ns = {"math": __import__("math")}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "train.py", "exec"), ns)
    steps = np.array([0, 1, 5, 998, 999, 1000, 19999, 20000, 20001, 80000, 400000])
    rec["lr_steps"] = steps
    rec["lr_factors"] = np.array([[ns["create_lr_lambda"](w, d)(int(s)) for s in steps] for w, d in ((1000, 20000), (1, 10), (100, 100))])
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **rec)
    print(tag, "ok")


if __name__ == "__main__":
    torch.set_num_threads(8)
    if len(sys.argv) > 1:  # regenerate selected fixtures only: python oracle/make_golden.py ema
        for a in sys.argv[1:]:
            {"ema": lambda: ema_case("ema"),
             # the bench path's kernels (256 tokens: fused tcgen05 attention backward; 80 x 256 rows: every block GEMM of
             # DiT-XS/2 qualifies for the cta_group::2 kernel), one t == 0 and one dropped label
             "train_xs2_b80": lambda: train_case("train_xs2_b80", "DiT-XS/2", 80, 9, store_inputs=False),
             "eval_b2": lambda: eval_case("eval_b2", "DiT-B/2", 2, 10),
             "eval_helpers": lambda: eval_helpers_case("eval_helpers")}[a]()
        sys.exit(0)
    eval_case("eval_xs8", "DiT-XS/8", 3, 1)
    eval_case("eval_s4", "DiT-S/4", 2, 2)
    eval_case("eval_xs2", "DiT-XS/2", 1, 3)
    cfg_case("cfg_xs4", "DiT-XS/4", 2, 4, 4.0)
    train_case("train_xs8", "DiT-XS/8", 4, 5)
    train_case("train_xs4", "DiT-XS/4", 3, 6)
    diffusion_case("diffusion")
    loop_case("loop_xs8", "DiT-XS/8", 2, 7, 6, False)
    loop_case("loop_xs4_cfg", "DiT-XS/4", 1, 8, 5, True)
    ema_case("ema")
    train_case("train_xs2_b80", "DiT-XS/2", 80, 9, store_inputs=False)
    eval_case("eval_b2", "DiT-B/2", 2, 10)
    eval_helpers_case("eval_helpers")
