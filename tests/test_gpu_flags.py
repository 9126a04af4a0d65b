"""README `--use-*` switches turned OFF (README.md:59-66; BASELINE config 4 "AdaLN baseline, no mp flags").

The reference snapshot hard-codes every switch on and ships no "off" branch, so the checker is
the oracle's own restatement of the vanilla DiT ops: these tests are SELF-REFERENTIAL — they pin the CUDA path to the
oracle, not to the reference (parity unpinned)."""
import pytest
import torch

from conftest import rel_l2
from oracle import mapdit_oracle as O

pytestmark = pytest.mark.gpu

FLAGS = ("use_cosine_attention", "use_weight_normalization", "use_forced_weight_normalization", "use_mp_residual",
         "use_mp_silu", "use_no_layernorm", "use_mp_pos_enc", "use_mp_embedding")
ALL_OFF = {k: False for k in FLAGS}
CASES = [{k: False} for k in FLAGS] + [ALL_OFF, {"use_cosine_attention": False, "use_no_layernorm": False},
                                       {"use_mp_residual": False, "use_mp_silu": False, "use_mp_pos_enc": False}]
IDS = [("-".join(sorted(c)) if len(c) < 8 else "all_off").replace("use_", "") for c in CASES]


def build(name, flags, dtype, seed=41):
    import mapdit_b200 as M
    cfg = O.config_for(name, **flags)
    sd = O.init_state_dict(cfg, seed=seed)
    m = M.DIT_MODELS[name](in_channels=4, input_size=32, num_classes=1000, compute_dtype=dtype, **flags)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == O.param_shapes(cfg)  # same keys with any flag set
    m.load_state_dict(sd)
    return m.cuda(), cfg, sd


def test_default_init_statistics_follow_the_flags():
    import mapdit_b200 as M
    m = M.DIT_MODELS["DiT-XS/4"](in_channels=4, input_size=32, num_classes=10, use_weight_normalization=False, use_mp_pos_enc=False)
    w = m.blocks[0].mlp.net[0].weight
    assert abs(float(w.std()) - w.shape[1] ** -0.5) < 0.1 * w.shape[1] ** -0.5
    assert torch.allclose(m.pos_embed, O.pos_embed_table(256, 8, normalized=False))
    assert m.variant == 4


@pytest.mark.parametrize("flags", CASES, ids=IDS)
@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-5), ("bf16", 3e-2)])
def test_flags_off_forward(flags, dtype, tol):
    m, cfg, sd = build("DiT-XS/4", flags, dtype)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 4, 32, 32, generator=g)
    t = torch.randint(0, 1000, (3,), generator=g)
    y = torch.randint(0, 1000, (3,), generator=g)
    with torch.no_grad():
        ref = O.dit_forward(sd, cfg, x, t, y)
        out = m.eval()(x.cuda(), t.cuda(), y.cuda())
    e = rel_l2(out.cpu(), ref)
    print(f"{flags} {dtype}: forward rel-L2 vs oracle {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("flags", CASES, ids=IDS)
@pytest.mark.parametrize("dtype,tol", [("fp32", 3e-4), ("bf16", 1e-1)])
def test_flags_off_training_gradients(flags, dtype, tol):
    import mapdit_b200 as M
    m, cfg, sd = build("DiT-XS/8", flags, dtype)
    m.train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 4, 32, 32, generator=g)
    t = torch.randint(0, 1000, (4,), generator=g)
    y = torch.randint(0, 1000, (4,), generator=g)
    noise = torch.randn(4, 4, 32, 32, generator=g)
    drop = torch.tensor([False, True, False, False])
    d = M.create_diffusion("")
    terms = d.training_losses(lambda xt, tt, **kw: m(xt, tt, kw["y"], drop_mask=drop.cuda()), x.cuda(), t.cuda(), dict(y=y.cuda()),
                              noise=noise.cuda())
    terms["loss"].mean().backward()
    p = O.make_params(sd)
    oterms, ograds = O.train_step_grads(p, cfg, O.make_tables(""), x, t, y, noise, drop_mask=drop)
    assert rel_l2(terms["loss"].detach().cpu(), oterms["loss"].detach()) < (2e-5 if dtype == "fp32" else 3e-2)
    worst, gain_scale = 0.0, max([float(v.abs()) for v in ograds.values() if v is not None and v.dim() == 0] + [1e-30])
    for k, prm in m.named_parameters():
        og = ograds[k]
        if og is None:  # LayerNorm adaLN does not use the gains: zero gradient here, None in autograd
            assert float(prm.grad.abs().max()) == 0.0, k
            continue
        if dtype == "bf16" and prm.dim() == 0:
            assert abs(float(prm.grad) - float(og)) < 0.1 * gain_scale, (k, float(prm.grad), float(og))
            continue
        e = rel_l2(prm.grad.cpu(), og)
        worst = max(worst, e)
        assert e < tol, (k, e)
        # forced weight normalisation writes back only when both switches are on
        forced = cfg.use_weight_normalization and cfg.use_forced_weight_normalization
        if prm.dim() == 2 and not forced and k != "y_embedder.embedding.weight":
            assert torch.equal(prm.detach().cpu(), sd[k]), k
    print(f"{flags} {dtype}: worst per-parameter grad rel-L2 vs oracle {worst:.2e}")


@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("bf16", 5e-2)])
def test_adaln_baseline_sampling_loop(dtype, tol):
    """BASELINE config 4's shape of work at test size: the all-flags-off model through the CUDA-graph sampling loop"""
    import mapdit_b200 as M
    from mapdit_b200.diffusion import gaussian_diffusion as gd
    m, cfg, sd = build("DiT-XS/4", ALL_OFF, dtype)
    m.eval()
    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 4, 32, 32, generator=g)
    y = torch.tensor([3, 7])
    noises = [torch.randn(2, 4, 32, 32, generator=g) for _ in range(4)]
    ref = O.p_sample_loop(O.make_tables("4"), lambda a, b: O.dit_forward(sd, cfg, a, b, y), z, noises)
    it = iter(noises)
    real = gd._randn_like
    gd._randn_like = lambda v: next(it).cuda()
    try:
        s = M.create_diffusion("4").p_sample_loop(m.forward, z.shape, z.cuda(), model_kwargs=dict(y=y.cuda()), device="cuda")
    finally:
        gd._randn_like = real
    e = rel_l2(s.cpu(), ref)
    print(f"all-off {dtype}: 4-step sampling rel-L2 vs oracle {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("patch", [8, 2])
def test_dit_xl_head_dim_72_runs_in_bf16(patch):
    """DiT-XL (head_dim 72, BASELINE config 5) on the bf16 path, two blocks deep: tcgen05 GEMMs with the fused residual + modulation
    epilogues; patch 8 (16 tokens) takes the generic attention kernels, patch 2 (256 tokens) the head_dim-72 tcgen05 attention
    forward (two-panel tiles) and the tcgen05 dq + dkv backward pair"""
    import mapdit_b200 as M
    name = f"DiT-XL/{patch}"
    cfg = O.config_for(name)
    cfg.depth = 2
    sd = O.init_state_dict(cfg, seed=5)
    m = M.DiT(depth=2, hidden_size=1152, patch_size=patch, num_heads=16, in_channels=4, input_size=32, num_classes=1000)
    m.load_state_dict(sd)
    m = m.cuda()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 32, 32, generator=g)
    t = torch.randint(0, 1000, (2,), generator=g)
    y = torch.randint(0, 1000, (2,), generator=g)
    with torch.no_grad():
        ref = O.dit_forward(sd, cfg, x, t, y)
        out = m.eval()(x.cuda(), t.cuda(), y.cuda())
    assert rel_l2(out.cpu(), ref) < 3e-2
    m.train()
    noise = torch.randn(2, 4, 32, 32, generator=g)
    drop = torch.tensor([False, True])
    d = M.create_diffusion("")
    terms = d.training_losses(lambda xt, tt, **kw: m(xt, tt, kw["y"], drop_mask=drop.cuda()), x.cuda(), t.cuda(), dict(y=y.cuda()),
                              noise=noise.cuda())
    terms["loss"].mean().backward()
    p = O.make_params(sd)
    oterms, ograds = O.train_step_grads(p, cfg, O.make_tables(""), x, t, y, noise, drop_mask=drop)
    assert rel_l2(terms["loss"].detach().cpu(), oterms["loss"].detach()) < 3e-2
    for k, prm in m.named_parameters():
        if prm.dim() == 2:
            assert rel_l2(prm.grad.cpu(), ograds[k]) < 1e-1, k
