"""Rotation modulation (north_star / BASELINE config 3 "rotation-and-scaling").  The reference snapshot contains no code
for it: the checker is the oracle's own restatement of the README's description, so these tests
are SELF-REFERENTIAL — they pin the CUDA path to the oracle, not to the reference."""
import pytest
import torch

from conftest import rel_l2
from oracle import mapdit_oracle as O

pytestmark = pytest.mark.gpu


def build(name, modulation, dtype, seed=31):
    import mapdit_b200 as M
    cfg = O.config_for(name, modulation=modulation)
    sd = O.init_state_dict(cfg, seed=seed)
    m = M.DIT_MODELS[name](in_channels=4, input_size=32, num_classes=1000, compute_dtype=dtype, modulation=modulation)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == O.param_shapes(cfg)
    m.load_state_dict(sd)
    return m.cuda(), cfg, sd


@pytest.mark.parametrize("modulation", ["rotation_scaling", "rotation"])
@pytest.mark.parametrize("dtype,tol", [("fp32", 1e-5), ("bf16", 3e-2)])
def test_rotation_forward(modulation, dtype, tol):
    m, cfg, sd = build("DiT-XS/4", modulation, dtype)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 4, 32, 32, generator=g)
    t = torch.randint(0, 1000, (3,), generator=g)
    y = torch.randint(0, 1000, (3,), generator=g)
    with torch.no_grad():
        ref = O.dit_forward(sd, cfg, x, t, y)
        out = m.eval()(x.cuda(), t.cuda(), y.cuda())
    e = rel_l2(out.cpu(), ref)
    print(f"{modulation} {dtype}: forward rel-L2 vs oracle {e:.2e}")
    assert e < tol


@pytest.mark.parametrize("modulation", ["rotation_scaling", "rotation"])
@pytest.mark.parametrize("dtype,tol", [("fp32", 2e-4), ("bf16", 8e-2)])
def test_rotation_training_gradients(modulation, dtype, tol):
    import mapdit_b200 as M
    m, cfg, sd = build("DiT-XS/8", modulation, dtype)
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
    worst = 0.0
    gain_scale = max(float(v.abs()) for k, v in ograds.items() if v.dim() == 0)
    for k, prm in m.named_parameters():
        if dtype == "bf16" and prm.dim() == 0:
            # a scalar gain's gradient is a sum of +/- terms over every token and channel pair: under bf16 rounding the
            # cancellation leaves an ABSOLUTE error set by the size of the summands, so it is bounded against the largest
            # gain gradient of the model rather than against its own (possibly tiny) value
            assert abs(float(prm.grad) - float(ograds[k])) < 0.1 * gain_scale, (k, float(prm.grad), float(ograds[k]))
            continue
        e = rel_l2(prm.grad.cpu(), ograds[k])
        worst = max(worst, e)
        assert e < tol, (k, e)
    print(f"{modulation} {dtype}: worst per-parameter grad rel-L2 vs oracle {worst:.2e}")
