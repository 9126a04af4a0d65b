"""Pins oracle/mapdit_oracle.py against outputs of the unmodified reference
(tests/golden/*.npz, made by oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import mapdit_oracle as O

TOL = 2e-6  # fp32 CPU restatement vs reference: same ops, allow BLAS/ISA reorder noise


def load(tag):
    return np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False)


def wsum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


@pytest.mark.parametrize("tag", ["eval_xs8", "eval_s4", "eval_xs2", "eval_b2"])
def test_eval_forward(tag):
    g = load(tag)
    cfg = O.config_for(str(g["name"]))
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    assert abs(wsum(sd) - float(g["wsum"])) <= 1e-9 * float(g["wsum"]), "weight generator drifted"
    with torch.no_grad():
        out = O.dit_forward(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]), torch.from_numpy(g["y"]))
    assert rel_l2(out, g["out"]) < TOL


def test_forward_with_cfg():
    g = load("cfg_xs4")
    cfg = O.config_for(str(g["name"]))
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    with torch.no_grad():
        out = O.dit_forward_with_cfg(sd, cfg, torch.from_numpy(g["x"]), torch.from_numpy(g["t"]),
                                     torch.from_numpy(g["y"]), float(g["cfg_scale"]))
    assert rel_l2(out, g["out"]) < TOL


def test_state_dict_contract():
    cfg = O.config_for("DiT-S/4")
    shp = O.param_shapes(cfg)
    assert len(shp) == 98  # 7*L + 14
    n = sum(int(np.prod(s)) for k, s in shp.items() if k not in O.BUFFER_KEYS)
    assert n == 32_855_849


@pytest.mark.parametrize("tag", ["train_xs8", "train_xs4", "train_xs2_b80"])
def test_train_step_grads(tag):
    g = load(tag)
    cfg = O.config_for(str(g["name"]))
    p = O.make_params(O.init_state_dict(cfg, seed=int(g["seed"])))
    T = O.make_tables("")
    x, t, y, noise, drop = O.golden_train_inputs(g, cfg)
    terms, grads = O.train_step_grads(p, cfg, T, x, t, y, noise, drop_mask=drop)
    for k in ("loss", "mse", "vb"):
        assert rel_l2(terms[k].detach(), g[k]) < 1e-5, k
    names = [str(s) for s in g["grad_names"]]
    stats = g["grad_stats"]
    assert names == [k for k in p if p[k].requires_grad]
    for i, k in enumerate(names):
        gr = grads[k].double()
        assert abs(gr.norm().item() - stats[i, 0]) <= 2e-5 * max(stats[i, 0], 1e-12), k
        proj = torch.from_numpy(np.random.default_rng(7000 + i).standard_normal(tuple(gr.shape))).double()
        assert abs((gr * proj).sum().item() - stats[i, 2]) <= 1e-4 * max(stats[i, 0], 1e-12), k
        # forced weight normalisation left the parameter normalised (mp_linear.py:38-40)
        assert abs(p[k].detach().double().norm().item() - stats[i, 3]) <= 1e-6 * stats[i, 3] + 1e-12, k
        if "grad::" + k in g.files:
            assert rel_l2(grads[k], g["grad::" + k]) < 2e-5, k


def test_diffusion_tables_and_maps():
    g = load("diffusion")
    for rs in ["", "50", "250", "10", "ddim25"]:
        T = O.make_tables(rs)
        key = rs or "full"
        assert T.timestep_map == [int(v) for v in g[f"{key}::timestep_map"]]
        np.testing.assert_allclose(T.betas, g[f"{key}::betas"], rtol=1e-13)
        for k, v in T.tabs.items():
            if k in ("log_betas", "alphas_cumprod", "alphas_cumprod_prev"):
                continue
            np.testing.assert_allclose(v, g[f"{key}::{k}"], rtol=1e-12, atol=0, err_msg=f"{key}::{k}")
    T = O.make_tables("50")
    assert T.timestep_map[:6] == [0, 20, 41, 61, 82, 102] and T.timestep_map[-4:] == [938, 958, 979, 999]
    assert O.make_tables("250").timestep_map[:3] == [0, 4, 8]


def test_training_losses_and_p_sample_synthetic():
    g = load("diffusion")
    mo = torch.from_numpy(g["tl_mo"])
    terms = O.training_losses(O.make_tables(""), lambda *a: mo, torch.from_numpy(g["tl_x0"]),
                              torch.from_numpy(g["tl_t"]), torch.from_numpy(g["tl_noise"]))
    for k in ("loss", "mse", "vb"):
        assert rel_l2(terms[k], g["tl_" + k]) < 1e-6, k
    T50 = O.make_tables("50")
    for clip in (1, 0):
        out = O.p_sample_step(T50, mo, torch.from_numpy(g["tl_x0"]), torch.from_numpy(g["ps_t"]),
                              torch.from_numpy(g["tl_noise"]), clip_denoised=bool(clip))
        assert rel_l2(out["sample"], g[f"ps_sample_clip{clip}"]) < 1e-6
        assert rel_l2(out["pred_xstart"], g[f"ps_x0_clip{clip}"]) < 1e-6


def test_ddim_step_synthetic():
    g = load("diffusion")
    mo, x, noise, t = (torch.from_numpy(g[k]) for k in ("tl_mo", "tl_x0", "tl_noise", "dd_t"))
    T = O.make_tables("ddim25")
    for eta in (0.0, 0.7):
        for clip in (1, 0):
            out = O.ddim_step(T, mo, x, t, noise, clip_denoised=bool(clip), eta=eta)
            assert rel_l2(out["sample"], g[f"dd_sample_eta{eta}_clip{clip}"]) < 1e-6
            assert rel_l2(out["pred_xstart"], g[f"dd_x0_eta{eta}_clip{clip}"]) < 1e-6


@pytest.mark.parametrize("tag", ["loop_xs8", "loop_xs4_cfg"])
def test_sampling_loop_free_running(tag):
    g = load(tag)
    cfg = O.config_for(str(g["name"]))
    sd = O.init_state_dict(cfg, seed=int(g["seed"]))
    T = O.make_tables(str(int(g["steps"])))
    y = torch.from_numpy(g["y"])
    if bool(g["use_cfg"]):
        fn = lambda x, t: O.dit_forward_with_cfg(sd, cfg, x, t, y, 1.5)
    else:
        fn = lambda x, t: O.dit_forward(sd, cfg, x, t, y)
    noises = [torch.from_numpy(n) for n in g["noises"]]
    trace = []
    O.p_sample_loop(T, fn, torch.from_numpy(g["z"]), noises, clip_denoised=True, trace=trace)
    ref = g["samples_clip1"]
    for k, o in enumerate(trace):
        assert rel_l2(o["sample"], ref[k]) < 1e-5, k
    # clip_denoised=False: the reference itself overflows on random weights;
    # compare teacher-forced on the finite prefix.
    ref0 = g["samples_clip0"]
    teacher = [torch.from_numpy(g["z"])] + [torch.from_numpy(r) for r in ref0[:-1]]
    trace = []
    O.p_sample_loop(T, fn, torch.from_numpy(g["z"]), noises, clip_denoised=False, teacher=teacher, trace=trace)
    checked = 0
    for k, o in enumerate(trace):
        if np.isfinite(ref0[k]).all() and np.isfinite(teacher[k].numpy()).all():
            assert rel_l2(o["sample"], ref0[k]) < 1e-5, k
            checked += 1
    assert checked >= 1
