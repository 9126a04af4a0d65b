"""CPU-only checks of the host side: API surface, state-dict contract, respacing tables, the C ABI
library loading and exporting every declared symbol, and the loud failure without a GPU."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from oracle import mapdit_oracle as O


def test_registry_names_and_alias():
    import mapdit_b200 as M
    assert set(M.DIT_MODELS) == {f"DiT-{s}/{p}" for s in ("XS", "S", "B", "L", "XL") for p in (2, 4, 8)}
    assert M.DiT_models is M.DIT_MODELS


@pytest.mark.parametrize("name", ["DiT-S/4", "DiT-XS/2", "DiT-B/2"])
def test_state_dict_contract_matches_reference(name):
    import mapdit_b200 as M
    m = M.DIT_MODELS[name](in_channels=4, input_size=32, num_classes=1000)
    cfg = O.config_for(name)
    want = O.param_shapes(cfg)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    # named_parameters order equals the reference's registration order (pinned by the golden grad_names)
    g = np.load(os.path.join(GOLDEN, "train_xs8.npz"))
    if name == "DiT-XS/2":
        assert [k for k, _ in m.named_parameters()] == [str(s) for s in g["grad_names"]]
    m.load_state_dict(O.init_state_dict(cfg, seed=0), strict=True)
    m2 = copy.deepcopy(m)  # src/ema.py:121 deep-copies the model
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert m.get_parameter("blocks.0.attn.qkv_proj.weight").shape == (3 * cfg.hidden_size, cfg.hidden_size)
    assert torch.allclose(m.pos_embed, O.pos_embed_table(cfg.hidden_size, 32 // cfg.patch_size))


def test_unbuilt_variants_fail_loudly():
    import mapdit_b200 as M
    with pytest.raises(NotImplementedError):  # LayerNorm adaLN is only defined for the adaln layout
        M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10, use_no_layernorm=False, modulation="rotation")
    off = M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10, use_mp_silu=False, use_cosine_attention=False)
    assert off.variant == 2 | 16 and off.state_dict().keys() == M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10).state_dict().keys()
    with pytest.raises(ValueError):
        M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10, modulation="bogus")
    m = M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10, modulation="rotation_scaling")
    assert m.blocks[0].modulation[1].weight.shape == (5 * 256, 256)  # 6D -> 5D: the README's "~5.4 % fewer parameters"
    with pytest.raises(TypeError):
        M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10, bogus=True)


def test_cpu_tensors_are_rejected_not_emulated():
    import mapdit_b200 as M
    m = M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=10).eval()
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.zeros(1, 4, 32, 32), torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long))


def test_diffusion_tables_match_reference_golden():
    import mapdit_b200 as M
    g = np.load(os.path.join(GOLDEN, "diffusion.npz"))
    for rs in ["", "50", "250", "10", "ddim25"]:
        d = M.create_diffusion(rs)
        key = rs or "full"
        assert d.timestep_map == [int(v) for v in g[f"{key}::timestep_map"]]
        assert d.num_timesteps == len(d.timestep_map)
        np.testing.assert_allclose(d.betas, g[f"{key}::betas"], rtol=1e-13)
        for k in ["sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]:
            np.testing.assert_allclose(getattr(d, k), g[f"{key}::{k}"], rtol=1e-12, err_msg=f"{key}::{k}")
    with pytest.raises(ValueError):
        M.create_diffusion("2000")


def test_space_timesteps_sections():
    from mapdit_b200.diffusion import space_timesteps
    assert sorted(space_timesteps(300, [10, 15, 20]))[:3] == sorted(O.space_timesteps(300, [10, 15, 20]))[:3]
    assert space_timesteps(300, "10,15,20") == set(O.space_timesteps(300, "10,15,20"))
    assert space_timesteps(1000, "ddim50") == set(range(0, 1000, 20))


def test_library_loads_and_exports_every_declared_symbol():
    from mapdit_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mapdit.h")).read()
    declared = set(re.findall(r"\b(mapdit_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mapdit_gemm_args"}
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.SIGNATURES) <= declared
    assert _lib.lib().mapdit_abi_version() == 1
    assert ctypes.sizeof(_lib.GemmArgs) == L.mapdit_sizeof_gemm_args()


def test_off_path_entry_points_say_so():
    import mapdit_b200 as M
    d = M.create_diffusion("ddim25")
    with pytest.raises(NotImplementedError):
        d.ddim_sample(None, torch.zeros(1, 4, 8, 8), torch.zeros(1).long(), cond_fn=lambda *a, **k: None)
    with pytest.raises(NotImplementedError):
        M.create_diffusion("", predict_xstart=True).training_losses(None, torch.zeros(1, 4, 8, 8), torch.zeros(1).long())


def test_ema_host_math_matches_reference_golden(tmp_path):
    """src/ema.py:10-114 — std<->gamma, beta schedule, post-hoc weights and reconstruction from fp16 snapshots."""
    from mapdit_b200 import ema as E
    g = np.load(os.path.join(GOLDEN, "ema.npz"))
    np.testing.assert_allclose(E.std_to_gamma(g["stds"]), g["gammas"], rtol=1e-12)
    np.testing.assert_allclose(E.gamma_to_std(g["gammas"]), g["back"], rtol=1e-12)
    for i, s in enumerate((0.05, 0.1)):
        np.testing.assert_allclose([E.calc_beta(s, t) for t in g["beta_ts"]], g["betas"][i], rtol=1e-12)
    w = E.solve_weights(g["snap_ts"], E.std_to_gamma(g["snap_stds"]), 1000, E.std_to_gamma(0.075))
    np.testing.assert_allclose(w, g["weights"], rtol=1e-10)
    for i, (s, t) in enumerate(zip(g["snap_stds"], g["snap_ts"])):
        sd = {"a.weight": torch.from_numpy(g[f"snap{i}_a"]).half(), "b": torch.from_numpy(g[f"snap{i}_b"]).half()}
        torch.save({"std": float(s), "t": int(t), "state_dict": sd}, tmp_path / f"{s:.3f}_{int(t):07d}.pt")
    out = E.calculate_posthoc_ema(0.075, str(tmp_path), verbose=False)
    np.testing.assert_allclose(out["a.weight"].numpy(), g["posthoc_a"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out["b"].numpy(), g["posthoc_b"], rtol=1e-6, atol=1e-7)
    hit = E.calculate_posthoc_ema(0.1, str(tmp_path), verbose=False)
    assert hit["a.weight"].dtype == torch.float16  # exact-std hit returns the stored fp16 snapshot (src/ema.py:93-98)
    np.testing.assert_array_equal(hit["a.weight"].float().numpy(), g["exact_a"])


def test_lr_schedule_and_experiment_layout_match_reference(tmp_path):
    """train.py:179-214: LambdaLR factor (golden from the reference's own function) and the NNN-Model/checkpoints layout"""
    from mapdit_b200 import data
    g = np.load(os.path.join(GOLDEN, "ema.npz"))
    for row, (w, d) in zip(g["lr_factors"], ((1000, 20000), (1, 10), (100, 100))):
        lam = data.create_lr_lambda(w, d)
        np.testing.assert_allclose([lam(int(s)) for s in g["lr_steps"]], row, rtol=1e-15)
    e0 = data.setup_experiment("DiT-B/2", str(tmp_path))
    e1 = data.setup_experiment("DiT-B/2", str(tmp_path))
    assert os.path.basename(e0) == "000-DiT-B-2" and os.path.basename(e1) == "001-DiT-B-2"
    assert os.path.isdir(os.path.join(e1, "checkpoints"))
    data.save_config(e0, dict(model="DiT-B/2", lr=1e-2, stats_mean=[0.1, 0.2]))
    import yaml
    assert yaml.safe_load(open(os.path.join(e0, "config.yaml")))["model"] == "DiT-B/2"


def test_state_dict_accepts_compiled_prefix():
    """checkpoints of the reference carry `_orig_mod.` keys (train.py:46,124-128)"""
    import mapdit_b200 as M
    cfg = O.config_for("DiT-XS/8")
    sd = O.init_state_dict(cfg, seed=3)
    m = M.DIT_MODELS["DiT-XS/8"](in_channels=4, input_size=32, num_classes=1000)
    m.load_state_dict({"_orig_mod." + k: v for k, v in sd.items()})
    assert all(torch.equal(m.state_dict()[k], v) for k, v in sd.items())


def test_bench_flops_model_matches_survey_table():
    """bench.py's algorithmic FLOPs per image (the numerator of `step_frac_of_sustained`) against SURVEY.md §8(d): DiT-B/2 forward
    46.011 GFLOP, training 139.240 GFLOP; the rotation-and-scaling headline has a 5D-wide modulation GEMM instead of 6D."""
    import importlib.util
    import os
    import types
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    # bench.py re-points file descriptor 1 at import time (stdout guard): give it a scratch descriptor to play with
    saved = os.dup(1)
    try:
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    m = types.SimpleNamespace(patch_size=2, in_channels=4, input_size=32, depth=12, hidden_size=768, modulation="adaln")
    fl = bench.flops_per_image(m)
    assert abs(fl["fwd"] / 1e9 - 46.011) < 0.01 and abs(fl["train"] / 1e9 - 139.240) < 0.02
    m.modulation = "rotation_scaling"
    fr = bench.flops_per_image(m)
    assert fr["fwd"] == fl["fwd"] - 12 * 2 * 768 * 768  # one D x D slice less per block
    assert bench.metric_name("train") == "dit_b2_map_train_img_per_s" and "rotation-and-scaling" in bench.workload_name("train")


def test_gemm_args_struct_mirror_matches_library():
    """the ctypes mirror of mapdit_gemm_args (with the ldrot field of the rotation epilogue) has the library's size"""
    import ctypes as C
    from mapdit_b200 import _lib
    L = _lib.lib()
    L.mapdit_sizeof_gemm_args.restype = C.c_int
    assert L.mapdit_sizeof_gemm_args() == C.sizeof(_lib.GemmArgs)
    assert _lib.EPI_RESID_ROT == 6


def test_launch_summary_tool_reads_the_committed_profile():
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "launch_summary.py"), os.path.join(root, "profiles", "r2_launches_train.csv")],
                         capture_output=True, text=True, check=True).stdout
    assert "360 launches" in out and "gemm_tc2_kernel<256>" in out and "attn_bwd_fused2_tc" in out


def test_epilogue_and_variant_constants_match_header():
    """the Python mirror of the header's enums (epilogue selectors, dtype codes) cannot drift from include/mapdit.h"""
    from mapdit_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mapdit.h")).read()
    defs = {k: int(v, 0) for k, v in re.findall(r"#define\s+(MAPDIT_[A-Z0-9_]+)\s+(-?(?:0x[0-9a-fA-F]+|\d+))\b", hdr)}
    for name in ("STORE", "QKNORM", "MPSILU", "RESID_MOD", "RESID", "SILU_BWD", "RESID_ROT", "STORE_DELTA"):
        assert defs[f"MAPDIT_EPI_{name}"] == getattr(_lib, f"EPI_{name}"), name
    assert len({defs[k] for k in defs if k.startswith("MAPDIT_EPI_")}) == 8  # no two selectors share a value


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference`: the reference's own CPU implementation (oracle/_ref, the unmodified reference mirrored by
    oracle/vendor_reference.py) when present, else the pinned restatement; one JSON line with the contract's keys"""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "DiT-XS/8", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "src", "models.py"))
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["sample50"]["value"] > 0 and d["sample50"]["cpu_baseline"]["kind"] == d["cpu_baseline"]["kind"]


def test_gradient_bucket_plan():
    """TrainStep's reduce buckets: every group in exactly one bucket, buckets are runs of consecutive groups in backward order, a
    bucket's last group is the one whose completion launches its all-reduce"""
    from mapdit_b200.parallel import bucket_plan
    for ng in (1, 2, 3, 13, 29):
        for bpb in (0, 1, 2, 3, 6, 7, 12, 100):
            bog, last = bucket_plan(ng, bpb)
            assert len(bog) == ng and bog[0] == 0 and all(0 <= bog[i + 1] - bog[i] <= 1 for i in range(ng - 1)), (ng, bpb, bog)
            assert last == [max(i for i in range(ng) if bog[i] == b) for b in range(max(bog) + 1)] and last[-1] == ng - 1
            if bpb <= 0 or ng == 1:
                assert max(bog) == 0
            elif bpb == 1:
                assert bog == list(range(ng))  # the historical layout: a bucket per block, the embedders / final layer on their own
            else:
                assert all(bog.count(b) <= bpb + 1 for b in set(bog)) and bog[-1] == bog[-2]  # the last group joins the last bucket
    assert bucket_plan(13, 6) == ([0] * 6 + [1] * 7, [5, 12])
