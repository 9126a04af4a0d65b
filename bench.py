#!/usr/bin/env python
"""Benchmark of the MaP-DiT hot path on B200 (contract: task prompt; metric: BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sample|train|forward] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of synthetic
latents: `sample` = one 50-step respaced p_sample_loop over the batch (model.forward, no CFG),
`train` = training_losses + backward + Adam, `forward` = one eval forward.  Workload at every N is
DiT-B/2 (BASELINE.json configs[2]; the configuration the metric is quoted on), batch 256 per GPU
(weak scaling, batch-sharded, no data-path collective except the final gather of samples).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  NCCL prints its version banner to file descriptor 1 from C (NCCL_DEBUG_FILE does not
# cover it), so fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit_line(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


MODEL = "DiT-B/2"
SAMPLING_STEPS = 50
# BASELINE.json configs[2] / north_star target: "DiT-B/2 MaP (all mp flags + rotation-and-scaling modulation)".  The reference
# snapshot ships only the MP-AdaLN modulation ("adaln", the parity-pinned variant): it is timed too and reported in `adaln`.
MODULATION = "rotation_scaling"
MOD_NAMES = {"adaln": "MP-AdaLN modulation (reference snapshot)", "rotation_scaling": "rotation-and-scaling modulation",
             "rotation": "rotation modulation"}


def flops_per_image(model):
    """algorithmic FLOPs (2*MAC) of one forward, split for the train multiplier."""
    p, C = model.patch_size, model.in_channels
    T, L, D = (model.input_size // p) ** 2, model.depth, model.hidden_size
    lin = T * L * 24 * D * D
    attn = T * L * 4 * T * D
    emb = T * (2 * (p * p * C + 1) * D + 2 * D * 2 * p * p * C)
    modw = {"adaln": 6, "rotation_scaling": 5, "rotation": 3}[getattr(model, "modulation", "adaln")]  # modulation GEMM width / D
    cond = L * 2 * modw * D * D + 4 * D * D + 2 * (256 * D + D * D) + 32 * D
    return dict(fwd=lin + attn + emb + cond, train=3 * (lin + emb + cond) + 3.5 * attn)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], src="measured")
    except Exception:
        return dict(tf_burst=1590.0, tf_sustained=1400.0, hbm=6650.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args, workload=None, emit=True):
    """The reference arm: the reference's own CPU implementation of the path on all host cores, on a bounded sample of the same
    workload.  When `oracle/_ref` is there (the UNMODIFIED reference mirrored by oracle/vendor_reference.py: it travels to the GPU
    box with the snapshot) that is the reference ITSELF through its public API — `DIT_MODELS[...]`, `create_diffusion`,
    `training_losses` + `torch.optim.Adam` as in train.py:86-96, `p_sample` as in sample.py:52-61 — with its MP-AdaLN modulation (the
    only one it has code for), `cpu_baseline.kind = "reference"`; otherwise the pinned restatement (oracle/, kind "port")."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    workload = workload or args.workload
    if workload == "both":
        line = run_reference(args, "train", emit=False)
        sl = run_reference(args, "sample", emit=False)
        line["metric"] = "dit_b2_map_train_img_per_s (+ sample50 img/s in `sample50`)"
        line["sample50"] = {k: sl[k] for k in ("value", "unit", "ms_per_step", "cpu_baseline", "e2e")}
        emit_line(line)
        return line
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, S = 8, args.input_size
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, S, S, generator=g)
    y = torch.randint(0, 1000, (B,), generator=g)
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    use_ref = os.path.isfile(os.path.join(ref_dir, "src", "models.py")) and not args.flags_off
    times = []
    nsub = 2 if workload == "sample" else 1
    if use_ref:
        kind, who, modulation = "reference", "the unmodified reference (oracle/_ref) on the host cores", "adaln (the reference has no other)"
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        from diffusion import create_diffusion as ref_create_diffusion  # the reference's packages
        from src.models import DIT_MODELS as REF_MODELS
        torch.manual_seed(0)
        model = REF_MODELS[MODEL](in_channels=4, input_size=S, num_classes=1000)
        if workload == "train":
            diffusion = ref_create_diffusion(timestep_respacing="")
            opt = torch.optim.Adam(model.parameters(), lr=1e-2, betas=(0.9, 0.99))
            model.train()

            def step():  # train.py:86-96
                t = torch.randint(0, diffusion.num_timesteps, (B,))
                loss = diffusion.training_losses(model, x, t, dict(y=y))["loss"].mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
        else:
            diffusion = ref_create_diffusion(str(SAMPLING_STEPS))
            model.eval()

            def step():  # the first `nsub` iterations of p_sample_loop (sample.py:52-61 without CFG), or one eval forward
                img = x
                with torch.no_grad():
                    for k in range(nsub):
                        tt = torch.full((B,), diffusion.num_timesteps - 1 - k, dtype=torch.long)
                        if workload == "sample":
                            img = diffusion.p_sample(model.forward, img, tt, clip_denoised=True, model_kwargs=dict(y=y))["sample"]
                        else:
                            model(img, tt, y)
    else:
        from oracle import mapdit_oracle as O
        kind, who, modulation = "port", "the pinned CPU restatement (oracle/mapdit_oracle.py)", MODULATION
        cfg = O.config_for(MODEL, modulation=MODULATION, input_size=S)
        sd = O.init_state_dict(cfg, seed=0)
        if workload == "train":
            p = O.make_params(sd)
            T = O.make_tables("")
            opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-2, betas=(0.9, 0.99))
            t = torch.randint(0, 1000, (B,), generator=g)
            noise = torch.randn(B, 4, S, S, generator=g)

            def step():
                opt.zero_grad()
                O.train_step_grads(p, cfg, T, x, t, y, noise)
                opt.step()
        else:
            T = O.make_tables(str(SAMPLING_STEPS))
            tm = torch.tensor(T.timestep_map)

            def step():
                img = x
                with torch.no_grad():
                    for k in range(nsub):
                        i = T.num_timesteps - 1 - k
                        tt = torch.full((B,), i, dtype=torch.long)
                        out = O.dit_forward(sd, cfg, img, tm[tt], y)
                        if workload == "sample":
                            img = O.p_sample_step(T, out, img, tt, torch.randn_like(img))["sample"]
    sample = {"train": f"{MODEL} training step (training_losses + backward + Adam) at batch {B}",
              "sample": f"{MODEL} {nsub} of {SAMPLING_STEPS} sampling steps at batch {B}, scaled to {SAMPLING_STEPS} steps",
              "forward": f"{MODEL} eval forward at batch {B}"}[workload] + f": {who}"
    per_step_images = B * nsub / SAMPLING_STEPS if workload == "sample" else B
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        step()
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = per_step_images / (ms / 1e3)
    unit = "img/s"
    line = {"impl": "reference", "metric": metric_name(workload), "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(workload), "model": MODEL, "modulation": modulation, "batch_per_step": B,
                       "note": "bounded sample of the workload at batch 8 on the host cores; img/s is per image, so it compares with the GPU arm's"},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if emit:
        emit_line(line)
    return line


def metric_name(w):
    return {"sample": "dit_b2_map_sample50_img_per_s", "train": "dit_b2_map_train_img_per_s", "forward": "dit_b2_map_forward_img_per_s"}[w]


def workload_name(w):
    mp = f"{MODEL} MaP + {MOD_NAMES[MODULATION]}"
    return {"sample": f"{mp}, {SAMPLING_STEPS}-step respaced p_sample_loop, 32x32x4 latents, batch 256/GPU, no CFG",
            "train": f"{mp} training step (training_losses + backward + Adam), 32x32x4 latents, batch 256/GPU",
            "forward": f"{mp} eval forward, 32x32x4 latents, batch 256/GPU"}[w]


# --------------------------------------------------------------------------------------------- our arm (B200)
VAL_TRAIN_N, VAL_SAMPLE_N = 8, 4  # sub-batch sizes of the oracle validation legs


def make_model(M, dtype, dev, input_size, flags_off):
    """the bench's synthetic network: torch seed 0, the reference's init distributions, gains / sigma reference moved off their
    zero init like a trained net (with the reference's zero gains the shift / rotation paths would carry no signal)"""
    import torch
    torch.manual_seed(0)  # same replica on every rank (and on the CPU for the oracle legs)
    off = {k: False for k in ("use_cosine_attention", "use_weight_normalization", "use_forced_weight_normalization", "use_mp_residual",
                              "use_mp_silu", "use_no_layernorm", "use_mp_pos_enc", "use_mp_embedding")} if flags_off else {}
    model = M.DIT_MODELS[MODEL](in_channels=4, input_size=input_size, num_classes=1000, compute_dtype=dtype, modulation=MODULATION, **off)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if prm.dim() == 0:
                prm.fill_(0.3)
        model.final_layer.sigma_scale.reference.normal_()
    return model.to(dev) if dev is not None else model, off


def host_inputs(B, S, rank):
    """seeded host-side inputs of one rank (pinned by the caller)"""
    import torch
    g = torch.Generator().manual_seed(1 + rank)
    z = torch.randn(B, 4, S, S, generator=g)
    y = torch.randint(0, 1000, (B,), generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    noise = torch.randn(B, 4, S, S, generator=g)
    drop = torch.rand(B, generator=g) < 0.1
    step_noises = torch.randn(SAMPLING_STEPS, VAL_SAMPLE_N, 4, S, S, generator=g)
    return dict(z=z, y=y, t=t, noise=noise, drop=drop, step_noises=step_noises)


def cpu_legs(args, workloads, baseline=True):
    """Rank 0, N = 1, BEFORE any NCCL initialisation (other ranks must not spin on the GPU while the host cores are timed): the
    pinned CPU oracle (oracle/mapdit_oracle.py) on bounded samples of the same workload with the bench's own weights and inputs.
    It returns the reported `cpu_baseline` AND the reference values the GPU results are validated against afterwards."""
    import torch
    import mapdit_b200 as M
    from oracle import mapdit_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, off = make_model(M, "fp32", None, args.input_size, args.flags_off)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    del model
    cfg = O.config_for(MODEL, modulation=MODULATION, input_size=args.input_size, **off)
    hi = host_inputs(args.batch, args.input_size, 0)
    out = {}
    if "train" in workloads:
        n = VAL_TRAIN_N
        p = O.make_params(sd)
        T = O.make_tables("")
        x, t, y, noise, drop = (hi[k][:n] for k in ("z", "t", "y", "noise", "drop"))
        terms, _ = O.train_step_grads(p, cfg, T, x, t, y, noise, drop_mask=drop)  # step-0 per-sample losses (also the warm-up)
        rec = {"loss_ref": terms["loss"].detach().clone()}
        if baseline:
            opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-2, betas=(0.9, 0.99))
            k, t0 = 0, time.perf_counter()
            while True:
                opt.zero_grad()
                O.train_step_grads(p, cfg, T, x, t, y, noise, drop_mask=drop)
                opt.step()
                k += 1
                el = time.perf_counter() - t0
                if el > 15.0 or k >= 10:
                    break
            rec["cpu_baseline"] = {"value": n * k / el, "unit": "img/s", "cores": cores, "kind": "port",
                                   "sample": f"{MODEL} training step (loss + backward + Adam) at batch {n}, {k} iterations in {el:.1f} s "
                                             "on the pinned CPU oracle (oracle/mapdit_oracle.py)"}
        out["train"] = rec
    if "sample" in workloads:
        n = VAL_SAMPLE_N
        T = O.make_tables(str(SAMPLING_STEPS))
        z, y = hi["z"][:n], hi["y"][:n]
        t0 = time.perf_counter()
        ref = O.p_sample_loop(T, lambda a, b: O.dit_forward(sd, cfg, a, b, y), z, list(hi["step_noises"]), clip_denoised=True)
        el = time.perf_counter() - t0
        out["sample"] = {"sample_ref": ref,
                         "cpu_baseline": {"value": n / el, "unit": "img/s", "cores": cores, "kind": "port",
                                          "sample": f"{MODEL} full {SAMPLING_STEPS}-step p_sample_loop of {n} images in {el:.1f} s on the "
                                                    "pinned CPU oracle (oracle/mapdit_oracle.py)"}}
    return out


def time_kernel(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return sum(ev[i].elapsed_time(ev[i + 1]) for i in range(iters)) / iters  # ms


def kernel_roofline(model, B, pk):
    """Isolated timing (CUDA events on the launching stream) of the dominant kernel: the fc1 block GEMM
    (gemm_tc2_kernel<256>: cta_group::2 256 x 256 tiles, fused mp_silu epilogue), plus the other block GEMMs and attention."""
    import torch
    from mapdit_b200 import _lib, ops
    D, T, H = model.hidden_size, (model.input_size // model.patch_size) ** 2, model.num_heads
    M = B * T
    dev = "cuda"
    mk = lambda *s: (torch.randn(*s, device=dev) * 0.05).bfloat16()
    h, u4 = mk(M, D), mk(M, 4 * D)
    wqkv, wo, w1, w2 = mk(3 * D, D), mk(D, D), mk(4 * D, D), mk(D, 4 * D)
    qkv, o, x = mk(M, 3 * D), mk(M, D), mk(M, D)
    mods = torch.randn(B, 6 * D, device=dev)
    gain = torch.tensor(0.3, device=dev)
    res = {}
    rot = getattr(model, "modulation", "adaln") != "adaln"
    cs = torch.empty(B, D, device=dev)
    ops.rot_table(mods[:, D:], gain, cs, 6 * D, D)

    def resid_mod(a, w):  # residual + next modulation fused into the GEMM epilogue (rotation table or shift/scale/gain)
        if rot:
            return ops.gemm_bf16(a, w, x, epilogue=_lib.EPI_RESID_ROT, out2=h, resid=x, gate=mods, shift=cs, scale=mods[:, 2 * D:],
                                 ldmod=6 * D, ldrot=D, tokens=T)
        return ops.gemm_bf16(a, w, x, epilogue=_lib.EPI_RESID_MOD, out2=h, resid=x, gate=mods, shift=mods[:, D:], scale=mods[:, 2 * D:],
                             gain=gain, ldmod=6 * D, tokens=T)
    # attention backward (delta pre-kernel + the fused kernel of the 256-token models, or the dq / dkv kernel pair): 5 GEMMs of algorithmic work
    dO, dqkv = mk(M, D), torch.empty(M, 3 * D, device=dev, dtype=torch.bfloat16)
    lse, delta = torch.empty(M, H, device=dev), torch.empty(M, H, device=dev)
    ops.cos_attn(qkv, o, B, T, H, D // H, lse=lse)
    o_saved = o.clone()
    cases = {
        "qkv_gemm_qknorm": (lambda: ops.gemm_bf16(h, wqkv, qkv, epilogue=_lib.EPI_QKNORM, tokens=T, head_dim=D // H, qk_cols=2 * D), 2 * M * D * 3 * D),
        "attn": (lambda: ops.cos_attn(qkv, o, B, T, H, D // H), 4 * M * T * D),
        "attn_bwd": (lambda: ops.cos_attn_bwd(qkv, o_saved, dO, lse, dqkv, delta, B, T, H, D // H), 10 * M * T * D),
        "out_gemm_resid_mod": (lambda: resid_mod(o, wo), 2 * M * D * D),
        "fc1_gemm_mpsilu": (lambda: ops.gemm_bf16(h, w1, u4, epilogue=_lib.EPI_MPSILU), 2 * M * D * 4 * D),
        "fc2_gemm_resid_mod": (lambda: resid_mod(u4, w2), 2 * M * 4 * D * D),
    }
    for k, (fn, fl) in cases.items():
        ms = time_kernel(fn)
        res[k] = {"ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1)}
    dom = res["fc1_gemm_mpsilu"]
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this shape, from the committed ncu --set full
    # capture of the shipped kernel (tools/ncu_traffic.py writes the file from the .ncu-rep); null when the capture is of another shape
    traffic, tsrc = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "round2_fc1_gemm_tc2_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("m") == M and tj.get("n") == 4 * D and tj.get("k") == D and "gemm_tc2_kernel" in tj.get("kernel", ""):
            traffic, tsrc = tj["dram_bytes_read"] + tj["dram_bytes_write"], "B/launch (ncu --set full, profiles/round2_fc1_gemm_tc2_traffic.json)"
    except Exception:
        pass
    algo_bytes = 2 * (M * D + 4 * D * D + M * 4 * D)  # A + B + out in bf16 (the eval flavour timed here; training also writes the pre-activation)
    roof = {"bound": "tensor", "kernel": "gemm_tc2_kernel<256> (fc1, cta_group::2, fused mp_silu epilogue), M=%d N=%d K=%d" % (M, 4 * D, D),
            "achieved": dom["tflops"], "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": round(dom["tflops"] / pk["tf_burst"], 4),
            "peak_source": f"MEASURED_PEAKS.json bf16 burst ({pk['src']})",
            "traffic": traffic, "traffic_unit": tsrc, "algorithmic_bytes": algo_bytes, "per_kernel": res}
    return roof


def run_ours(args, workload, finalize=True, cpu=None):
    """`cpu` = cpu_legs()[workload] (rank 0 at N = 1): the oracle's reference values + the reported cpu_baseline"""
    import torch
    import torch.distributed as dist
    import mapdit_b200 as M
    from mapdit_b200 import _lib
    from mapdit_b200.diffusion import gaussian_diffusion as gd
    if args.gemm_2cta is not None:
        _lib.set_option("gemm_2cta", args.gemm_2cta)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    B = args.batch
    S = args.input_size
    model, _ = make_model(M, args.dtype, dev, S, args.flags_off)
    fl = flops_per_image(model)
    pk = peaks()
    # host-side (pinned) inputs for the e2e leg, device-resident copies for the kernel-only leg
    hi = host_inputs(B, S, rank)
    z_host, y_host, t_host = hi["z"].pin_memory(), hi["y"].pin_memory(), hi["t"].pin_memory()
    z_dev, y_dev, t_dev = z_host.to(dev), y_host.to(dev), t_host.to(dev)
    out_host = torch.empty(B, 4, S, S).pin_memory()
    diffusion = M.create_diffusion(str(SAMPLING_STEPS) if workload == "sample" else "")
    validation = {}
    state = {}

    if workload == "sample":
        model.eval()
        gather = [torch.empty(B, 4, S, S, device=dev) for _ in range(world)] if world > 1 else None
        # clip_denoised=True: on random weights the unclipped x0 prediction of sample.py:52-61 (clip_denoised=False on TRAINED
        # weights) diverges to inf within ~37 steps; the clamp is one instruction of the fused step kernel, the work is identical
        CLIP = True

        def step_dev():
            s = diffusion.p_sample_loop(model.forward, z_dev.shape, z_dev, clip_denoised=CLIP, model_kwargs=dict(y=y_dev), device=dev)
            if world > 1:
                dist.all_gather(gather, s)
            state["last"] = s
            return s

        def step_e2e():
            z = z_host.to(dev, non_blocking=True)
            y = y_host.to(dev, non_blocking=True)
            s = diffusion.p_sample_loop(model.forward, z.shape, z, clip_denoised=CLIP, model_kwargs=dict(y=y), device=dev)
            if world > 1:
                dist.all_gather(gather, s)
            out_host.copy_(s, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        images_per_step = B
        flops_step = fl["fwd"] * B * SAMPLING_STEPS
        h2d, d2h = z_host.numel() * 4 + y_host.numel() * 8, out_host.numel() * 4
        if cpu is not None and rank == 0:
            # end-of-loop divergence of a sub-batch against the CPU oracle on the same weights, start noise and per-step noise
            n = VAL_SAMPLE_N
            it = iter(hi["step_noises"].to(dev))
            real = gd._randn_like
            gd._randn_like = lambda v: next(it)
            try:
                sv = diffusion.p_sample_loop(model.forward, (n, 4, S, S), z_dev[:n].clone(), clip_denoised=True,
                                             model_kwargs=dict(y=y_dev[:n]), device=dev)
            finally:
                gd._randn_like = real
            ref = cpu["sample_ref"].double()
            validation["sample_rel_l2_vs_oracle"] = float((sv.cpu().double() - ref).norm() / ref.norm())
            validation["sample_check"] = (f"{n} images, {SAMPLING_STEPS} steps, shared noise, free running, clip_denoised=True, {args.dtype} vs the "
                                          "fp32 CPU oracle" + ("" if MODULATION == "adaln" else " (rotation modulation: self-referential oracle)"))
    elif workload == "forward":
        model.eval()

        def step_dev():
            with torch.no_grad():
                state["last"] = model(z_dev, t_dev, y_dev)
                return state["last"]

        out8 = torch.empty(B, 8, S, S).pin_memory()

        def step_e2e():
            with torch.no_grad():
                o = model(z_host.to(dev, non_blocking=True), t_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True))
            out8.copy_(o, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        images_per_step = B
        flops_step = fl["fwd"] * B
        h2d, d2h = z_host.numel() * 4 + 2 * B * 8, out8.numel() * 4
    else:
        from mapdit_b200.train import TrainStep
        model.train()
        if args.wgrad_stream is not None:
            model.engine.trainer.wgrad_stream = bool(args.wgrad_stream)
        ts = TrainStep(model, diffusion, lr=1e-2, betas=(0.9, 0.99), world_size=world)
        noise_host, drop_host = hi["noise"].pin_memory(), hi["drop"].pin_memory()
        noise_dev, drop_dev = noise_host.to(dev), drop_host.to(dev)
        losses = []
        if cpu is not None and rank == 0:
            # step-0 per-sample losses of a sub-batch against the CPU oracle (same weights, inputs, noise, label-dropout mask)
            n = VAL_TRAIN_N
            l0 = ts.compute_grads(z_dev[:n], t_dev[:n], y_dev[:n], noise_dev[:n], drop_dev[:n], reduce=False)
            ref = cpu["loss_ref"].double()
            validation["loss_rel_l2_vs_oracle"] = float((l0.cpu().double() - ref).norm() / ref.norm())
            validation["loss_check"] = (f"per-sample losses of {n} samples before the first optimiser step, {args.dtype} vs the fp32 CPU oracle"
                                        + ("" if MODULATION == "adaln" else " (rotation modulation: self-referential oracle)"))

        def step_dev():
            loss = ts.step(z_dev, t_dev, y_dev, noise_dev, drop_mask=drop_dev)
            losses.append(loss)
            return loss

        def step_e2e():
            loss = ts.step(z_host.to(dev, non_blocking=True), t_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True),
                           noise_host.to(dev, non_blocking=True), drop_mask=drop_host.to(dev, non_blocking=True))
            return float(loss)  # D2H read of the loss, like train.py:99
        images_per_step = B
        flops_step = fl["train"] * B
        h2d, d2h = (z_host.numel() + noise_host.numel()) * 4 + 2 * B * 8 + B, 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        n0 = _lib.total_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = _lib.total_launches() - n0
        clk = sampler.stop() if sampler else None
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms, launches, clk

    # the dominant kernel is timed alone (against the burst peak), so it is timed first: after the long step loops the GPU sits at
    # its power-capped clock and a stand-alone kernel would be compared with the burst peak in the sustained state
    roof_early = kernel_roofline(model, B, pk) if (rank == 0 and not args.no_roofline) else None
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    warm = max(args.warmup, 3)
    ms, launches, clk = timed(step_dev, args.steps, warm, sampler)
    # self-validation of the timed region's own results (every rank; reported by rank 0)
    if workload == "train":
        lv = torch.stack(losses).float().cpu()
        validation.update(loss_first=float(lv[0]), loss_first_timed=float(lv[warm]), loss_last=float(lv[-1]),
                          losses_finite=bool(torch.isfinite(lv).all()),
                          loss_note="the same synthetic batch every step: the loss must fall from loss_first (step 0, before any update)")
        fin = bool(torch.isfinite(lv).all())
    else:
        fin = bool(torch.isfinite(state["last"]).all())
        validation.update(finite=fin, out_abs_max=float(state["last"].abs().max()))
    if world > 1:
        ft = torch.tensor([1 if fin else 0], device=dev)
        dist.all_reduce(ft, op=dist.ReduceOp.MIN)
        fin = bool(int(ft))
    validation["finite_all_ranks"] = fin
    ms_e2e, _, _ = timed(step_e2e, args.steps, 1)
    value = images_per_step * world * args.steps / (ms / 1e3)
    e2e_value = images_per_step * world * args.steps / (ms_e2e / 1e3)
    if rank == 0:
        roof = roof_early
        if roof is not None:
            step_tf = flops_step * args.steps / (ms / 1e3) / 1e12
            roof["step_tflops_per_gpu"] = round(step_tf, 1)
            roof["step_frac_of_sustained"] = round(step_tf / pk["tf_sustained"], 4)
        line = {"metric": metric_name(workload), "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": workload_name(workload), "model": MODEL, "modulation": MODULATION, "batch_per_gpu": B,
                           "global_batch": B * world,
                           "parallelism": f"dp{world}" if world > 1 else "single",
                           "l2": "per-step working set (activations ~100 MB per [M,D] tensor, 2.4 GB per block) exceeds the 126 MB L2; no flush needed",
                           "weights": "random init (torch seed 0), reference init distributions, gains 0.3", "clip_denoised": True},
                "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clk, "roofline": roof, "validation": validation,
                "cpu_baseline": cpu.get("cpu_baseline") if cpu is not None else None}
    else:
        line = None
    del model
    torch.cuda.empty_cache()
    if world > 1 and finalize:
        dist.destroy_process_group()
    return line


def ref_on_b200(args):
    """informational: the unmodified reference (oracle/_ref) through its own PyTorch path on this GPU (rank 0, N = 1).  The eager
    arms are timed live; the torch.compile arms (minutes of compilation) come from the committed run of tools/ref_on_b200.py."""
    out = {"what": "unmodified reference DiT-B/2 (MP-AdaLN) on this B200 through its own PyTorch path; compare with `adaln`",
           "batch": args.batch}
    try:
        with open(os.path.join(ROOT, "profiles", "round2_ref_on_b200.json")) as f:
            out["recorded"] = {"source": "profiles/round2_ref_on_b200.json (tools/ref_on_b200.py --compile on a B200 of this pool)", **json.load(f)}
    except Exception:
        out["recorded"] = None
    if args.no_ref_on_b200:
        return out
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import ref_on_b200 as R
        out["live"] = R.measure(MODEL, args.batch, arms=("eager_tf32", "eager_bf16_autocast"), train_steps=3, sampling_steps=SAMPLING_STEPS,
                                input_size=args.input_size)
    except Exception as e:  # noqa: BLE001
        out["live"] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MAPDIT_BENCH_WORKLOAD", "both"), choices=["both", "sample", "train", "forward"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--gemm-2cta", type=int, default=None, help="override the GEMM kernel choice: 1 = cta_group::2 256xBN tiles, 0 = 1-CTA 128xBN")
    ap.add_argument("--wgrad-stream", type=int, default=None, help="A/B: 0 = weight-gradient GEMMs on the main stream, 1 = on a second stream (default)")
    ap.add_argument("--model", default=None, help="other BASELINE.json configs (e.g. DiT-S/2, DiT-L/2, DiT-XL/2); default DiT-B/2")
    ap.add_argument("--input-size", type=int, default=32, help="latent size (64 for BASELINE config 5)")
    ap.add_argument("--sampling-steps", type=int, default=None, help="respaced steps of the sample workload (default 50)")
    ap.add_argument("--flags-off", action="store_true", help="AdaLN baseline of BASELINE config 4: every --use-* MaP switch off")
    ap.add_argument("--modulation", default=None, choices=["adaln", "rotation_scaling", "rotation"],
                    help="default rotation_scaling (BASELINE configs[2]); adaln = the reference snapshot's MP-AdaLN")
    ap.add_argument("--no-adaln-arm", action="store_true", help="skip the extra MP-AdaLN timing of the default run")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-on-b200", action="store_true", help="skip the live eager timing of the unmodified reference on the GPU")
    args = ap.parse_args()
    global MODEL, SAMPLING_STEPS, MODULATION
    if args.modulation:
        MODULATION = args.modulation
    if args.flags_off:
        MODULATION = "adaln"
    if args.model:
        MODEL = args.model
    if args.sampling_steps:
        SAMPLING_STEPS = args.sampling_steps
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    rank = int(os.environ.get("RANK", "0"))
    # CPU legs (oracle reference values + the reported cpu_baseline): rank 0 at N = 1 only, before anything touches NCCL or the GPU
    do_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    wl = ["train", "sample"] if args.workload == "both" else ([args.workload] if args.workload in ("train", "sample") else [])
    cpu = cpu_legs(args, wl) if (do_cpu and wl) else {}
    if args.workload == "both":
        # BASELINE.json's metric has two halves: the training step is the primary value, the 50-step sampler rides along
        cpu_adaln = {}
        if do_cpu and MODULATION != "adaln" and not args.no_adaln_arm:
            main_mod, MODULATION = MODULATION, "adaln"
            cpu_adaln = cpu_legs(args, wl, baseline=False)
            MODULATION = main_mod
        refb = ref_on_b200(args) if (rank == 0 and world == 1 and not args.flags_off and args.model is None) else None
        line = run_ours(args, "train", finalize=False, cpu=cpu.get("train"))
        sline = run_ours(args, "sample", finalize=False, cpu=cpu.get("sample"))
        adaln = None
        if MODULATION != "adaln" and not args.no_adaln_arm:
            # the parity-pinned variant (the only modulation the reference snapshot has code for), device-timed only
            main_mod, MODULATION = MODULATION, "adaln"
            args.no_roofline = True
            at = run_ours(args, "train", finalize=False, cpu=cpu_adaln.get("train"))
            asmp = run_ours(args, "sample", finalize=False, cpu=cpu_adaln.get("sample"))
            MODULATION = main_mod
            if at is not None:
                adaln = {"modulation": "adaln", "note": "reference snapshot's MP-AdaLN modulation (parity pinned by the reference)",
                         "train": {k: at[k] for k in ("value", "unit", "ms_per_step", "gpu_launches", "validation")},
                         "sample50": {k: asmp[k] for k in ("value", "unit", "ms_per_step", "gpu_launches", "validation")},
                         "e2e": {"train": at["e2e"], "sample50": asmp["e2e"]}}
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
        if line is not None:
            if adaln is not None:
                line["adaln"] = adaln
            line["metric"] = "dit_b2_map_train_img_per_s (+ sample50 img/s in `sample50`)"
            if refb is not None:
                line["ref_on_b200"] = refb
            keep = ("value", "unit", "ms_per_step", "e2e", "gpu_launches", "clocks", "cpu_baseline", "validation")
            line["sample50"] = {k: sline[k] for k in keep}
            line["sample50"]["metric"] = sline["metric"]
            line["sample50"]["workload"] = sline["config"]["workload"]
            if sline.get("roofline"):
                line["sample50"]["step_tflops_per_gpu"] = sline["roofline"]["step_tflops_per_gpu"]
                line["sample50"]["step_frac_of_sustained"] = sline["roofline"]["step_frac_of_sustained"]
    else:
        line = run_ours(args, args.workload, cpu=cpu.get(args.workload))
    if line is not None:
        emit_line(line)


if __name__ == "__main__":
    main()
