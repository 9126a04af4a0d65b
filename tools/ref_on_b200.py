#!/usr/bin/env python
"""The same-box bar: the UNMODIFIED reference (oracle/_ref, mirrored by oracle/vendor_reference.py) timed on the B200 through its
own PyTorch path — informational, never on the product path.

    python tools/ref_on_b200.py [--compile] [--model DiT-B/2] [--batch 256] [--out gpurun_out/ref_on_b200.json]

Arms (what a user of the reference gets on this GPU without this repo):
  eager_tf32            train.py as written minus torch.compile: fp32 parameters/activations, TF32 matmuls (train.py:222-223)
  eager_bf16_autocast   the same under torch.autocast(bfloat16)
  compile_tf32          + torch.compile(model) (train.py:46, sample.py:25) — the reference's actual configuration
  compile_bf16_autocast both
Workloads: the training step of train.py:86-96 (training_losses -> mean -> zero_grad -> backward -> Adam(lr, betas=(0.9, 0.99))) and
the 50-step respaced sampler of sample.py:52-61 (p_sample_loop over model.forward, batch-sharded, no CFG: bench.py's workload),
both at batch 256 of 32x32x4 latents on DiT-B/2 (MP-AdaLN: the only modulation the reference has code for).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _fix_host_compiler():
    """torch.compile builds small host-side C++ kernels with $CXX; this image's /opt/gcc/bin/g++ wrapper cannot find libgomp.spec
    (`-fopenmp` fails), the system compiler can"""
    if os.environ.get("CXX", "").startswith("/opt/gcc") and os.path.exists("/usr/bin/g++"):
        os.environ["CXX"] = "/usr/bin/g++"
        os.environ["CC"] = "/usr/bin/gcc"


def _import_reference():
    if not os.path.isfile(os.path.join(REF, "src", "models.py")):
        raise RuntimeError("oracle/_ref is missing: run `python oracle/vendor_reference.py` in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from diffusion import create_diffusion  # noqa: E402  (the reference's package)
    from src.models import DIT_MODELS  # noqa: E402
    return DIT_MODELS, create_diffusion


def _events_ms(fn, iters):
    import torch
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def measure(model_name="DiT-B/2", batch=256, arms=("eager_tf32", "eager_bf16_autocast"), train_steps=5, sampling_steps=50, input_size=32,
            clip_denoised=True, log=lambda s: None):
    """-> {arm: {"train_img_s", "train_ms_per_step", "sample_img_s", "sample_ms_per_loop"}}; failures are recorded per arm"""
    import torch
    DIT_MODELS, create_diffusion = _import_reference()
    torch.backends.cuda.matmul.allow_tf32 = True  # train.py:222-223
    torch.backends.cudnn.allow_tf32 = True
    dev = torch.device("cuda")
    out = {}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 4, input_size, input_size, generator=g).to(dev)
    y = torch.randint(0, 1000, (batch,), generator=g).to(dev)
    for arm in arms:
        rec = {}
        try:
            torch.manual_seed(0)
            model = DIT_MODELS[model_name](in_channels=4, input_size=input_size, num_classes=1000).to(dev)
            with torch.no_grad():
                for p in model.parameters():
                    if p.dim() == 0:
                        p.fill_(0.3)
            net = torch.compile(model) if arm.startswith("compile") else model  # train.py:46
            bf16 = arm.endswith("bf16_autocast")
            ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if bf16 else (lambda: torch.autocast("cuda", enabled=False))
            # ---- training step (train.py:86-96)
            diffusion = create_diffusion(timestep_respacing="")
            opt = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0.9, 0.99))
            net.train()

            def train_step():
                t = torch.randint(0, diffusion.num_timesteps, (batch,), device=dev)
                with ctx():
                    loss = diffusion.training_losses(net, x, t, dict(y=y))["loss"].mean()
                opt.zero_grad()
                loss.backward()
                opt.step()
                return loss
            t0 = time.time()
            for _ in range(3):
                last = train_step()
            torch.cuda.synchronize()
            rec["train_warmup_s"] = round(time.time() - t0, 1)
            ms = _events_ms(train_step, train_steps)
            rec.update(train_ms_per_step=round(ms, 2), train_img_s=round(batch / ms * 1e3, 1), train_loss=float(last))
            log(f"{arm}: train {ms:.1f} ms/step")
            del opt
            # ---- 50-step sampler (sample.py:52-61 call pattern without CFG, like bench.py's workload)
            net.eval()
            d = create_diffusion(str(sampling_steps))
            d3 = create_diffusion("3")

            def loop(dd):
                with torch.no_grad(), ctx():
                    return dd.p_sample_loop(net.forward, x.shape, x, clip_denoised=clip_denoised, model_kwargs=dict(y=y), device=dev)
            t0 = time.time()
            loop(d3)
            torch.cuda.synchronize()
            rec["sample_warmup_s"] = round(time.time() - t0, 1)
            ms = _events_ms(lambda: loop(d), 1)
            rec.update(sample_ms_per_loop=round(ms, 1), sample_img_s=round(batch / ms * 1e3, 2))
            log(f"{arm}: sample{sampling_steps} {ms:.0f} ms/loop")
            del net, model
        except Exception as e:  # noqa: BLE001  (an arm that cannot run on this box is reported, not fatal)
            rec["error"] = f"{type(e).__name__}: {str(e)[:300]}"
            log(f"{arm}: FAILED {rec['error']}")
        torch.cuda.empty_cache()
        out[arm] = rec
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="DiT-B/2")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--input-size", type=int, default=32)
    ap.add_argument("--compile", action="store_true", help="also time torch.compile(model) (minutes of compilation)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ref_on_b200.json"))
    a = ap.parse_args()
    _fix_host_compiler()
    import torch
    arms = ["eager_tf32", "eager_bf16_autocast"] + (["compile_tf32", "compile_bf16_autocast"] if a.compile else [])
    res = measure(a.model, a.batch, arms, input_size=a.input_size, log=lambda s: print("[ref_on_b200]", s, file=sys.stderr, flush=True))
    line = {"what": "unmodified reference (oracle/_ref) on this GPU through its own PyTorch path", "model": a.model, "modulation": "adaln",
            "batch": a.batch, "input_size": a.input_size, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
            "clip_denoised": True, "arms": res, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(line, f, indent=1)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
