#!/usr/bin/env python
"""Isolated CUDA-event timings of the hot-path kernels at the DiT-B/2, batch 256 shapes (M = 65536 tokens).

    python tools/kbench.py [names...] [--iters 10] [--json out.json]

Every case reports ms per launch and the achieved fraction of its roofline (tensor: algorithmic FLOPs vs the measured
bf16 burst peak; hbm: algorithmic bytes vs the measured copy bandwidth, MEASURED_PEAKS.json).  HBM-bound cases rotate over
enough buffer sets to exceed the 126 MB L2 between launches.  Also the command `ncu` is pointed at (one kernel per case).
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdit_b200 import _lib, ops  # noqa: E402

D, T, H, B = 768, 256, 12, 256
M = B * T
dev = "cuda"


def peaks():
    try:
        p = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
        return p["bf16_tflops"], p["hbm_gbs"]
    except Exception:
        return 1590.0, 6650.0


def mk(*s, dtype=torch.bfloat16, scale=0.05):
    return (torch.randn(*s, device=dev) * scale).to(dtype)


def cases():
    c = {}
    mods = torch.randn(B, 6 * D, device=dev)
    dmods = torch.zeros(B, 6 * D, device=dev)
    gain = torch.tensor(0.3, device=dev)

    def attn():
        qkv, o = mk(M, 3 * D, scale=1.0), mk(M, D)
        lse = torch.empty(M, H, device=dev)
        return (lambda: ops.cos_attn(qkv, o, B, T, H, D // H, lse=lse)), ("tensor", 4 * M * T * D)

    def attn_bwd():
        qkv, o, do = mk(M, 3 * D, scale=1.0), mk(M, D), mk(M, D)
        lse = torch.empty(M, H, device=dev)
        ops.cos_attn(qkv, o, B, T, H, D // H, lse=lse)
        dqkv, delta = torch.empty_like(qkv), torch.empty(M, H, device=dev)
        return (lambda: ops.cos_attn_bwd(qkv, o, do, lse, dqkv, delta, B, T, H, D // H)), ("tensor", 10 * M * T * D)

    def wn_fwd(rows, cols):
        def make():
            sets = []
            for _ in range(4):
                w = torch.randn(rows, cols, device=dev)
                sets.append((w, torch.empty(rows, cols, device=dev, dtype=torch.bfloat16), torch.empty(cols, rows, device=dev, dtype=torch.bfloat16)))
            it = [0]

            def fn():
                w, e, et = sets[it[0] % len(sets)]
                it[0] += 1
                ops.weight_norm_fwd(w, force=True, eff_bf16=e, eff_bf16_t=et)
            return fn, ("hbm", rows * cols * 12)
        return make

    def wn_bwd(rows, cols):
        def make():
            sets = [(torch.randn(rows, cols, device=dev), torch.randn(rows, cols, device=dev), torch.empty(rows, cols, device=dev)) for _ in range(4)]
            it = [0]

            def fn():
                v, g, o = sets[it[0] % len(sets)]
                it[0] += 1
                ops.weight_norm_bwd(v, g, o)
            return fn, ("hbm", rows * cols * 12)
        return make

    def modulate_bwd():
        sets = [(mk(M, D), mk(M, D), mk(M, D)) for _ in range(3)]
        dgp = torch.empty(ops.modulate_bwd_partials(B, D), device=dev)
        it = [0]

        def fn():
            dh, x, R = sets[it[0] % 3]
            it[0] += 1
            ops.modulate_bwd(dh, x, R, mods, mods[:, D:], gain, dmods, dmods[:, D:], dgp, 6 * D, B, T, True)
        return fn, ("hbm", M * D * 2 * 4)

    def modulate_resid_bwd():
        sets = [(mk(M, D), mk(M, D), mk(M, D), mk(M, D), mk(M, D)) for _ in range(2)]
        dgp = torch.empty(ops.modulate_bwd_partials(B, D), device=dev)
        it = [0]

        def fn():
            dh, x, R, y, dy = sets[it[0] % 2]
            it[0] += 1
            ops.modulate_resid_bwd(dh, x, R, mods, mods[:, D:], gain, dmods, dmods[:, D:], dgp, y, dy, mods[:, 2 * D:], dmods[:, 2 * D:],
                                   6 * D, B, T, True)
        return fn, ("hbm", M * D * 2 * 6)

    def rotmod_resid_bwd():
        sets = [(mk(M, D), mk(M, D), mk(M, D), mk(M, D), mk(M, D)) for _ in range(2)]
        dgp = torch.empty(ops.rotmod_bwd_partials(B, D), device=dev)
        it = [0]

        def fn():
            dh, x, R, y, dy = sets[it[0] % 2]
            it[0] += 1
            ops.rotmod_resid_bwd(dh, x, R, mods, mods[:, D:], gain, dmods, dmods[:, D:], dgp, y, dy, mods[:, 2 * D:], dmods[:, 2 * D:],
                                 6 * D, B, T, True)
        return fn, ("hbm", M * D * 2 * 6)

    def patch_embed():
        x = torch.randn(B, 4, 32, 32, device=dev)
        wx, pos = torch.randn(D, 17, device=dev) * 0.2, torch.randn(T, D, device=dev)
        outs = [(mk(M, D), mk(M, D)) for _ in range(2)]
        it = [0]

        def fn():
            x0, h = outs[it[0] % 2]
            it[0] += 1
            ops.patch_embed(x, wx, pos, x0, h, mods, mods[:, D:], gain, 6 * D, 2)
        return fn, ("hbm", M * D * 2 * 2)

    def patch_embed_wgrad():
        x = torch.randn(B, 4, 32, 32, device=dev)
        Rs = [mk(M, D) for _ in range(3)]
        dW = torch.empty(D, 17, device=dev)
        it = [0]

        def fn():
            it[0] += 1
            ops.patch_embed_wgrad(Rs[it[0] % 3], x, dW, 2, 1.0)
        return fn, ("hbm", M * D * 2)

    def gemm_f32_cond():
        a, w = torch.randn(B, D, device=dev), torch.randn(D, D, device=dev)
        out = torch.empty(B, D, device=dev)
        return (lambda: ops.gemm_f32(a, w, out=out)), ("tensor", 2 * B * D * D)

    def resid_bwd():
        sets = [(mk(M, D), mk(M, D), mk(M, D)) for _ in range(3)]
        it = [0]

        def fn():
            R, y, dy = sets[it[0] % 3]
            it[0] += 1
            ops.resid_bwd(R, y, dy, mods[:, 2 * D:], dmods[:, 2 * D:], 6 * D, B, T)
        return fn, ("hbm", M * D * 2 * 4)

    def qk_norm_bwd():
        sets = [(mk(M, 3 * D), mk(M, 3 * D), torch.rand(M, 2 * H, device=dev)) for _ in range(2)]
        it = [0]

        def fn():
            dq, q, sc = sets[it[0] % 2]
            it[0] += 1
            ops.qk_norm_bwd(dq, q, sc, D, D // H)
        return fn, ("hbm", M * 2 * D * 2 * 3)

    def adam():
        n = 130_000_000
        p, g, m, v = (torch.randn(n, device=dev) for _ in range(4))
        v.abs_()
        return (lambda: ops.adam_step(p, g, m, v, 1e-2, 0.9, 0.99, 1e-8, 3)), ("hbm", n * 28)

    def gemm(n, k, epi):
        def make():
            a, w, out = mk(M, k), mk(n, k), mk(M, n)
            x, h = mk(M, n), mk(M, n)
            kw = {}
            if epi == _lib.EPI_RESID_MOD:
                kw = dict(out2=h, resid=x, gate=mods, shift=mods[:, D:], scale=mods[:, 2 * D:], gain=gain, ldmod=6 * D, tokens=T)
            elif epi == _lib.EPI_QKNORM:
                kw = dict(tokens=T, head_dim=D // H, qk_cols=2 * D)
            return (lambda: ops.gemm_bf16(a, w, out, epilogue=epi, **kw)), ("tensor", 2 * M * n * k)
        return make

    def wgrad(n, k):
        def make():
            dy, x, out = mk(M, n), mk(M, k), torch.empty(n, k, device=dev)
            return (lambda: ops.gemm_bf16_tn(dy, x, out)), ("tensor", 2 * M * n * k)
        return make

    c["attn_fwd"] = attn
    c["attn_bwd"] = attn_bwd

    # DiT-XL/2 @ 64x64 (BASELINE config 5): 1024 tokens, 16 heads of 72 channels, batch 32
    XT, XH, XHD, XB = 1024, 16, 72, 32
    XM, XD = XB * XT, XH * XHD

    def attn_xl():
        qkv, o = mk(XM, 3 * XD, scale=1.0), mk(XM, XD)
        lse = torch.empty(XM, XH, device=dev)
        return (lambda: ops.cos_attn(qkv, o, XB, XT, XH, XHD, lse=lse)), ("tensor", 4 * XM * XT * XD)

    def attn_bwd_xl():
        qkv, o, do = mk(XM, 3 * XD, scale=1.0), mk(XM, XD), mk(XM, XD)
        lse = torch.empty(XM, XH, device=dev)
        ops.cos_attn(qkv, o, XB, XT, XH, XHD, lse=lse)
        dqkv, delta = torch.empty_like(qkv), torch.empty(XM, XH, device=dev)
        return (lambda: ops.cos_attn_bwd(qkv, o, do, lse, dqkv, delta, XB, XT, XH, XHD)), ("tensor", 14 * XM * XT * XD)
    c["attn_fwd_xl"] = attn_xl
    c["attn_bwd_xl"] = attn_bwd_xl
    c["wn_fwd_3072x768"] = wn_fwd(3072, 768)
    c["wn_fwd_768x3072"] = wn_fwd(768, 3072)
    c["wn_bwd_3072x768"] = wn_bwd(3072, 768)
    c["modulate_bwd"] = modulate_bwd
    c["resid_bwd"] = resid_bwd
    c["modulate_resid_bwd"] = modulate_resid_bwd
    c["rotmod_resid_bwd"] = rotmod_resid_bwd
    c["patch_embed"] = patch_embed
    c["patch_embed_wgrad"] = patch_embed_wgrad
    c["gemm_f32_cond_256x768x768"] = gemm_f32_cond
    c["qk_norm_bwd"] = qk_norm_bwd
    c["adam_130M"] = adam
    c["gemm_qkv"] = gemm(3 * D, D, _lib.EPI_QKNORM)
    c["gemm_out"] = gemm(D, D, _lib.EPI_RESID_MOD)
    c["gemm_fc1"] = gemm(4 * D, D, _lib.EPI_MPSILU)
    c["gemm_fc2"] = gemm(D, 4 * D, _lib.EPI_RESID_MOD)
    c["gemm_dgrad_fc1"] = gemm(D, 4 * D, _lib.EPI_STORE)
    c["gemm_dgrad_out"] = gemm(D, D, _lib.EPI_STORE)
    c["wgrad_fc1"] = wgrad(4 * D, D)
    c["wgrad_out"] = wgrad(D, D)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("names", nargs="*")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--json", default=None)
    ap.add_argument("--opt", action="append", default=[], help="runtime option name=value (mapdit_set_option), e.g. attn_v2=0")
    ap.add_argument("--variant", type=int, default=0, help="MAPDIT_VAR_* word (16 = dot-product attention -> mma.sync kernels)")
    a = ap.parse_args()
    for kv in a.opt:
        k, v = kv.split("=")
        _lib.set_option(k, int(v))
    ops.set_variant(a.variant)
    tf, hbm = peaks()
    res = {}
    all_cases = cases()
    for name in (a.names or list(all_cases)):
        fn, (bound, work) = all_cases[name]()
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
        ev[0].record()
        for i in range(a.iters):
            fn()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters))[a.iters // 2]
        if bound == "tensor":
            ach, frac, unit = work / ms / 1e9, work / ms / 1e9 / tf, "TFLOP/s"
        else:
            ach, frac, unit = work / ms / 1e6, work / ms / 1e6 / hbm, "GB/s"
        res[name] = dict(ms=round(ms, 4), achieved=round(ach, 1), unit=unit, frac=round(frac, 3), bound=bound)
        print(f"{name:18s} {ms:8.4f} ms  {ach:9.1f} {unit:8s} {100 * frac:5.1f}% of measured {bound} peak", flush=True)
        del fn
        torch.cuda.empty_cache()
    if a.json:
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
