#!/usr/bin/env python
"""Out-proj GEMM (K = N = 768) against problem size, with a plain store / residual / residual + modulate epilogue: shows the
~22 us single-wave floor and the per-wave cost of each epilogue (developer tool; DESIGN.md §7 'open')."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdit_b200 import _lib, ops
D, T = 768, 256
for B in (16, 32, 64, 128, 256):
    M = B * T
    mk = lambda *s: (torch.randn(*s, device="cuda") * 0.05).bfloat16()
    o, wo, x, h = mk(M, D), mk(D, D), mk(M, D), mk(M, D)
    mods = torch.randn(B, 6 * D, device="cuda"); gain = torch.tensor(0.3, device="cuda")
    def f_mod(): ops.gemm_bf16(o, wo, x, epilogue=_lib.EPI_RESID_MOD, out2=h, resid=x, gate=mods, shift=mods[:, D:], scale=mods[:, 2*D:], gain=gain, ldmod=6*D, tokens=T)
    def f_res(): ops.gemm_bf16(o, wo, x, epilogue=_lib.EPI_RESID, resid=x, gate=mods, ldmod=6*D, tokens=T)
    def f_st(): ops.gemm_bf16(o, wo, h)
    res = []
    for fn in (f_st, f_res, f_mod):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res.append(f"{ms*1e3:7.1f} us {2*M*D*D/ms/1e9:7.0f} TF")
    print(f"M={M:6d}: store {res[0]} | resid {res[1]} | resid_mod {res[2]}")
