#!/usr/bin/env python
"""Timeline of CTA 0 of the fused attention backward (attn_bwd_fused_tc; clock64 stamps, developer tool): per (kt, qb)
iteration g, when the MMA warp had issued S/dP of g+1, saw P/dS of g, had issued the dV/dK/dQ MMAs; when softmax warp 2
started waiting for S, saw it, finished the exp pass, saw the staging tiles free and published them; epilogue end; and
when the TMA producer saw each input tile of the next items free."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdit_b200 import _lib, ops  # noqa: E402

D, T, H, B = 768, 256, 12, 256
M = B * T
qkv = torch.randn(M, 3 * D, device="cuda")
sc = torch.empty(M, 2 * H, device="cuda")
ops.qk_normalize_save(qkv, sc, D, 64)
qkv = qkv.bfloat16()
o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
do = torch.randn(M, D, device="cuda").bfloat16()
lse = torch.empty(M, H, device="cuda")
ops.cos_attn(qkv, o, B, T, H, 64, lse=lse)
dqkv = torch.empty_like(qkv)
delta = torch.empty(M, H, device="cuda")
for _ in range(2):
    ops.cos_attn_bwd_qknorm(qkv, o, do, lse, sc, dqkv, delta, B, T, H, 64)
dbg = torch.zeros(4096, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.mapdit_attn_debug_buffer.argtypes = [C.c_void_p]
L.mapdit_attn_debug_buffer(C.c_void_p(dbg.data_ptr()))
ops.cos_attn_bwd_qknorm(qkv, o, do, lse, sc, dqkv, delta, B, T, H, 64)
torch.cuda.synchronize()
L.mapdit_attn_debug_buffer(None)
d = dbg.cpu()
roles = d[:768].view(3, 64, 4)
t0 = int(roles[roles > 0].min())
rel = lambda v: int(v) - t0 if int(v) else -1
print("  g | MMA: polled, P seen, phase2 issued | softmax w2: wait S, S seen, exp done, published | read-out / stash end | i==0 top: before stash, after stash")
for g in range(24):
    m, s, e = roles[0, g], roles[1, g], roles[2, g]
    print(f"{g:3d} | {rel(m[0]):7d} {rel(m[1]):7d} {rel(m[2]):7d} | {rel(s[0]):7d} {rel(s[1]):7d} {rel(s[2]):7d} {rel(s[3]):7d} | {rel(e[0]):7d} | top {rel(e[1]):7d} {rel(e[2]):7d}")
prod = d[768:768 + 64].view(8, 8)
er = roles[2]
print("read-out warp 10 (fused2 only): per item kv0 done, dQ(second) done, item done:")
for it in range(6):
    print(f"  item {it}: " + " ".join(f"{rel(v):7d}" for v in er[4 * it][:3]))
print("TMA producer: clock at which tile k of item it was seen free (order K0 V0 Q0 dO0 Q1 dO1 K1 V1)")
for it in range(6):
    print(f"  item {it}: " + " ".join(f"{rel(v):7d}" for v in prod[it]))

