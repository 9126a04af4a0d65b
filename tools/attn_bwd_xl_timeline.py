#!/usr/bin/env python
"""Timeline of CTA 0 of the persistent attn_bwd_dq_tc<72> / attn_bwd_dkv_tc<72> (DiT-XL/2 @ 64x64 shapes; clock64 stamps, developer tool): per key block j, when the MMA
warp had issued S/dP of j+1, saw dS of j and had issued the dQ MMAs; when softmax warp 2 started waiting for S, saw it, finished
the exp pass and published dS; when the TMA producer saw the ring slot of block j free."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdit_b200 import _lib, ops  # noqa: E402

B, T, H, HD = 32, 1024, 16, 72
D, M = H * HD, B * T
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
do = torch.randn(M, D, device="cuda").bfloat16()
lse = torch.empty(M, H, device="cuda")
ops.cos_attn(qkv, o, B, T, H, HD, lse=lse)
dqkv, delta = torch.empty_like(qkv), torch.empty(M, H, device="cuda")
for _ in range(2):
    ops.cos_attn_bwd(qkv, o, do, lse, dqkv, delta, B, T, H, HD)
dbg = torch.zeros(2048, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.mapdit_attn_debug_buffer.argtypes = [C.c_void_p]
L.mapdit_attn_debug_buffer(C.c_void_p(dbg.data_ptr()))
ops.cos_attn_bwd(qkv, o, do, lse, dqkv, delta, B, T, H, HD)
torch.cuda.synchronize()
L.mapdit_attn_debug_buffer(None)
def show(d, name):
    t0 = int(d[d > 0].min())
    rel = lambda v: int(v) - t0 if int(v) else -1
    print(f"---- {name}: CTA 0, blocks counted across its items ({T // 64} per item)")
    print("  g | MMA: loop top, ring slot g+2 seen, tile(g) seen, MMAs issued | softmax w2: wait S, S seen, exp done, published | TMA: slot free")
    for g in range(2 * T // 64 + 4):
        m, s, t = d[0, g], d[1, g], d[2, g]
        print(f"{g:3d} | {rel(m[3]):7d} {rel(m[0]):7d} {rel(m[1]):7d} {rel(m[2]):7d} | {rel(s[0]):7d} {rel(s[1]):7d} {rel(s[2]):7d} {rel(s[3]):7d} | {rel(t[0]):7d}")
    print("item | softmax w2: next item's rows in TMEM, accumulators in registers, stored")
    for it in range(4):
        c = d[3, it]
        print(f"{it:4d} | {rel(c[0]):7d} {rel(c[1]):7d} {rel(c[2]):7d}")


allr = dbg.cpu().view(8, 64, 4)
show(allr[:4], "attn_bwd_dq_tc<72>")
show(allr[4:], "attn_bwd_dkv_tc<72>")
