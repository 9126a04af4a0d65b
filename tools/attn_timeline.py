#!/usr/bin/env python
"""Timeline of CTA 0 of the attn_v2 kernel (clock64 stamps, developer tool): per step, when the MMA warp saw P and finished
issuing, and when the softmax warps of tiles A/B started waiting for S, saw S and published P."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdit_b200 import _lib, ops  # noqa: E402

if len(sys.argv) > 2:
    _lib.set_option("attn_v2", int(sys.argv[2]))
D, T, H, B = 768, 256, 12, 256
M = B * T
qkv = torch.randn(M, 3 * D, device="cuda").bfloat16()
o = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(M, H, device="cuda")
for _ in range(2):
    ops.cos_attn(qkv, o, B, T, H, 64, lse=lse)
dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
L = _lib.lib()
L.mapdit_attn_debug_buffer.argtypes = [C.c_void_p]
L.mapdit_attn_debug_buffer(C.c_void_p(dbg.data_ptr()))
ops.cos_attn(qkv, o, B, T, H, 64, lse=lse)
torch.cuda.synchronize()
L.mapdit_attn_debug_buffer(None)
d = dbg.cpu().view(4, 64, 4)
t0 = int(d[d > 0].min())
rel = lambda v: int(v) - t0 if int(v) else -1
print("step |  MMA A: P seen, issued |  MMA B: P seen, issued |  SM A: wait, S seen, P out |  SM B: wait, S seen, P out")
for g in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    print(f"{g:4d} | {rel(d[0, g, 0]):7d} {rel(d[0, g, 1]):7d} | {rel(d[1, g, 0]):7d} {rel(d[1, g, 1]):7d} | "
          f"{rel(d[2, g, 0]):7d} {rel(d[2, g, 1]):7d} {rel(d[2, g, 2]):7d} | {rel(d[3, g, 0]):7d} {rel(d[3, g, 1]):7d} {rel(d[3, g, 2]):7d}")
