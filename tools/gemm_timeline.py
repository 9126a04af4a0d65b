#!/usr/bin/env python
"""Per-tile timeline of CTA 0 of the 2-CTA GEMM (clock64 stamps through the mapdit_attn_debug_buffer hook; developer tool):
when the MMA warp saw its accumulator stage free and had issued the tile's last MMA, and when epilogue warps 4 and 11 started
waiting, saw the accumulator complete and finished the tile -- for the out-proj shape with a plain store and with the fused
residual + modulation epilogue."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mapdit_b200 import _lib, ops
D, T, B = 768, 256, 256
M = B * T
mk = lambda *s: (torch.randn(*s, device="cuda") * 0.05).bfloat16()
o, wo, x, h = mk(M, D), mk(D, D), mk(M, D), mk(M, D)
mods = torch.randn(B, 6 * D, device="cuda"); gain = torch.tensor(0.3, device="cuda")
L = _lib.lib(); L.mapdit_attn_debug_buffer.argtypes = [C.c_void_p]
for name, fn in (("store", lambda: ops.gemm_bf16(o, wo, h)),
                 ("resid_mod", lambda: ops.gemm_bf16(o, wo, x, epilogue=_lib.EPI_RESID_MOD, out2=h, resid=x, gate=mods, shift=mods[:, D:], scale=mods[:, 2*D:], gain=gain, ldmod=6*D, tokens=T))):
    for _ in range(3): fn()
    dbg = torch.zeros(1024, dtype=torch.int64, device="cuda")
    L.mapdit_attn_debug_buffer(C.c_void_p(dbg.data_ptr())); fn(); torch.cuda.synchronize(); L.mapdit_attn_debug_buffer(None)
    d = dbg.cpu(); ew = d[:512].view(8, 16, 4); mm = d[512:576].view(16, 4)
    t0 = int(d[d > 0].min()); rel = lambda v: int(v) - t0 if int(v) else -1
    print(f"== {name}: per tile of CTA 0: MMA [acc free seen, last MMA issued] | epilogue warp 4: [start, acc ready, done] | warp 11: [start, ready, done]")
    for t in range(10):
        print(f"  tile {t}: MMA {rel(mm[t,0]):7d} {rel(mm[t,1]):7d} | w4 {rel(ew[0,t,0]):7d} {rel(ew[0,t,1]):7d} {rel(ew[0,t,2]):7d} | w11 {rel(ew[7,t,0]):7d} {rel(ew[7,t,1]):7d} {rel(ew[7,t,2]):7d}")
    if name == "resid_mod":
        fine = d[576:576 + 128].view(4, 4, 8)  # [warp 4 tile 2, warp 4 tile 3, warp 8 tile 2, warp 8 tile 3][chunk][stamp]
        ev = ["wait resid", "resid ok", "bulk wait ok", "tmem ld ok", "x' written", "stores issued", "h bulk wait ok", "h store issued"]
        for wi, wn in enumerate(("w4 tile2", "w4 tile3", "w8 tile2", "w8 tile3")):
            for c in range(4):
                row = [rel(fine[wi, c, e]) for e in range(8)]
                print(f"  {wn} chunk {c}: " + " ".join(f"{ev[e]}={row[e]}" for e in range(8)) + "  | deltas " + " ".join(str(row[e + 1] - row[e]) for e in range(7)))
