#!/usr/bin/env python
"""Kernel timeline of one training step from torch.profiler (CUPTI): per stream busy time, how much of the second
(weight-gradient) stream overlaps the main stream, and the largest gaps.  Developer tool; numbers taken under the
profiler are not bench values.

    python tools/train_timeline.py [model] [batch] [modulation] [wgrad_stream 0|1] [input_size]
"""
import collections
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mapdit_b200 as M  # noqa: E402
from mapdit_b200.train import TrainStep  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "DiT-B/2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
modulation = sys.argv[3] if len(sys.argv) > 3 else "rotation_scaling"
ws = int(sys.argv[4]) if len(sys.argv) > 4 else 1
S = int(sys.argv[5]) if len(sys.argv) > 5 else 32
torch.manual_seed(0)
m = M.DIT_MODELS[name](in_channels=4, input_size=S, num_classes=1000, modulation=modulation).cuda().train()
with torch.no_grad():
    for p in m.parameters():
        if p.dim() == 0:
            p.fill_(0.3)
m.engine.trainer.wgrad_stream = bool(ws)
ts = TrainStep(m, M.create_diffusion(""))
x = torch.randn(B, 4, S, S, device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")
n = torch.randn(B, 4, S, S, device="cuda")
for _ in range(3):
    ts.step(x, t, y, n)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(x, t, y, n)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
end = max(e["ts"] + e["dur"] for e in ev)
streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"]["stream"]].append((e["ts"] - t0, e["ts"] - t0 + e["dur"], e["name"].replace("(anonymous namespace)::", "").replace("void ", "")))
print(f"{name} B={B} {modulation} wgrad_stream={ws}: {len(ev)} kernels, span {(end - t0) / 1e3:.2f} ms")
for sid, ks in streams.items():
    busy = sum(b - a for a, b, _ in ks)
    print(f"  stream {sid}: {len(ks)} kernels, busy {busy / 1e3:.2f} ms")
if len(streams) > 1:
    main_id = max(streams, key=lambda s: len(streams[s]))
    main = streams[main_id]
    ov = 0.0
    by = collections.Counter()
    for sid, ks in streams.items():
        if sid == main_id:
            continue
        for a, b, _ in ks:
            for c, d, nm in main:
                if d <= a:
                    continue
                if c >= b:
                    break
                o = min(b, d) - max(a, c)
                ov += o
                by[nm.split("(")[0][-40:]] += o
    print(f"  overlap of the other streams with the main stream: {ov / 1e3:.2f} ms")
    for nm, o in by.most_common(6):
        print(f"    {o / 1e3:6.2f} ms with {nm}")
allk = sorted((a, b, nm) for ks in streams.values() for a, b, nm in ks)
cover, gaps = 0.0, []
cur_a, cur_b = allk[0][0], allk[0][1]
for a, b, nm in allk[1:]:
    if a > cur_b:
        gaps.append((a - cur_b, cur_b, nm))
        cover += cur_b - cur_a
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
cover += cur_b - cur_a
print(f"  GPU busy (union over streams) {cover / 1e3:.2f} ms, idle {((end - t0) - cover) / 1e3:.2f} ms in {len(gaps)} gaps")
for g, at, nm in sorted(gaps, reverse=True)[:5]:
    print(f"    gap {g:7.1f} us at {at / 1e3:7.2f} ms before {nm.split('(')[0][-50:]}")
agg = collections.Counter()
for a, b, nm in allk:
    agg[nm.split("(")[0][-48:]] += b - a
print("  kernel time by name (in-step, warm):")
for nm, d in agg.most_common(16):
    print(f"    {d / 1e3:7.2f} ms  {nm}")
