#!/usr/bin/env python
"""How long the HOST needs to issue one training step (all launches are asynchronous) next to the GPU time of the step:
if the host time approaches the GPU time the step is launch bound and belongs in a CUDA graph."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mapdit_b200 as M  # noqa: E402
from mapdit_b200.train import TrainStep  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "DiT-B/2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
m = M.DIT_MODELS[name](in_channels=4, input_size=32, num_classes=1000).cuda().train()
ts = TrainStep(m, M.create_diffusion(""))
x = torch.randn(B, 4, 32, 32, device="cuda")
t = torch.randint(0, 1000, (B,), device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")
n = torch.randn(B, 4, 32, 32, device="cuda")
for _ in range(3):
    ts.step(x, t, y, n)
torch.cuda.synchronize()
host, K = 0.0, 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K):
    t0 = time.perf_counter()
    ts.step(x, t, y, n)
    host += time.perf_counter() - t0
e1.record()
torch.cuda.synchronize()
torch.cuda.synchronize()
t0 = time.perf_counter()
ts.step(x, t, y, n)
first = time.perf_counter() - t0  # queue empty: pure host cost of issuing one step
torch.cuda.synchronize()
print(f"{name} B={B}: host cost of one step on an empty queue {1e3 * first:.2f} ms")
print(f"{name} B={B}: host issue time {1e3 * host / K:.2f} ms/step, GPU time {e0.elapsed_time(e1) / K:.2f} ms/step")
