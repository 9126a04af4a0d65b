"""N>1 host logic on CPU: world_size-2 gloo processes exercising batch sharding, the bucketed gradient
reducer used by TrainStep, and the final sample gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mapdit_b200.parallel import GradReducer, gather_samples, shard_range
    ok = True
    # batch sharding covers the batch exactly once
    lo, hi = shard_range(11, rank, world)
    cover = torch.zeros(11)
    cover[lo:hi] = 1
    dist.all_reduce(cover)
    ok &= bool((cover == 1).all())
    # bucketed reducer: buckets reduced out of order / lazily still give the rank sum
    flat = torch.arange(100, dtype=torch.float32) * (rank + 1)
    red = GradReducer(flat, [(60, 100), (20, 60), (0, 20), (100, 100)])
    red.start_step()
    red.ready(1)
    red.ready(1)  # idempotent
    red.finish()
    ok &= bool(torch.equal(flat, torch.arange(100, dtype=torch.float32) * sum(r + 1 for r in range(world))))
    # N-rank average == 1-rank gradient on the concatenated batch (mean of per-rank means with equal shards)
    g_local = torch.full((4,), float(rank))
    red2 = GradReducer(g_local, [(0, 4)])
    red2.start_step()
    red2.finish()
    ok &= bool(torch.allclose(g_local / world, torch.full((4,), (world - 1) / 2)))
    # compressed reduction (TrainStep's bf16 gradient all-reduce): the rank sum lands in the bf16 copy, the fp32 span keeps
    # the local gradient; values chosen exactly representable in bf16
    loc = torch.arange(64, dtype=torch.float32) * (rank + 1)
    c16 = torch.zeros(64, dtype=torch.bfloat16)
    red3 = GradReducer(loc, [(32, 64), (0, 32)], compressed=c16, compress=lambda a, b: b.copy_(a))
    red3.start_step()
    red3.ready(0)
    red3.finish()
    ok &= bool(torch.equal(c16.float(), torch.arange(64, dtype=torch.float32) * sum(r + 1 for r in range(world))))
    ok &= bool(torch.equal(loc, torch.arange(64, dtype=torch.float32) * (rank + 1)))
    # deferred waits (TrainStep.step: the optimiser updates bucket i behind its own all-reduce while bucket i+1 is on the wire)
    f4 = torch.arange(40, dtype=torch.float32) * (rank + 1)
    red4 = GradReducer(f4, [(0, 16), (16, 40)])
    red4.start_step()
    red4.ready(0)
    red4.finish(wait=False)
    red4.wait_bucket(0)
    ok &= bool(torch.equal(f4[:16], torch.arange(16, dtype=torch.float32) * sum(r + 1 for r in range(world))))
    red4.wait_bucket(1)
    red4.wait_bucket(1)  # idempotent
    ok &= bool(torch.equal(f4, torch.arange(40, dtype=torch.float32) * sum(r + 1 for r in range(world))))
    # final gather is rank-major
    s = gather_samples(torch.full((2, 3), float(rank)))
    ok &= s.shape == (2 * world, 3) and bool((s[:2] == 0).all()) and bool((s[2:4] == 1).all())
    out[rank] = ok
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
