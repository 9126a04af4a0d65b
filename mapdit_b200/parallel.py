"""Data-parallel plumbing (one process per GPU, torch.distributed over NCCL/NVLink; gloo in CPU tests).

The hot path shards by sample only: training needs one gradient all-reduce per step, sampling
needs no traffic until the final gather.  Nothing here touches kernels; it is device-agnostic so the N>1 logic
is testable with world_size-2 gloo on CPU.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(global_batch: int, rank: int, world_size: int):
    """contiguous [lo, hi) slice of the global batch owned by `rank` (remainder spread over the first ranks)"""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_samples(local: torch.Tensor, group=None) -> torch.Tensor:
    """final all-gather of batch-sharded sampling results, rank-major (the only collective of the sampling path)"""
    _, w = world()
    if w == 1:
        return local
    outs = [torch.empty_like(local) for _ in range(w)]
    dist.all_gather(outs, local.contiguous(), group=group)
    return torch.cat(outs, dim=0)


def bucket_plan(n_groups: int, blocks_per_bucket: int):
    """Which reduce bucket each gradient group goes into.  Groups 0 .. n_groups-2 are the blocks in backward order, the last group
    holds the embedders / final layer.  `blocks_per_bucket` consecutive blocks share a bucket and the last group joins the last
    bucket; 1 keeps the last group in a bucket of its own (one bucket per block); <= 0 puts everything into one bucket.
    Returns (bucket_of_group, last_group_of_bucket): a bucket is complete when its last group has reported."""
    if n_groups <= 1 or blocks_per_bucket <= 0:
        bog = [0] * n_groups
    else:
        bog = [min(gi, n_groups - 2) // blocks_per_bucket for gi in range(n_groups)]
        if blocks_per_bucket == 1:
            bog[n_groups - 1] = bog[n_groups - 2] + 1
    nb = (max(bog) + 1) if bog else 0
    return bog, [max(gi for gi in range(n_groups) if bog[gi] == b) for b in range(nb)]


class GradReducer:
    """Bucketed, overlapped gradient all-reduce over a flat buffer.

    `slices[i] = (lo, hi)` are contiguous regions of `flat`, ordered as the backward finishes them.  `ready(i)`
    launches the async all-reduce (SUM) of bucket i; `finish()` launches whatever is left and waits for all of
    them, after which `flat` holds the sum over ranks (the optimiser applies the 1/world factor).

    With `compressed` (a second flat buffer of the same length, e.g. bf16) and `compress(src_slice, dst_slice)`, each bucket
    is first copied into `compressed` on the current stream and THAT slice is all-reduced: afterwards `compressed` holds the
    rank sum and `flat` keeps the local gradient."""

    def __init__(self, flat: torch.Tensor, slices, group=None, compressed=None, compress=None):
        self.flat, self.slices, self.group = flat, list(slices), group
        self.compressed, self.compress = compressed, compress
        assert compressed is None or (compressed.numel() == flat.numel() and compress is not None)
        self._works, self._done, self._work_of = [], set(), {}

    def start_step(self):
        self._works, self._done, self._work_of = [], set(), {}

    def ready(self, i: int):
        _, w = world()
        if w == 1 or i in self._done:
            return
        lo, hi = self.slices[i]
        if hi > lo:
            buf = self.flat[lo:hi]
            if self.compressed is not None:
                self.compress(buf, self.compressed[lo:hi])
                buf = self.compressed[lo:hi]
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._work_of[i] = self._works[-1]
        self._done.add(i)

    def finish(self, wait: bool = True):
        """launch whatever is left; `wait=False` leaves the waiting to `wait_bucket` (the optimiser then updates bucket i while
        bucket i+1 is still being reduced)"""
        for i in range(len(self.slices)):
            self.ready(i)
        if wait:
            for wk in self._works:
                wk.wait()
            self._works, self._work_of = [], {}

    def wait_bucket(self, i: int):
        """the current stream waits for bucket i's all-reduce (no-op for one rank or an already awaited bucket)"""
        wk = self._work_of.pop(i, None)
        if wk is not None:
            wk.wait()
        if not self._work_of:
            self._works = []
