"""Fused training step: the inner loop of the reference's train.py:86-96
(training_losses -> loss.mean() -> backward -> Adam(lr, betas=(0.9, 0.99))) as one fixed kernel schedule.

Parameters, gradients and Adam moments live in flat fp32 buffers (parameters are re-pointed into the flat
buffer, so `model.parameters()` / `state_dict()` keep working): the optimiser is ONE kernel launch over the
span and data-parallel training needs no per-tensor bookkeeping.  Gradient layout follows the order in which
the backward finishes them (last block first), so each block's gradient slice is all-reduced over NCCL
(async, overlapping the rest of the backward) the moment it is final; ranks hold identical replicas.
"""
import torch

from . import ops
from .parallel import GradReducer, bucket_plan


class TrainStep:
    def __init__(self, model, diffusion, lr=1e-2, betas=(0.9, 0.99), eps=1e-8, world_size=1, lr_lambda=None, ema=None,
                 grad_reduce_dtype=None):
        """`lr_lambda(step)` multiplies `lr` like the reference's LambdaLR (train.py:66,104: the k-th optimiser step, 0-based,
        uses lr * lr_lambda(k)); `ema` (mapdit_b200.ema.EMA) is updated after every step like train.py:105.
        `grad_reduce_dtype`: "bf16" (default when world_size > 1; MAPDIT_GRAD_REDUCE overrides) all-reduces a bf16 copy of each
        gradient bucket — half the NVLink bytes and half the SM time NCCL takes from the backward GEMMs — and Adam reads the
        reduced bf16 gradients with fp32 moments / parameters; "fp32" reduces the fp32 span in place."""
        self.model, self.diffusion = model, diffusion
        self.lr, self.betas, self.eps = lr, betas, eps
        self.lr_lambda, self.ema = lr_lambda, ema
        self.world = world_size
        self.step_count = 0
        m = model
        dev = next(m.parameters()).device
        # flat layout: blocks in reverse (the order the backward finishes them; a block's modulation weight included: its
        # gradient is taken inside the block loop), then the embedders / final layer
        groups = []
        for b in reversed(list(m.blocks)):
            groups.append(list(b.parameters()))
        seen = {id(p) for g in groups for p in g}
        groups.append([p for p in m.parameters() if id(p) not in seen])
        pad = lambda n: (n + 63) // 64 * 64  # every parameter starts 256-byte aligned (vectorised kernels, TMA)
        total = sum(pad(p.numel()) for g in groups for p in g)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad_views, self.slices, self.offset_of = {}, [], {}
        off = 0
        for g in groups:
            start = off
            for p in g:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + n].view(p.shape)
                self.grad_views[id(p)] = self.flat_g[off:off + n].view(p.shape)
                self.offset_of[id(p)] = off
                off += pad(n)
            self.slices.append((start, off))
        self._group_of = {}
        for gi, g in enumerate(groups):
            for p in g:
                self._group_of[id(p)] = gi
        model.engine.invalidate()
        import os
        if grad_reduce_dtype is None:
            grad_reduce_dtype = os.environ.get("MAPDIT_GRAD_REDUCE", "bf16")
        assert grad_reduce_dtype in ("bf16", "fp32")
        self.flat_g16 = None
        if world_size > 1 and grad_reduce_dtype == "bf16":
            self.flat_g16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        # reduce buckets: `blocks_per_bucket` consecutive groups (blocks, in backward order) share one all-reduce; the embedders / final
        # layer group joins the last bucket.  0 = one bucket for everything, launched when the backward is done; 1 = one bucket per
        # block.  Default: two buckets (half the blocks each).  With bf16 buckets the whole all-reduce is ~0.8 ms of a 42 ms step, and
        # every NCCL kernel that starts mid-backward takes SMs from the persistent GEMM it lands on (whose CTAs then run as a second
        # wave), so few large buckets beat many small ones: N = 4 on one box 43.5-43.7 ms against 43.9 with a bucket per block
        # and 43.8 with a single bucket at the end (N = 1: 42.6 ms); N = 2 on another box 43.3 against 43.5-44.4 (N = 1: 42.5).
        n_blocks = len(self.slices) - 1
        bpb = int(os.environ.get("MAPDIT_DP_BLOCKS_PER_BUCKET", str(max(1, (n_blocks + 1) // 2))))
        ng = len(self.slices)
        self._bucket_of_group, self._last_group_of_bucket = bucket_plan(ng, bpb)
        nb = len(self._last_group_of_bucket)
        bucket_slices = [(min(self.slices[gi][0] for gi in range(ng) if self._bucket_of_group[gi] == b),
                          max(self.slices[gi][1] for gi in range(ng) if self._bucket_of_group[gi] == b)) for b in range(nb)]
        self.reducer = GradReducer(self.flat_g, bucket_slices, compressed=self.flat_g16,
                                   compress=(lambda src, dst: ops.cast(src, dst)) if self.flat_g16 is not None else None)

    # gradient hook from the backward: every parameter in `pairs` has its final gradient -> reduce finished buckets.
    # The last bucket (embedders / final layer, ~1 % of the span) only completes at the very end.
    def _on_grads(self, pairs):
        for gi in sorted({self._group_of[id(p)] for p, _ in pairs}):
            b = self._bucket_of_group[gi]
            if gi < len(self.slices) - 1 and gi == self._last_group_of_bucket[b]:
                self.reducer.ready(b)

    def compute_grads(self, x, t, y, noise=None, drop_mask=None, loss_divisor=None, reduce=True, _defer_wait=False):
        """q_sample -> forward -> loss -> backward (-> all-reduce) of the local batch into the flat gradient span `flat_g`
        (train.py:86-95); returns the per-sample losses [N].  `loss_divisor` (default: the local batch size, i.e. loss.mean())
        is what the per-sample losses are divided by before the backward — shard gradients taken with the GLOBAL batch size
        add up to the full-batch gradient."""
        m, d = self.model, self.diffusion
        assert m.training, "TrainStep needs model.train() (forced weight normalisation + label dropout)"
        with torch.cuda.device(x.device):
            tr = m.engine.trainer
            tr.grad_buffers = self.grad_views
            tr.grad_hook = self._on_grads if reduce else None
            try:
                self.reducer.start_step()
                x0 = x.contiguous().float()
                tl = t.contiguous().long()
                if noise is None:
                    noise = torch.randn_like(x0)
                noise = noise.contiguous().float()
                N = x0.shape[0]
                tab = d.device_tables(x0.device)
                x_t = torch.empty_like(x0)
                ops.q_sample(x0, noise, tl, tab, x_t)
                t_model = d._map_tensor(tl.device, tl.dtype)[tl] if hasattr(d, "_map_tensor") else tl
                with torch.no_grad():
                    out, saved = tr.forward(x_t, t_model, y, drop_mask)
                    loss = torch.empty(N, device=x0.device)
                    dout = torch.empty_like(out)
                    gs = torch.full((N,), 1.0 / (loss_divisor or N), device=x0.device)
                    ops.loss_fwd_bwd(out, x0, x_t, noise, tl, tab, loss, None, None, dout, gs, gs)
                    tr.backward(saved, dout)
                    if reduce:
                        self.reducer.finish(wait=not _defer_wait)  # step(): apply_grads waits bucket by bucket
                    self._reduced = bool(reduce)
            finally:
                tr.grad_buffers = None
                tr.grad_hook = None
        return loss

    def apply_grads(self):
        """one fused Adam launch over the flat span (train.py:96,104: optimizer.step(); scheduler.step()), then the EMA hook"""
        if self.flat_g16 is not None and not getattr(self, "_reduced", False):
            raise RuntimeError("TrainStep.apply_grads: the gradients of this step were not all-reduced (compute_grads(reduce=False)); "
                               "with world_size > 1 the optimiser reads the reduced bf16 copy")
        with torch.cuda.device(self.flat_p.device), torch.no_grad():
            lr = self.lr * (self.lr_lambda(self.step_count) if self.lr_lambda is not None else 1.0)
            self.step_count += 1
            # one launch per reduce bucket, each behind its own all-reduce: the update of the first bucket (reduced long ago) runs while
            # the last bucket is still on the wire.  One rank / buckets already awaited: the waits are no-ops.
            spans = self.reducer.slices if self.world > 1 else [(0, self.flat_p.numel())]
            for bi, (lo, hi) in enumerate(spans):
                if hi <= lo:
                    continue
                if self.world > 1:
                    self.reducer.wait_bucket(bi)
                if self.flat_g16 is not None:
                    ops.adam_step_g16(self.flat_p[lo:hi], self.flat_g16[lo:hi], self.flat_m[lo:hi], self.flat_v[lo:hi], lr, self.betas[0],
                                      self.betas[1], self.eps, self.step_count, grad_scale=1.0 / self.world)
                else:
                    ops.adam_step(self.flat_p[lo:hi], self.flat_g[lo:hi], self.flat_m[lo:hi], self.flat_v[lo:hi], lr, self.betas[0],
                                  self.betas[1], self.eps, self.step_count, grad_scale=1.0 / self.world)
            self.model.engine.invalidate()  # parameters moved through raw pointers: cached effective weights are stale
            if self.ema is not None:
                self.ema.update(self.step_count, self.model)

    def step(self, x, t, y, noise=None, drop_mask=None):
        """one optimisation step on the local batch; returns the mean loss (device scalar)"""
        loss = self.compute_grads(x, t, y, noise, drop_mask, _defer_wait=True)
        self.apply_grads()
        return loss.mean()
