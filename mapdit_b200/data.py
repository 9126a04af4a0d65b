"""Data / IO either side of the training hot path ("next" row N3) with the reference's file formats:

* ``LatentDataset``  — train.py:144-176 ``CustomDataset`` (``posterior_means.pt``, ``posterior_stds.pt``, ``labels.pt``,
  ``stats.pt``) resident in HBM; ``sample_batch`` is one fused gather + reparameterise + normalise kernel.
* ``create_lr_lambda`` — train.py:179-197 warm-up / inverse-sqrt decay factor.
* ``setup_experiment`` / ``save_config`` — train.py:200-214, :35-40 (``NNN-Model-Name/checkpoints``, ``config.yaml``).
* ``save_checkpoint`` / ``load_checkpoint`` — train.py:124-132: ``{"model": state_dict, "opt": Adam state_dict}`` with the
  ``_orig_mod.`` key prefix the reference's ``torch.compile`` wrapper produces; the optimiser entry is a genuine
  ``torch.optim.Adam`` state dict built from TrainStep's flat moment buffers, so either side can resume the other's run.
"""
import math
import os
from glob import glob

import torch

from . import ops

PREFIX = "_orig_mod."


class LatentDataset:
    def __init__(self, data_path=None, device="cuda", tensors=None):
        if tensors is None:
            ld = lambda n: torch.load(os.path.join(data_path, n), weights_only=True)
            tensors = dict(posterior_means=ld("posterior_means.pt"), posterior_stds=ld("posterior_stds.pt"), labels=ld("labels.pt"),
                           stats=ld("stats.pt"))
        self.posterior_means = tensors["posterior_means"].to(device, torch.float32).contiguous()
        self.posterior_stds = tensors["posterior_stds"].to(device, torch.float32).contiguous()
        self.labels = tensors["labels"].to(device)
        self.stats = tensors["stats"]
        assert self.posterior_means.shape[0] == self.labels.shape[0] == self.posterior_stds.shape[0]
        self._mean = torch.as_tensor(self.stats["mean"], dtype=torch.float32).reshape(-1).to(device)
        self._std = torch.as_tensor(self.stats["std"], dtype=torch.float32).reshape(-1).to(device)

    @property
    def data_size(self):
        return self.posterior_means.shape[2]

    @property
    def channels(self):
        return self.posterior_means.shape[1]

    def __len__(self):
        return self.posterior_means.shape[0]

    def sample_batch(self, idx, eps=None):
        """(x, y) for the items `idx` (int64 device tensor): x = Normalize(mean + eps * std) (train.py:168-176)"""
        idx = idx.to(self.labels.device, torch.int64).contiguous()
        n = idx.shape[0]
        shape = (n,) + tuple(self.posterior_means.shape[1:])
        if eps is None:
            eps = torch.randn(shape, device=idx.device)
        out = torch.empty(shape, device=idx.device)
        ops.latent_sample(self.posterior_means, self.posterior_stds, idx, eps.contiguous().float(), self._mean, self._std, out)
        return out, self.labels[idx]

    def batches(self, batch_size, generator=None):
        """one shuffled epoch, drop_last=True (the DataLoader settings of train.py:31)"""
        perm = torch.randperm(len(self), device=self.labels.device, generator=generator)
        for i in range(0, len(self) - batch_size + 1, batch_size):
            yield self.sample_batch(perm[i:i + batch_size])


def create_lr_lambda(num_lin_warmup, start_decay):
    """train.py:179-197"""
    def lr_lambda(step):
        if step + 1 < num_lin_warmup:
            return (step + 1) / num_lin_warmup
        if step >= start_decay:
            return 1.0 / math.sqrt(max(step / start_decay, 1))
        return 1.0
    return lr_lambda


def setup_experiment(model_name, results_dir):
    """train.py:200-214"""
    os.makedirs(results_dir, exist_ok=True)
    index = len(glob(os.path.join(results_dir, "*")))
    exp = os.path.join(results_dir, f"{index:03d}-{model_name.replace('/', '-')}")
    os.makedirs(os.path.join(exp, "checkpoints"), exist_ok=True)
    return exp


def save_config(exp_dir, args: dict):
    """train.py:35-40 (yaml dump of the argparse namespace the samplers read back, sample.py:20-21)"""
    import yaml
    with open(os.path.join(exp_dir, "config.yaml"), "w") as f:
        yaml.dump(dict(args), f)


def strip_prefix(sd):
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in sd.items()}


def adam_state_dict(train_step):
    """torch.optim.Adam.state_dict() of a TrainStep (parameter order = model.parameters(), like train.py:57)"""
    ts = train_step
    params = list(ts.model.parameters())
    state = {}
    for i, p in enumerate(params):
        lo = ts.offset_of[id(p)]
        n = p.numel()
        state[i] = {"step": torch.tensor(float(ts.step_count)), "exp_avg": ts.flat_m[lo:lo + n].view(p.shape).clone(),
                    "exp_avg_sq": ts.flat_v[lo:lo + n].view(p.shape).clone()}
    # under the reference's LambdaLR (train.py:66,104) "lr" is the already-scheduled value of the NEXT step and the base value
    # lives in "initial_lr"; without a schedule torch writes no "initial_lr"
    cur = ts.lr * (ts.lr_lambda(ts.step_count) if ts.lr_lambda is not None else 1.0)
    group = dict(lr=cur, **({"initial_lr": ts.lr} if ts.lr_lambda is not None else {}), betas=tuple(ts.betas), eps=ts.eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                 capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False, params=list(range(len(params))))
    return {"state": state if ts.step_count > 0 else {}, "param_groups": [group]}


def load_adam_state_dict(train_step, sd):
    ts = train_step
    params = list(ts.model.parameters())
    for i, p in enumerate(params):
        st = sd["state"].get(i)
        if st is None:
            continue
        lo, n = ts.offset_of[id(p)], p.numel()
        ts.flat_m[lo:lo + n].copy_(st["exp_avg"].reshape(-1))
        ts.flat_v[lo:lo + n].copy_(st["exp_avg_sq"].reshape(-1))
        ts.step_count = int(st["step"])
    g = sd["param_groups"][0]
    # base learning rate: LambdaLR keeps it in "initial_lr" ("lr" is already multiplied by the schedule factor)
    ts.lr, ts.betas, ts.eps = g.get("initial_lr", g["lr"]), tuple(g["betas"]), g["eps"]


def save_checkpoint(path, model, train_step=None, compiled_prefix=True):
    """train.py:124-132; `compiled_prefix` reproduces the `_orig_mod.` keys of the torch.compile-wrapped reference model"""
    sd = {((PREFIX + k) if compiled_prefix else k): v.detach().clone() for k, v in model.state_dict().items()}
    ckpt = {"model": sd}
    if train_step is not None:
        ckpt["opt"] = adam_state_dict(train_step)
    torch.save(ckpt, path)


def load_checkpoint(path, model, train_step=None, map_location=None):
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    model.load_state_dict(ckpt["model"])  # DiT.load_state_dict accepts the `_orig_mod.` prefix
    if train_step is not None and "opt" in ckpt:
        load_adam_state_dict(train_step, ckpt["opt"])
    return ckpt
