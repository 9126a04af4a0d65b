"""Power-function EMA tracking and post-hoc EMA reconstruction with the reference's API and file format
(src/ema.py; Karras et al., arXiv 2312.02696) — "next" row N2.

The host math (std <-> gamma, beta schedule, the small linear system of the post-hoc reconstruction) is numpy; the
per-step update of every parameter of every tracked copy is ONE kernel launch per copy over a device-resident chunk
table (csrc/optim.cu: multi_lerp) instead of the reference's per-parameter `lerp_` loop.
"""
import copy
import os
import re

import numpy as np
import torch

from . import ops

_CHUNK = 1 << 16


def std_to_gamma(std):
    """largest real root of g^3 + 7 g^2 + (16 - s^-2) g + (12 - s^-2) = 0 (src/ema.py:10-20)"""
    std = np.asarray(std)
    inv = std.astype(np.float64).flatten() ** -2
    roots = [np.roots([1, 7, 16 - v, 12 - v]).real.max() for v in inv]
    return np.array(roots, dtype=np.float64).reshape(std.shape)


def gamma_to_std(gammas):
    """src/ema.py:23-30"""
    g = np.asarray(gammas).astype(np.float64)
    return np.sqrt((g + 1) / (np.square(g + 2) * (g + 3)))


def calc_beta(std, t):
    """lerp weight of the step-t update, (1 - 1/t)^(gamma+1) (src/ema.py:33-40)"""
    return (1 - 1 / t) ** (std_to_gamma(np.array(std)) + 1)


def p_dot_p(t_a, gamma_a, t_b, gamma_b):
    """inner product of two power-function profiles (src/ema.py:43-53)"""
    ratio = t_a / t_b
    expo = np.where(t_a < t_b, gamma_b, -gamma_a)
    return (gamma_a + 1) * (gamma_b + 1) * ratio ** expo / ((gamma_a + gamma_b + 1) * np.maximum(t_a, t_b))


def solve_weights(t_i, gamma_i, t_r, gamma_r):
    """least-squares weights of the snapshots (t_i, gamma_i) for the target profile (t_r, gamma_r) (src/ema.py:56-65)"""
    col = lambda v: np.float64(v).reshape(-1, 1)
    row = lambda v: np.float64(v).reshape(1, -1)
    A = p_dot_p(col(t_i), col(gamma_i), row(t_i), row(gamma_i))
    B = p_dot_p(col(t_i), col(gamma_i), row(t_r), row(gamma_r))
    return np.linalg.solve(A, B)


def calculate_posthoc_ema(out_std, ema_dir, verbose=True):
    """state_dict of the EMA profile `out_std` reconstructed from the snapshots `{std:.3f}_{t:07d}.pt` in `ema_dir`
    (src/ema.py:68-114; snapshots are fp16 state dicts under the key "state_dict")."""
    found = []
    for name in os.listdir(ema_dir):
        m_std, m_t = re.search(r"[0-9]*\.[0-9]+", name), re.search(r"_(\d+)\.pt$", name)
        if m_std and m_t:
            found.append((float(m_std.group(0)), int(m_t.group(1)), name))
    assert found, "No EMA snapshots found in the results directory"
    stds = np.array([f[0] for f in found])
    ts = np.array([f[1] for f in found])
    t_out = ts.max()
    if out_std in stds:
        pick = int(np.argmax((stds == out_std) & (ts == t_out)))
        return torch.load(os.path.join(ema_dir, found[pick][2]), weights_only=True)["state_dict"]
    w = solve_weights(ts, std_to_gamma(stds), t_out, std_to_gamma(out_std)).flatten()
    first = torch.load(os.path.join(ema_dir, found[0][2]), weights_only=True)["state_dict"]
    acc = {k: torch.zeros_like(v, dtype=torch.float32) for k, v in first.items()}
    for wi, (_, _, name) in zip(w, found):
        sd = torch.load(os.path.join(ema_dir, name), weights_only=True)["state_dict"]
        for k in acc:
            acc[k] += sd[k].float() * wi
    return acc


class EMA:
    """src/ema.py:117-155: one frozen copy of the network per std, updated every step, snapshotted as fp16."""

    @torch.no_grad()
    def __init__(self, net, results_dir, stds=(0.05, 0.1)):
        self.emas = {s: copy.deepcopy(net).eval().requires_grad_(False) for s in stds}
        self.ema_dir = os.path.join(results_dir, "ema")
        os.makedirs(self.ema_dir, exist_ok=True)
        self._tables = {}

    def _table(self, std, model):
        """device chunk table {dst, src, n} pairing every EMA parameter with the live model's parameter"""
        key = (std, id(model))
        hit = self._tables.get(key)
        live = [model.get_parameter(n) for n, _ in self.emas[std].named_parameters()]
        sig = tuple(p.data_ptr() for p in live)
        if hit is not None and hit[2] == sig:
            return hit[0], hit[1]
        rows = []
        for (name, dst), src in zip(self.emas[std].named_parameters(), live):
            assert dst.is_cuda and src.is_cuda and dst.is_contiguous() and src.is_contiguous() and dst.dtype == torch.float32
            for off in range(0, dst.numel(), _CHUNK):
                n = min(_CHUNK, dst.numel() - off)
                rows.append((dst.data_ptr() + 4 * off, src.data_ptr() + 4 * off, n))
        tab = torch.tensor(rows, dtype=torch.int64).to(live[0].device)
        self._tables[key] = (tab, len(rows), sig)
        return tab, len(rows)

    @torch.no_grad()
    def update(self, t, model):
        """ema <- lerp(ema, model, calc_beta(std, t)) for every parameter (src/ema.py:124-140)"""
        for std in self.emas:
            tab, n = self._table(std, model)
            ops.multi_lerp(tab, n, float(calc_beta(std, t)))
            # the copy's parameters moved through raw pointers (no version bump): drop its cached effective weights, or
            # later eval forwards / graphed sampling loops of the copy would keep using the old normalised weights
            eng = getattr(self.emas[std], "_engine", None)
            if eng is not None:
                eng.invalidate()

    @torch.no_grad()
    def save_snapshot(self, t):
        """`{std:.3f}_{t:07d}.pt` = {"std", "t", "state_dict" (fp16, cpu)} (src/ema.py:142-155)"""
        for std, ema in self.emas.items():
            sd = {k: v.detach().cpu().half() if v.is_floating_point() else v.detach().cpu() for k, v in ema.state_dict().items()}
            torch.save({"std": std, "t": t, "state_dict": sd}, os.path.join(self.ema_dir, f"{std:.3f}_{t:07d}.pt"))
