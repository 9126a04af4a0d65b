"""Timestep respacing with the reference's entry points (diffusion/respace.py)."""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion, _ModelWrapperBase


def space_timesteps(num_timesteps, section_counts):
    """Set of retained timesteps (diffusion/respace.py:12-62): "N" / "a,b,c" per-section counts with
    fractional striding, or "ddimN" for the fixed integer stride of the DDIM paper."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            wanted = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == wanted:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(tok) for tok in section_counts.split(",")]
    base, remainder = divmod(num_timesteps, len(section_counts))
    kept, offset = [], 0
    for idx, count in enumerate(section_counts):
        length = base + (1 if idx < remainder else 0)
        if length < count:
            raise ValueError(f"cannot divide section of {length} steps into {count}")
        step = 1 if count <= 1 else (length - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(offset + round(pos))
            pos += step
        offset += length
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """Diffusion over a subset of the base timesteps (diffusion/respace.py:65-114): betas are re-derived from
    the retained cumulative alphas and the model is always called with the ORIGINAL timestep value."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        self.timestep_map = [i for i in range(self.original_num_steps) if i in self.use_timesteps]
        ac = base.alphas_cumprod[self.timestep_map]
        prev = np.concatenate([[1.0], ac[:-1]])
        kwargs["betas"] = 1 - ac / prev
        super().__init__(**kwargs)
        self._map_dev = {}

    def _map_tensor(self, device, dtype):
        key = (str(device), dtype)
        m = self._map_dev.get(key)
        if m is None:
            m = th.tensor(self.timestep_map, device=device, dtype=dtype)
            self._map_dev[key] = m
        return m

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.original_num_steps, self)

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def p_sample(self, model, *args, **kwargs):
        return super().p_sample(self._wrap_model(model), *args, **kwargs)

    def ddim_sample(self, model, *args, **kwargs):
        return super().ddim_sample(self._wrap_model(model), *args, **kwargs)

    def training_losses(self, model, *args, **kwargs):
        return super().training_losses(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def condition_score(self, cond_fn, *args, **kwargs):
        return super().condition_score(self._wrap_model(cond_fn), *args, **kwargs)

    def _timestep_for_model(self, i):
        return self.timestep_map[i]


class _WrappedModel(_ModelWrapperBase):
    """Remaps respaced indices to original timesteps before calling the model (diffusion/respace.py:117-129);
    the map lives on the device once instead of being re-uploaded per call."""

    def __init__(self, model, timestep_map, original_num_steps, owner=None):
        self.model = model
        self.timestep_map = timestep_map
        self.original_num_steps = original_num_steps
        self._owner = owner

    def __call__(self, x, ts, **kwargs):
        if self._owner is not None:
            map_tensor = self._owner._map_tensor(ts.device, ts.dtype)
        else:
            map_tensor = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        return self.model(x, map_tensor[ts], **kwargs)
