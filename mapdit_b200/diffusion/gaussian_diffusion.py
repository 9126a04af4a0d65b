"""Gaussian diffusion process with the reference's entry points (diffusion/gaussian_diffusion.py).

Host side (this file): float64 coefficient tables, argument handling, the step loop and CUDA-graph
management.  Device side: the fused kernels K5 (step update) and K6 (q_sample + loss and its
gradient) in csrc/diffusion.cu.  Only the configuration the reference ever instantiates has
kernels — epsilon prediction, learned-range variance, MSE(+VB) loss (diffusion/__init__.py:10-46 as
called from train.py:43 and sample.py:52); other enum values raise NotImplementedError.
"""
import enum

import numpy as np
import torch as th

from .. import ops


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """linear schedule of Ho et al. scaled to the step count (gaussian_diffusion.py:98-115)."""
    if schedule_name != "linear":
        raise NotImplementedError(f"beta schedule {schedule_name!r}: only 'linear' is used by the reference scripts")
    k = 1000 / num_diffusion_timesteps
    return np.linspace(k * 1e-4, k * 2e-2, num_diffusion_timesteps, dtype=np.float64)


def _randn_like(x):
    """noise source of q_sample / p_sample (gaussian_diffusion.py:225,410); tests patch this."""
    return th.randn_like(x)


def _is_dit_callable(model):
    """-> (dit_module, uses_cfg) when `model` is our DiT, DiT.forward or DiT.forward_with_cfg, else (None, False)."""
    from ..dit import DiT
    if isinstance(model, DiT):
        return model, False
    owner = getattr(model, "__self__", None)
    fn = getattr(model, "__func__", None)
    if isinstance(owner, DiT):
        if fn is DiT.forward_with_cfg:
            return owner, True
        if fn is DiT.forward:
            return owner, False
    return None, False


class _FusedLoss(th.autograd.Function):
    """K6: loss/mse/vb [N] from the model output; backward re-runs the kernel to emit d/d model_output."""

    @staticmethod
    def forward(ctx, model_output, x_start, x_t, noise, t, tables):
        n = x_start.shape[0]
        dev = x_start.device
        loss, mse, vb = (th.empty(n, device=dev, dtype=th.float32) for _ in range(3))
        mo = model_output.detach().contiguous().float()
        ops.loss_fwd_bwd(mo, x_start, x_t, noise, t, tables, loss, mse, vb, None, None, None)
        ctx.save_for_backward(mo, x_start, x_t, noise, t, tables)
        return loss, mse, vb

    @staticmethod
    def backward(ctx, g_loss, g_mse, g_vb):
        mo, x_start, x_t, noise, t, tables = ctx.saved_tensors
        z = th.zeros_like(g_loss) if g_loss is not None else None

        def pick(g):
            return g if g is not None else (z if z is not None else 0.0)
        gs_eps = (pick(g_loss) + pick(g_mse)).contiguous().float()
        gs_var = (pick(g_loss) + pick(g_vb)).contiguous().float()
        grad = th.empty_like(mo)
        ops.loss_fwd_bwd(mo, x_start, x_t, noise, t, tables, None, None, None, grad, gs_eps, gs_var)
        return grad, None, None, None, None, None


class GaussianDiffusion:
    """Tables + entry points of gaussian_diffusion.py:144-201 (same attribute names)."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        betas = np.array(betas, dtype=np.float64)
        assert betas.ndim == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        pv = self.posterior_variance
        self.posterior_log_variance_clipped = np.log(np.append(pv[1], pv[1:])) if len(pv) > 1 else np.array([])
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - ac)
        self._dev_tables = {}
        self._graphs = {}

    # ------------------------------------------------------------------ helpers
    def _kernels_available(self):
        return (self.model_mean_type == ModelMeanType.EPSILON and self.model_var_type == ModelVarType.LEARNED_RANGE
                and self.loss_type == LossType.MSE)

    def _require_kernels(self, what):
        if not self._kernels_available():
            raise NotImplementedError(
                f"{what}: only EPSILON + LEARNED_RANGE + MSE (the configuration create_diffusion() is called with in "
                "train.py:43 / sample.py:52) has CUDA kernels")

    def device_tables(self, device):
        """fp32 [8, steps] coefficient table resident on the device (layout: include/mapdit.h)."""
        key = str(device)
        tab = self._dev_tables.get(key)
        if tab is None:
            rows = [self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod, self.sqrt_recip_alphas_cumprod,
                    self.sqrt_recipm1_alphas_cumprod, self.posterior_mean_coef1, self.posterior_mean_coef2,
                    self.posterior_log_variance_clipped, np.log(self.betas), self.alphas_cumprod, self.alphas_cumprod_prev]
            tab = th.from_numpy(np.stack(rows)).to(device=device).float().contiguous()
            self._dev_tables[key] = tab
        return tab

    def _scale_timesteps(self, t):
        return t

    def _call_model(self, model, x, t, model_kwargs):
        out = model(x, self._scale_timesteps(t), **(model_kwargs or {}))
        if isinstance(out, tuple):
            return out
        return out, None

    # ------------------------------------------------------------------ q(x_t | x_0)
    def q_sample(self, x_start, t, noise=None):
        """gaussian_diffusion.py:215-230"""
        if noise is None:
            noise = _randn_like(x_start)
        assert noise.shape == x_start.shape
        x0 = x_start.contiguous().float()
        out = th.empty_like(x0)
        ops.q_sample(x0, noise.contiguous().float(), t.contiguous().long(), self.device_tables(x0.device), out)
        return out

    # ------------------------------------------------------------------ p(x_{t-1} | x_t)
    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None):
        """gaussian_diffusion.py:254-332 -> dict(mean, variance, log_variance, pred_xstart, extra)"""
        self._require_kernels("p_mean_variance")
        B, C = x.shape[:2]
        assert t.shape == (B,)
        model_output, extra = self._call_model(model, x, t, model_kwargs)
        assert model_output.shape == (B, C * 2, *x.shape[2:])
        return self._p_mean_variance_from_output(model_output, x, t, clip_denoised, denoised_fn, extra)

    def _p_mean_variance_from_output(self, model_output, x, t, clip_denoised, denoised_fn, extra=None):
        xc = x.contiguous().float()
        mo = model_output.contiguous().float()
        tl = t.contiguous().long()
        mean, var, logvar, x0 = (th.empty_like(xc) for _ in range(4))
        tab = self.device_tables(xc.device)
        if denoised_fn is None:
            ops.p_mean_variance(mo, xc, tl, tab, mean, var, logvar, x0, clip_denoised)
        else:
            # user callback between x0 prediction and the posterior mean (gaussian_diffusion.py:310-323):
            # kernel for the unclipped x0 / variance, callback in python, kernel again for the mean.
            ops.p_mean_variance(mo, xc, tl, tab, mean, var, logvar, x0, False)
            x0 = denoised_fn(x0)
            if clip_denoised:
                x0 = x0.clamp(-1, 1)
            x0 = x0.contiguous().float()
            ops.posterior_mean(x0, xc, tl, tab, mean)
        return {"mean": mean, "variance": var, "log_variance": logvar, "pred_xstart": x0, "extra": extra}

    def condition_mean(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        """gaussian_diffusion.py:346-356 (user-supplied guidance gradient; plain tensor arithmetic)"""
        gradient = cond_fn(x, t, **(model_kwargs or {}))
        return p_mean_var["mean"].float() + p_mean_var["variance"] * gradient.float()

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None):
        """gaussian_diffusion.py:376-417 -> dict(sample, pred_xstart)"""
        self._require_kernels("p_sample")
        B, C = x.shape[:2]
        assert t.shape == (B,)
        if denoised_fn is None and cond_fn is None:
            model_output, _ = self._call_model(model, x, t, model_kwargs)
            assert model_output.shape == (B, C * 2, *x.shape[2:])
            noise = _randn_like(x)
            xc = x.contiguous().float()
            sample, x0 = th.empty_like(xc), th.empty_like(xc)
            ops.diffusion_step(model_output.contiguous().float(), xc, noise.contiguous().float(), t.contiguous().long(),
                               self.device_tables(xc.device), sample, x0, clip_denoised)
            return {"sample": sample, "pred_xstart": x0}
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, model_kwargs=model_kwargs)
        noise = _randn_like(x)
        if cond_fn is not None:
            out["mean"] = self.condition_mean(cond_fn, out, x, t, model_kwargs=model_kwargs)
        sample = th.empty_like(out["mean"])
        ops.noise_add(out["mean"].contiguous(), out["log_variance"], noise.contiguous().float(), t.contiguous().long(), sample)
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                      device=None, progress=False):
        """gaussian_diffusion.py:419-462"""
        final = None
        for sample in self.p_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised,
                                                     denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                                                     device=device, progress=progress, _want_xstart=False):
            final = sample
        return final["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False, _want_xstart=True):
        """gaussian_diffusion.py:464-511.  When `model` is a mapdit_b200 DiT (or its forward /
        forward_with_cfg) and no python callbacks are given, each step is one CUDA-graph replay of
        [timestep remap -> DiT forward -> (CFG combine) -> fused step update]."""
        self._require_kernels("p_sample_loop")
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        dit, uses_cfg = _is_dit_callable(getattr(model, "model", model) if isinstance(model, _ModelWrapperBase) else model)
        if dit is not None and denoised_fn is None and cond_fn is None and th.device(device).type == "cuda":
            yield from self._graphed_loop(dit, uses_cfg, img, indices, clip_denoised, model_kwargs or {}, _want_xstart)
            return
        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            with th.no_grad():
                out = self.p_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                    model_kwargs=model_kwargs)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ CUDA-graph fast path
    def _timestep_for_model(self, i):
        """index of respaced step i in the model's own time axis (identity here; SpacedDiffusion remaps)"""
        return i

    def _graphed_loop(self, dit, uses_cfg, img, indices, clip_denoised, model_kwargs, want_xstart, ddim_eta=None):
        y = model_kwargs["y"]
        cfg_scale = float(model_kwargs["cfg_scale"]) if uses_cfg else None
        N = img.shape[0]
        dev = img.device
        # the captured graph bakes in raw pointers of parameters and buffers (gains, embedding table, pos_embed, the engine's
        # weight buffers): the key carries them, so re-pointed parameters (TrainStep's flat span, .to(), a recycled id()) miss
        ptrs = tuple(p.data_ptr() for p in dit.parameters()) + tuple(b.data_ptr() for b in dit.buffers())
        key = (id(dit), ptrs, N, tuple(img.shape), str(dev), uses_cfg, cfg_scale, bool(clip_denoised), dit.compute_dtype, dit.training,
               ddim_eta)
        st = self._graphs.get(key)
        if st is None:  # drop stale graphs of the same model object (their pointers are dead)
            for k in [k for k in self._graphs if k[0] == id(dit) and k[1] != ptrs]:
                del self._graphs[k]
        tab = self.device_tables(dev)
        # Pipelined conditioning (eval): everything at the head of the forward that depends on (t, y) only — Fourier features, the
        # embedder MLPs, the modulation GEMM of all blocks, the rotation tables, MPScale: ~26 small dependent launches, ~0.25 ms of
        # a 12 ms DiT-B/2 step — is computed for step k+1 on a side branch of step k's graph, into the conditioning slot the blocks
        # of step k do not read.  Two graphs (even / odd steps) alternate the slots.
        pipelined = not dit.training
        eng = dit.engine
        with th.no_grad():
            if st is None:
                st = {"img": th.empty(img.shape, device=dev, dtype=th.float32), "noise": th.empty(img.shape, device=dev, dtype=th.float32),
                      "t": th.zeros(N, device=dev, dtype=th.int64), "tm": th.zeros(N, device=dev, dtype=th.int64),
                      "tm_next": th.zeros(N, device=dev, dtype=th.int64),
                      "y": th.zeros(N, device=dev, dtype=th.int64), "x0": th.empty(img.shape, device=dev, dtype=th.float32)}
                branch = th.cuda.Stream(device=dev)

                def body(p=None):
                    cur = th.cuda.current_stream(dev)
                    if p is not None:  # next step's conditioning beside this step's blocks
                        branch.wait_stream(cur)
                        with th.cuda.stream(branch):
                            eng.conditioning(st["tm_next"], st["y"], slot=1 - p)
                    if uses_cfg:
                        if p is None:
                            mo = dit.forward_with_cfg(st["img"], st["tm"], st["y"], cfg_scale)
                        else:  # forward_with_cfg (src/dit.py:107-118) with the conditioning slot handed over
                            half = st["img"][: N // 2]
                            mo = eng.forward(th.cat([half, half], dim=0), st["tm"], st["y"], train=False, cond_ready=p)
                            ops.cfg_combine(mo, dit.in_channels, cfg_scale)
                    elif p is None:
                        mo = dit.forward(st["img"], st["tm"], st["y"])
                    else:
                        mo = eng.forward(st["img"], st["tm"], st["y"], train=False, cond_ready=p)
                    if ddim_eta is None:
                        ops.diffusion_step(mo, st["img"], st["noise"], st["t"], tab, st["img"], st["x0"], clip_denoised)
                    else:
                        ops.ddim_step(mo, st["img"], st["noise"], st["t"], tab, st["img"], st["x0"], clip_denoised, ddim_eta)
                    if p is not None:
                        cur.wait_stream(branch)

                st["y"].copy_(y)
                st["img"].copy_(img)
                st["noise"].zero_()
                variants = (0, 1) if pipelined else (None,)
                side = th.cuda.Stream(device=dev)
                side.wait_stream(th.cuda.current_stream(dev))
                with th.cuda.stream(side):  # warm-up: builds weight caches / workspaces / the second conditioning slot outside capture
                    if pipelined:
                        eng.conditioning(st["tm_next"], st["y"], slot=0)
                    for p in variants:
                        body(p)
                th.cuda.current_stream(dev).wait_stream(side)
                from .. import _lib
                st["graphs"], st["kernels"] = [], []
                for p in variants:
                    g = th.cuda.CUDAGraph()
                    n_before = _lib.launch_count()
                    with th.cuda.graph(g):
                        body(p)
                    st["graphs"].append(g)
                    st["kernels"].append(_lib.launch_count() - n_before)
                self._graphs[key] = st
            from .. import _lib
            eng.weights(dit.compute_dtype, train=False)  # refresh cached effective weights if parameters changed
            st["y"].copy_(y)
            st["img"].copy_(img)
            first = True
            for k, i in enumerate(indices):
                st["t"].fill_(i)
                st["tm"].fill_(self._timestep_for_model(i))
                if pipelined:
                    if first:  # the first step's conditioning has no previous graph to ride in
                        st["tm_next"].fill_(self._timestep_for_model(i))
                        eng.conditioning(st["tm_next"], st["y"], slot=0)
                        first = False
                    st["tm_next"].fill_(self._timestep_for_model(max(i - 1, 0)))
                st["noise"].copy_(_randn_like(st["img"]))
                which = (k & 1) if pipelined else 0
                st["graphs"][which].replay()
                _lib.note_graph_replay(st["kernels"][which])
                if want_xstart:
                    yield {"sample": st["img"].clone(), "pred_xstart": st["x0"].clone()}
            if not want_xstart:
                yield {"sample": st["img"].clone(), "pred_xstart": st["x0"].clone()}

    # ------------------------------------------------------------------ training
    def training_losses(self, model, x_start, t, model_kwargs=None, noise=None):
        """gaussian_diffusion.py:715-787 (MSE + learned-range VB) -> dict(loss, mse, vb), each [N]"""
        self._require_kernels("training_losses")
        if model_kwargs is None:
            model_kwargs = {}
        if noise is None:
            noise = _randn_like(x_start)
        x0 = x_start.contiguous().float()
        noise = noise.contiguous().float()
        tl = t.contiguous().long()
        tab = self.device_tables(x0.device)
        x_t = th.empty_like(x0)
        ops.q_sample(x0, noise, tl, tab, x_t)
        model_output = model(x_t, self._scale_timesteps(tl), **model_kwargs)
        B, C = x_t.shape[:2]
        assert model_output.shape == (B, C * 2, *x_t.shape[2:])
        loss, mse, vb = _FusedLoss.apply(model_output, x0, x_t, noise, tl, tab)
        return {"loss": loss, "mse": mse, "vb": vb}

    # ------------------------------------------------------------------ not on the hot path
    def _off_path(self, name):
        raise NotImplementedError(f"{name} is outside the accelerated hot path (SURVEY.md §2 row 8: never called by the "
                                  "reference scripts); DDIM is a 'next' row (N4)")

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0):
        """gaussian_diffusion.py:513-560 -> dict(sample, pred_xstart); one fused kernel per step ("next" row N4)"""
        self._require_kernels("ddim_sample")
        if denoised_fn is not None or cond_fn is not None:
            self._off_path("ddim_sample with denoised_fn / cond_fn")
        B, C = x.shape[:2]
        assert t.shape == (B,)
        model_output, _ = self._call_model(model, x, t, model_kwargs)
        assert model_output.shape == (B, C * 2, *x.shape[2:])
        noise = _randn_like(x)
        xc = x.contiguous().float()
        sample, x0 = th.empty_like(xc), th.empty_like(xc)
        ops.ddim_step(model_output.contiguous().float(), xc, noise.contiguous().float(), t.contiguous().long(),
                      self.device_tables(xc.device), sample, x0, clip_denoised, eta)
        return {"sample": sample, "pred_xstart": x0}

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                         device=None, progress=False, eta=0.0):
        """gaussian_diffusion.py:600-631"""
        final = None
        for sample in self.ddim_sample_loop_progressive(model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                                        cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress,
                                                        eta=eta, _want_xstart=False):
            final = sample
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, _want_xstart=True):
        """gaussian_diffusion.py:633-680; same CUDA-graph step loop as p_sample_loop_progressive with the DDIM update kernel"""
        self._require_kernels("ddim_sample_loop")
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise if noise is not None else th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        dit, uses_cfg = _is_dit_callable(model)
        if dit is not None and denoised_fn is None and cond_fn is None and th.device(device).type == "cuda":
            yield from self._graphed_loop(dit, uses_cfg, img, indices, clip_denoised, model_kwargs or {}, _want_xstart, ddim_eta=float(eta))
            return
        for i in indices:
            t = th.tensor([i] * shape[0], device=device)
            with th.no_grad():
                out = self.ddim_sample(model, img, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                                       model_kwargs=model_kwargs, eta=eta)
                yield out
                img = out["sample"]

    # ------------------------------------------------------------------ evaluation helpers (never called by the reference scripts)
    # Thin host-side passthroughs: the model call and p_mean_variance run on the kernels, the handful of
    # [N, C, H, W] elementwise expressions around them are plain tensor arithmetic — these are not on the hot path.
    def _coef(self, table, t, like):
        """table[t] broadcast against `like` (gaussian_diffusion.py:861-873)"""
        v = th.from_numpy(np.asarray(table)).to(device=t.device)[t.long()].float()
        return v.view(-1, *([1] * (like.dim() - 1)))

    def condition_score(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        """gaussian_diffusion.py:358-374: guidance applied to the predicted epsilon (Song et al.)"""
        abar = self._coef(self.alphas_cumprod, t, x)
        x0 = p_mean_var["pred_xstart"]
        eps = (self._coef(self.sqrt_recip_alphas_cumprod, t, x) * x - x0) / self._coef(self.sqrt_recipm1_alphas_cumprod, t, x)
        eps = eps - (1 - abar).sqrt() * cond_fn(x, t, **(model_kwargs or {}))
        out = dict(p_mean_var)
        out["pred_xstart"] = self._coef(self.sqrt_recip_alphas_cumprod, t, x) * x - self._coef(self.sqrt_recipm1_alphas_cumprod, t, x) * eps
        out["mean"] = (self._coef(self.posterior_mean_coef1, t, x) * out["pred_xstart"] + self._coef(self.posterior_mean_coef2, t, x) * x)
        return out

    def ddim_reverse_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0):
        """gaussian_diffusion.py:562-598: one step of the deterministic DDIM ODE run forwards, x_t -> x_{t+1}"""
        assert eta == 0.0, "Reverse ODE only for deterministic path"
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn, model_kwargs=model_kwargs)
        if cond_fn is not None:
            out = self.condition_score(cond_fn, out, x, t, model_kwargs=model_kwargs)
        x0 = out["pred_xstart"]
        eps = (self._coef(self.sqrt_recip_alphas_cumprod, t, x) * x - x0) / self._coef(self.sqrt_recipm1_alphas_cumprod, t, x)
        abar_next = self._coef(self.alphas_cumprod_next, t, x)
        return {"sample": x0 * abar_next.sqrt() + (1 - abar_next).sqrt() * eps, "pred_xstart": x0}

    def _vb_terms_bpd(self, model, x_start, x_t, t, clip_denoised=True, model_kwargs=None):
        """gaussian_diffusion.py:682-713 -> dict(output [N] in bits, pred_xstart): KL(q(x_{t-1}|x_t,x_0) || p) or, at t == 0, the
        discretised-Gaussian decoder NLL (diffusion_utils.py:10-36,62-88)"""
        out = self.p_mean_variance(model, x_t, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
        q_mean = self._coef(self.posterior_mean_coef1, t, x_t) * x_start + self._coef(self.posterior_mean_coef2, t, x_t) * x_t
        q_logvar = self._coef(self.posterior_log_variance_clipped, t, x_t)
        p_mean, p_logvar = out["mean"], out["log_variance"]
        kl = 0.5 * (-1.0 + p_logvar - q_logvar + th.exp(q_logvar - p_logvar) + (q_mean - p_mean) ** 2 * th.exp(-p_logvar))
        kl = kl.flatten(1).mean(1) / np.log(2.0)
        # decoder NLL of x_start under N(p_mean, exp(p_logvar)) discretised to 256 bins on [-1, 1]
        cdf = lambda v: 0.5 * (1.0 + th.tanh(np.sqrt(2.0 / np.pi) * (v + 0.044715 * v ** 3)))
        centered, inv_std = x_start - p_mean, th.exp(-0.5 * p_logvar)
        cdf_plus, cdf_min = cdf(inv_std * (centered + 1.0 / 255.0)), cdf(inv_std * (centered - 1.0 / 255.0))
        log_probs = th.where(x_start < -0.999, th.log(cdf_plus.clamp(min=1e-12)),
                             th.where(x_start > 0.999, th.log((1.0 - cdf_min).clamp(min=1e-12)),
                                      th.log((cdf_plus - cdf_min).clamp(min=1e-12))))
        nll = -log_probs.flatten(1).mean(1) / np.log(2.0)
        return {"output": th.where(t == 0, nll, kl), "pred_xstart": out["pred_xstart"]}

    def _prior_bpd(self, x_start):
        """gaussian_diffusion.py:789-804: KL(q(x_T | x_0) || N(0, I)) in bits per dimension"""
        t = th.full((x_start.shape[0],), self.num_timesteps - 1, device=x_start.device, dtype=th.long)
        mean = self._coef(self.sqrt_alphas_cumprod, t, x_start) * x_start
        logvar = self._coef(self.log_one_minus_alphas_cumprod, t, x_start)
        kl = 0.5 * (-1.0 - logvar + th.exp(logvar) + mean ** 2)
        return kl.flatten(1).mean(1) / np.log(2.0)

    def calc_bpd_loop(self, model, x_start, clip_denoised=True, model_kwargs=None):
        """gaussian_diffusion.py:806-858 -> dict(total_bpd [N], prior_bpd [N], vb / xstart_mse / mse [N, T])"""
        x_start = x_start.float()
        n = x_start.shape[0]
        vb, xstart_mse, mse = [], [], []
        for i in range(self.num_timesteps - 1, -1, -1):
            t = th.full((n,), i, device=x_start.device, dtype=th.long)
            noise = _randn_like(x_start)
            x_t = self.q_sample(x_start, t, noise=noise)
            with th.no_grad():
                out = self._vb_terms_bpd(model, x_start, x_t, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
            vb.append(out["output"])
            xstart_mse.append(((out["pred_xstart"] - x_start) ** 2).flatten(1).mean(1))
            eps = ((self._coef(self.sqrt_recip_alphas_cumprod, t, x_t) * x_t - out["pred_xstart"])
                   / self._coef(self.sqrt_recipm1_alphas_cumprod, t, x_t))
            mse.append(((eps - noise) ** 2).flatten(1).mean(1))
        vb, xstart_mse, mse = th.stack(vb, dim=1), th.stack(xstart_mse, dim=1), th.stack(mse, dim=1)
        prior = self._prior_bpd(x_start)
        return {"total_bpd": vb.sum(dim=1) + prior, "prior_bpd": prior, "vb": vb, "xstart_mse": xstart_mse, "mse": mse}


class _ModelWrapperBase:
    pass
