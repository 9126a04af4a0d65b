"""DiT with the reference's module tree and state-dict keys, executed by hand-written CUDA kernels.

Mirrors src/dit.py:12-118 (DiT, forward, forward_with_cfg) and the parameter-holding modules of
src/basic, src/layers, src/blocks so that ``state_dict()`` / ``load_state_dict()`` /
``named_parameters()`` / ``get_parameter()`` / ``copy.deepcopy`` behave like the reference
(SURVEY.md §A.1, §8(b)).  The forward itself is not composed from these modules: it is one pass
of the kernel schedule in ``engine.py``.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .engine import Engine

def mod_layout(modulation, D):
    """column offsets of one block's modulation vector (fp32 [N, width]); a = attention branch, m = MLP branch"""
    h = D // 2
    if modulation == "adaln":
        return dict(width=6 * D, shift_a=0, scale_a=D, gate_a=2 * D, shift_m=3 * D, scale_m=4 * D, gate_m=5 * D)
    if modulation == "rotation_scaling":
        return dict(width=5 * D, rot_a=0, scale_a=h, gate_a=h + D, rot_m=h + 2 * D, scale_m=2 * h + 2 * D, gate_m=2 * h + 3 * D)
    if modulation == "rotation":
        return dict(width=3 * D, rot_a=0, gate_a=h, rot_m=h + D, gate_m=2 * h + D)
    raise ValueError(modulation)


MOD_LAYOUTS = ("adaln", "rotation_scaling", "rotation")
MAP_FLAGS = ("use_cosine_attention", "use_weight_normalization", "use_forced_weight_normalization", "use_mp_residual",
             "use_mp_silu", "use_no_layernorm", "use_mp_pos_enc", "use_mp_embedding")


class MPLinear(nn.Module):
    """weight holder of src/basic/mp_linear.py:9-28 (gain is the python float 1.0)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.weight = nn.Parameter(torch.empty(out_dim, in_dim))
        nn.init.normal_(self.weight)


class MPLinearChunk(nn.Module):
    """src/basic/mp_linear.py:48-63"""

    def __init__(self, in_dim, out_dim, n_chunks):
        super().__init__()
        self.in_dim, self.n_chunks = in_dim, n_chunks
        self.weight = nn.Parameter(torch.empty(n_chunks * out_dim, in_dim))
        nn.init.normal_(self.weight)


class MPSiLU(nn.Module):
    """src/basic/mp_silu.py:5-7 (parameter-free; present so Sequential indices match)"""


class MPEmbedding(nn.Module):
    """src/basic/mp_embedding.py:8-13"""

    def __init__(self, num_embeddings, embedding_dim):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(num_embeddings, embedding_dim))
        nn.init.normal_(self.weight)


class MLP(nn.Module):
    """src/layers/mlp.py:7-22"""

    def __init__(self, in_dim, out_dim, mlp_ratio=4.0, hidden_dim=None):
        super().__init__()
        self.hidden_dim = int(in_dim * mlp_ratio) if hidden_dim is None else hidden_dim
        self.net = nn.Sequential(MPLinear(in_dim, self.hidden_dim), MPSiLU(), MPLinear(self.hidden_dim, out_dim))


class Attention(nn.Module):
    """src/layers/attention.py:9-25"""

    def __init__(self, in_dim, num_heads):
        super().__init__()
        assert in_dim % num_heads == 0
        self.num_heads, self.head_dim = num_heads, in_dim // num_heads
        self.qkv_proj = MPLinearChunk(in_dim, in_dim, 3)
        self.out_proj = MPLinear(in_dim, in_dim)


class DiTBlock(nn.Module):
    """src/blocks/dit_block.py:11-29"""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, modulation="adaln"):
        super().__init__()
        self.attn = Attention(hidden_size, num_heads)
        self.mlp = MLP(hidden_size, hidden_size, mlp_ratio=mlp_ratio)
        if modulation == "adaln":      # shift, scale, gate x 2 (src/blocks/dit_block.py:24-27)
            self.modulation = nn.Sequential(MPSiLU(), MPLinearChunk(hidden_size, hidden_size, 6))
        else:                          # rotation (D/2) [+ scale (D)] + gate (D), x 2 -- UNPINNED
            width = {"rotation_scaling": 5 * hidden_size, "rotation": 3 * hidden_size}[modulation]
            self.modulation = nn.Sequential(MPSiLU(), MPLinearChunk(hidden_size, width, 1))
        self.gain_msa = nn.Parameter(torch.tensor(0.0))
        self.gain_mlp = nn.Parameter(torch.tensor(0.0))


class MPScale(nn.Module):
    """src/blocks/final_layer.py:12-18"""

    def __init__(self, in_dim, angle_dim=8, zero_init=True):
        super().__init__()
        self.angle_dim = angle_dim
        self.linear = MPLinear(in_dim, angle_dim)
        self.reference = nn.Parameter(torch.zeros(angle_dim) if zero_init else torch.ones(angle_dim))


class FinalLayer(nn.Module):
    """src/blocks/final_layer.py:24-51"""

    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.linear = MPLinearChunk(hidden_size, patch_size * patch_size * out_channels, 2)
        self.modulation = nn.Sequential(MPSiLU(), MPLinearChunk(hidden_size, hidden_size, 2))
        self.gain_mod = nn.Parameter(torch.tensor(0.0))
        self.mean_scale = MPScale(hidden_size, zero_init=False)
        self.sigma_scale = MPScale(hidden_size, zero_init=True)


class MPFourier(nn.Module):
    """src/blocks/timestep_embedder.py:8-16"""

    def __init__(self, num_channels):
        super().__init__()
        self.register_buffer("scale", (2 * torch.pi * torch.randn(num_channels)).to(torch.float32))
        self.register_buffer("shift", (2 * torch.pi * torch.rand(num_channels)).to(torch.float32))


class TimestepEmbedder(nn.Module):
    """src/blocks/timestep_embedder.py:24-40"""

    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = MLP(frequency_embedding_size, hidden_size, hidden_dim=hidden_size)
        self.embedding = MPFourier(frequency_embedding_size)


class LabelEmbedder(nn.Module):
    """src/blocks/label_embedder.py:6-17"""

    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        self.embedding = MPEmbedding(num_classes + (1 if dropout_prob > 0 else 0), hidden_size)
        self.num_classes, self.dropout_prob = num_classes, dropout_prob


def _sincos_1d(dim, pos):
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embed(dim, grid, normalized=True):
    """constant table of src/pos_embed.py:4-61, row-normalised as in src/dit.py:45-47 (raw with use_mp_pos_enc=False)."""
    coords = np.arange(grid, dtype=np.float32)
    mesh = np.stack(np.meshgrid(coords, coords), axis=0).reshape(2, 1, grid, grid)
    emb = np.concatenate([_sincos_1d(dim // 2, mesh[0]), _sincos_1d(dim // 2, mesh[1])], axis=1)
    t = torch.from_numpy(emb).float().unsqueeze(0)
    if not normalized:
        return t
    norm = torch.linalg.vector_norm(t, dim=-1, keepdim=True)
    return t * math.sqrt(dim) / (norm + 1e-4)


class DiT(nn.Module):
    """Diffusion transformer (src/dit.py:12-62).  Extra keyword arguments select the MaP switches of the
    reference README (README.md:59-66); the defaults are the snapshot's behaviour (all on, MP-AdaLN
    modulation).  ``compute_dtype`` picks the kernel set: "bf16" (tcgen05 tensor-core path) or
    "fp32" (CUDA-core parity mode)."""

    def __init__(self, depth, hidden_size, patch_size, input_size=32, in_channels=3, num_heads=16, mlp_ratio=4.0,
                 class_dropout_prob=0.1, num_classes=1000, learn_sigma=True, compute_dtype="bf16", modulation="adaln",
                 **flags):
        super().__init__()
        for k in flags:
            if k not in MAP_FLAGS:
                raise TypeError(f"DiT got an unexpected keyword argument '{k}'")
        # README.md:59-66 switches.  "On" is the snapshot's hard-coded behaviour (pinned by the reference); the "off"
        # branches are this repo's restatement of the vanilla DiT ops (UNPINNED), same parameters/keys.
        self.flags = {k: bool(flags.get(k, True)) for k in MAP_FLAGS}
        if modulation not in MOD_LAYOUTS:
            raise ValueError(f"modulation must be one of {sorted(MOD_LAYOUTS)}")
        if not self.flags["use_no_layernorm"] and modulation != "adaln":
            raise NotImplementedError("use_no_layernorm=False (LayerNorm + adaLN) is only defined for modulation='adaln'")
        self.modulation = modulation
        if not learn_sigma:
            raise NotImplementedError("learn_sigma=False raises TypeError in the reference (src/blocks/final_layer.py:60-61)")
        assert compute_dtype in ("bf16", "fp32")
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels
        self.input_size = input_size
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.depth = depth
        self.num_classes = num_classes
        self.compute_dtype = compute_dtype

        self.x_embedder = MPLinear(patch_size * patch_size * in_channels + 1, hidden_size)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.y_embedder = LabelEmbedder(num_classes, hidden_size, class_dropout_prob)
        self.register_buffer("pos_embed", sincos_pos_embed(hidden_size, input_size // patch_size,
                                                           normalized=self.flags["use_mp_pos_enc"]))
        self.blocks = nn.ModuleList([DiTBlock(hidden_size, num_heads, mlp_ratio=mlp_ratio, modulation=modulation) for _ in range(depth)])
        self.final_layer = FinalLayer(hidden_size, patch_size, self.out_channels)
        if not self.flags["use_weight_normalization"]:
            # without weight normalisation N(0,1) weights would scale activations by sqrt(fan_in) per layer
            with torch.no_grad():
                for name, p in self.named_parameters():
                    if p.dim() == 2 and name != "y_embedder.embedding.weight":
                        p.mul_(1.0 / math.sqrt(p.shape[1]))
        self._engine = None

    @property
    def variant(self) -> int:
        """MAPDIT_VAR_* word (include/mapdit.h) of this model's switched-off MaP flags"""
        f = self.flags
        return ((0 if f["use_mp_residual"] else 1) | (0 if f["use_mp_silu"] else 2) | (0 if f["use_mp_pos_enc"] else 4)
                | (0 if f["use_mp_embedding"] else 8) | (0 if f["use_cosine_attention"] else 16))

    # the engine holds workspaces / CUDA graphs: never deep-copied or serialised with the module
    def __deepcopy__(self, memo):
        import copy
        eng, self._engine = self._engine, None
        try:
            cls = self.__class__
            new = cls.__new__(cls)
            memo[id(self)] = new
            for k, v in self.__dict__.items():
                setattr(new, k, copy.deepcopy(v, memo))
        finally:
            self._engine = eng
        return new

    def load_state_dict(self, state_dict, *args, **kwargs):
        """accepts the `_orig_mod.` key prefix of checkpoints written from the reference's torch.compile-wrapped model
        (train.py:46,124-128)"""
        pre = "_orig_mod."
        if any(k.startswith(pre) for k in state_dict):
            state_dict = {(k[len(pre):] if k.startswith(pre) else k): v for k, v in state_dict.items()}
        out = super().load_state_dict(state_dict, *args, **kwargs)
        if self._engine is not None:
            self._engine.invalidate()
        return out

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engine"] = None
        return st

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self)
        return self._engine

    def forward(self, x, t, y, drop_mask=None):
        """x [N,C,H,W] fp32, t [N] int64, y [N] int64 -> [N,2C,H,W] fp32 (src/dit.py:70-105).
        ``drop_mask`` (bool [N], optional) pins the train-mode label dropout of
        src/blocks/label_embedder.py:19-27; by default it is drawn with torch.rand like the reference."""
        return self.engine.forward(x, t, y, train=self.training, drop_mask=drop_mask)

    def forward_with_cfg(self, x, t, y, cfg_scale):
        """src/dit.py:107-118"""
        half = x[: len(x) // 2]
        out = self.engine.forward(torch.cat([half, half], dim=0), t, y, train=self.training)
        from . import ops
        ops.cfg_combine(out, self.in_channels, cfg_scale)
        return out
