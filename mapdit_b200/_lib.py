"""ctypes binding of libmapdit.so (C ABI declared in include/mapdit.h).

There is no CPU fallback: importing the ops without the built library raises.  Build it with
``python -m mapdit_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmapdit.so")

F32, BF16 = 0, 1
EPI_STORE, EPI_QKNORM, EPI_MPSILU, EPI_RESID_MOD, EPI_RESID, EPI_SILU_BWD, EPI_RESID_ROT, EPI_STORE_DELTA = 0, 1, 2, 3, 4, 5, 6, 7

_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float


class GemmArgs(C.Structure):
    """mirror of mapdit_gemm_args (include/mapdit.h)"""
    _fields_ = [("a", _p), ("b", _p), ("out", _p), ("out2", _p), ("resid", _p), ("gate", _p), ("shift", _p),
                ("scale", _p), ("gain", _p), ("aux", _p), ("lda", _i64), ("ldb", _i64), ("ldo", _i64), ("ldmod", _i64),
                ("m", _i), ("n", _i), ("k", _i), ("tokens", _i), ("head_dim", _i), ("qk_cols", _i),
                ("epilogue", _i), ("out_dtype", _i), ("eps", _f), ("ldrot", _i64)]


SIGNATURES = {
    "mapdit_weight_norm_fwd": [_p, _i, _i, _f, _i, _p, _p, _p, _i64, _p, _p],
    "mapdit_weight_norm_bwd": [_p, _p, _p, _i, _i, _f, _i, _p],
    "mapdit_weight_norm_bwd_multi": [_p, _i, _i, _f, _p],
    "mapdit_adam_step_g16": [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _f, _f, _p],
    "mapdit_cast_2d": [_p, _i64, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_weight_norm_fwd_multi": [_p, _i, _i, _i, _f, _i, _p, _p],
    "mapdit_gemm_f32": [_p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_gemm_bf16": [C.POINTER(GemmArgs), _p],
    "mapdit_gemm_bf16_tn": [_p, _i64, _p, _i64, _p, _i64, _i, _i, _i, _p],
    "mapdit_multi_lerp": [_p, _i, _f, _p],
    "mapdit_set_variant": [_i],
    "mapdit_latent_sample": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "mapdit_timestep_sincos": [_p, _p, _i, _i, _f, _p],
    "mapdit_ln_modulate_fwd": [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_ln_modulate_bwd": [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p],
    "mapdit_adam_step": [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _f, _f, _p],
    "mapdit_modulate_fwd": [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_resid_fwd": [_p, _p, _p, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_mp_silu_fwd": [_p, _p, _i64, _i, _i, _p],
    "mapdit_qk_normalize": [_p, _i, _i, _i, _f, _i, _p],
    "mapdit_cast": [_p, _p, _i64, _i, _i, _p],
    "mapdit_cos_attn_fwd": [_p, _p, _p, _i, _i, _i, _i, _i, _p],
    "mapdit_cos_attn_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "mapdit_cos_attn_bwd_qknorm": [_p, _p, _p, _p, _p, _f, _p, _p, _i, _i, _i, _i, _i, _p],
    "mapdit_resid_bwd": [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_modulate_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p],
    "mapdit_modulate_resid_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p],
    "mapdit_sum_partials": [_p, _i, _p, _i, _p],
    "mapdit_mp_silu_bwd": [_p, _p, _p, _i64, _i, _p],
    "mapdit_qk_normalize_save": [_p, _p, _i, _i, _i, _f, _i, _p],
    "mapdit_qk_norm_bwd": [_p, _p, _p, _i, _i, _i, _f, _i, _p],
    "mapdit_final_bwd": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "mapdit_mp_scale_from_lin": [_p, _p, _p, _i, _i, _p],
    "mapdit_mp_scale_bwd": [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "mapdit_cond_combine_bwd": [_p, _p, _p, _p, _i64, _p],
    "mapdit_embed_rows_bwd": [_p, _p, _i64, _p, _p, _p, _i, _i, _f, _p],
    "mapdit_patchify": [_p, _p, _i, _i, _i, _i, _p],
    "mapdit_axpby": [_p, _p, _f, _i, _i64, _p],
    "mapdit_rotmod_fwd": [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _p],
    "mapdit_rotmod_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p],
    "mapdit_rotmod_resid_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _p],
    "mapdit_rot_table": [_p, _p, _p, _p, _p, _p, _i64, _i64, _i, _i, _p],
    "mapdit_patch_embed_wgrad": [_p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _p],
    "mapdit_patch_embed": [_p, _p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i, _i, _i, _i, _p],
    "mapdit_fourier": [_p, _p, _p, _p, _i, _i, _p],
    "mapdit_embed_rows": [_p, _p, _i64, _p, _p, _i, _i, _f, _p],
    "mapdit_cond_combine": [_p, _p, _p, _p, _p, _i64, _p],
    "mapdit_mp_scale": [_p, _p, _p, _p, _i, _i, _i, _p],
    "mapdit_final_unpatchify": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "mapdit_cfg_combine": [_p, _i, _i, _i, _f, _p],
    "mapdit_diffusion_step": [_p, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _p],
    "mapdit_ddim_step": [_p, _p, _p, _p, _p, _i, _p, _p, _i, _i, _i, _i, _f, _p],
    "mapdit_q_sample": [_p, _p, _p, _p, _i, _p, _i, _i, _p],
    "mapdit_loss_fwd_bwd": [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p],
    "mapdit_p_mean_variance": [_p, _p, _p, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p],
    "mapdit_posterior_mean": [_p, _p, _p, _p, _i, _p, _i, _i, _p],
    "mapdit_noise_add": [_p, _p, _p, _p, _p, _i, _i, _p],
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"mapdit_b200: {LIB_PATH} is missing — the CUDA kernels are the only implementation "
                "(no CPU fallback). Build with `python -m mapdit_b200.build`.")
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _i
        L.mapdit_last_error.restype = C.c_char_p
        L.mapdit_abi_version.restype = _i
        L.mapdit_launch_count.restype = _i64
        L.mapdit_modulate_bwd_partials.argtypes = [_i, _i]
        L.mapdit_set_option.argtypes = [C.c_char_p, _i]
        L.mapdit_set_option.restype = _i
        L.mapdit_modulate_bwd_partials.restype = _i
        L.mapdit_rotmod_bwd_partials.argtypes = [_i, _i]
        L.mapdit_rotmod_bwd_partials.restype = _i
        _lib = L
        for env, opt in (("MAPDIT_GEMM_2CTA", b"gemm_2cta"), ("MAPDIT_GEMM_2CTA_BN", b"gemm_2cta_bn"),
                         ("MAPDIT_GEMM_FUSED_RESID", b"gemm_fused_resid"), ("MAPDIT_ATTN_BWD_FUSED", b"attn_bwd_fused"),
                         ("MAPDIT_ATTN_V2", b"attn_v2")):  # developer A/B switches
            if os.environ.get(env) is not None:
                L.mapdit_set_option(opt, int(os.environ[env]))
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libmapdit {what} failed ({rc}): {lib().mapdit_last_error().decode()}")


def launch_count() -> int:
    return int(lib().mapdit_launch_count())


# kernels launched through CUDA-graph replays never pass through the C entry points again, so the
# Python side adds (kernels captured in the graph) x (replays) to the library's own counter.
_replayed = 0


def note_graph_replay(kernels_in_graph: int):
    global _replayed
    _replayed += int(kernels_in_graph)


def total_launches() -> int:
    return launch_count() + _replayed


def set_option(name: str, value: int):
    check(lib().mapdit_set_option(name.encode(), int(value)), "set_option")
