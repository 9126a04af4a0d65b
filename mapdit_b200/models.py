"""Model-size registry with the reference's names and kwargs (src/models.py:4-56)."""
from .dit import DiT

_SIZES = {"XL": dict(depth=28, hidden_size=1152, num_heads=16), "L": dict(depth=24, hidden_size=1024, num_heads=16),
          "B": dict(depth=12, hidden_size=768, num_heads=12), "S": dict(depth=12, hidden_size=384, num_heads=6),
          "XS": dict(depth=6, hidden_size=256, num_heads=4)}


def _make(size, patch):
    def ctor(**kwargs):
        return DiT(patch_size=patch, **_SIZES[size], **kwargs)
    ctor.__name__ = f"DiT_{size}_{patch}"
    return ctor


DIT_MODELS = {}
for _s in _SIZES:
    for _p in (2, 4, 8):
        _f = _make(_s, _p)
        globals()[_f.__name__] = _f
        DIT_MODELS[f"DiT-{_s}/{_p}"] = _f

# spelling used by north_star / upstream DiT
DiT_models = DIT_MODELS


def get_model(args):
    """utils.py:9-17 of the reference: build from a Namespace / dict with model, in_channels, input_size, num_classes."""
    if not isinstance(args, dict):
        args = vars(args)
    return DIT_MODELS[args["model"]](in_channels=args["in_channels"], input_size=args["input_size"],
                                     num_classes=args["num_classes"])
