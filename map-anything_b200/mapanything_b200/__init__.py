"""mapanything_b200: B200-native (sm_100a) implementation of MapAnything's feed-forward inference hot path.

    from mapanything_b200 import MapAnything, mapanything_config
    model = MapAnything(**mapanything_config()).to("cuda").eval()
    preds = model.infer(views)

All arithmetic runs in `libmapanything_b200.so` (hand-written CUDA, C ABI in include/mapanything_b200.h).
"""
from . import _lib  # noqa: F401
from .config import mapanything_config, mapanything_variant_config, pred_head_variant_config, tiny_config  # noqa: F401
from .image import load_images  # noqa: F401
from .model import MapAnything  # noqa: F401

__all__ = ["MapAnything", "load_images", "mapanything_config", "mapanything_variant_config", "pred_head_variant_config",
           "tiny_config"]
