"""ssds_pytorch_b200 — B200-native (sm_100a) hot path of ShuangXieIrene/ssds.pytorch.

Importing this package loads libssdsb200.so; there is no CPU or PyTorch-eager fallback.
"""
from . import _lib                      # noqa: F401  (raises if the CUDA library is missing)
from .box import (configure_ratio_scale, generate_anchors, anchor_grid, box2delta, delta2box,  # noqa: F401
                  decode, decode_levels, nms, extract_targets)
from .decoder import Decoder            # noqa: F401
from .criterion import MultiBoxLoss, FocalLoss, SmoothL1Loss, IOULoss     # noqa: F401
from . import _C                        # noqa: F401

__version__ = "0.1.0"


def install(ssds_module=None):
    """Expose the extension as `ssds._C` (the seam at reference box.py:3-4, export.py:134-139)."""
    import sys
    if ssds_module is None:
        import ssds as ssds_module      # the reference package, if it is importable
    ssds_module._C = _C
    sys.modules[ssds_module.__name__ + "._C"] = _C
    return _C
