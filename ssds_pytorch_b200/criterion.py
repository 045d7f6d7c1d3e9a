"""`MultiBoxLoss` — reference ssds/core/criterion.py:8-71, on the sm_100a kernels.

forward(pred_logits [B,A,C,H,W], target [B,A,C,H,W], depth [B,A,1,H,W]) -> unreduced loss
[B,A,C,H,W], with hard negatives mined per image (the reference's expand_as at criterion.py:66 only
runs for B == 1; per-image is the intended meaning).  Forward only (inference/timing
scope of this round); `forward_sum` is the fused variant that never materialises the one-hot target.
"""
import torch

from . import _lib
from ._lib import lib, check, ptr, dev_f32, stream_ptr


class MultiBoxLoss(torch.nn.Module):
    def __init__(self, negpos_ratio=3, **kwargs):
        super().__init__()
        self.negpos_ratio = negpos_ratio

    def forward(self, pred_logits, target, depth):
        logits = dev_f32(pred_logits)
        device = logits.device
        target = dev_f32(target, device)
        depth = dev_f32(depth, device)
        B, A, C, H, W = logits.shape
        out = torch.empty_like(logits)
        with torch.cuda.device(device):
            need = lib.ssdsb_multibox_loss_workspace_bytes(B, A, C, H, W)
            ws = _lib.workspace(need, device)
            check(lib.ssdsb_multibox_loss(ptr(logits), ptr(target), ptr(depth), B, A, C, H, W,
                                          int(self.negpos_ratio), ptr(out), ptr(ws), ws.numel(),
                                          stream_ptr()), "MultiBoxLoss")
        return out

    def forward_sum(self, pred_logits, depth):
        """Fused: per image (sum(loss * (depth >= 0)), #positives) — what
        pipeline_anchor_basic.py:76-82 reduces the unreduced loss to.  The class of a positive is
        depth-1 (valid when MATCHER.CENTER_SAMPLING_RADIUS == 0, the default)."""
        logits = dev_f32(pred_logits)
        device = logits.device
        depth = dev_f32(depth, device)
        B, A, C, H, W = logits.shape
        loss_sum = torch.empty((B,), dtype=torch.float32, device=device)
        num_pos = torch.empty((B,), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            need = lib.ssdsb_multibox_loss_workspace_bytes(B, A, C, H, W)
            ws = _lib.workspace(need, device)
            check(lib.ssdsb_multibox_loss_sum(ptr(logits), ptr(depth), B, A, C, H, W,
                                              int(self.negpos_ratio), ptr(loss_sum), ptr(num_pos),
                                              ptr(ws), ws.numel(), stream_ptr()),
                  "MultiBoxLoss.forward_sum")
        return loss_sum, num_pos


class _SumMixin:
    @staticmethod
    def _ws(B, A, H, W, device):
        return _lib.workspace(lib.ssdsb_loss_sum_workspace_bytes(B, A, H, W), device)


class FocalLoss(torch.nn.Module, _SumMixin):
    """reference criterion.py:74-108.  forward(pred_logits, target, depth) -> unreduced [B,A,C,H,W]
    (depth is unused there too); forward_sum(pred_logits, depth) -> per image
    (sum(loss * (depth >= 0)), #positives) with the class taken from depth (pipeline_anchor_basic.py:76-82)."""

    def __init__(self, alpha=0.25, gamma=2, **kwargs):
        super().__init__()
        self.alpha = alpha
        self.gamma = gamma

    def forward(self, pred_logits, target, depth=None):
        logits = dev_f32(pred_logits)
        target = dev_f32(target, logits.device)
        B, A, C, H, W = logits.shape
        out = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            check(lib.ssdsb_focal_loss(ptr(logits), ptr(target), B, A, C, H, W, float(self.alpha),
                                       float(self.gamma), ptr(out), stream_ptr()), "FocalLoss")
        return out

    def forward_sum(self, pred_logits, depth):
        logits = dev_f32(pred_logits)
        device = logits.device
        depth = dev_f32(depth, device)
        B, A, C, H, W = logits.shape
        loss_sum = torch.empty((B,), dtype=torch.float32, device=device)
        num_pos = torch.empty((B,), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            ws = self._ws(B, A, H, W, device)
            check(lib.ssdsb_focal_loss_sum(ptr(logits), ptr(depth), B, A, C, H, W, float(self.alpha),
                                           float(self.gamma), ptr(loss_sum), ptr(num_pos), ptr(ws),
                                           ws.numel(), stream_ptr()), "FocalLoss.forward_sum")
        return loss_sum, num_pos


class _LocLoss(torch.nn.Module, _SumMixin):
    _type = 0
    beta = 0.11

    def forward(self, pred, target):
        pred = dev_f32(pred)
        target = dev_f32(target, pred.device)
        B, A, four, H, W = target.shape
        pred = pred.reshape(target.shape)
        out = torch.empty((B, A, 4 if self._type == 0 else 1, H, W), dtype=torch.float32,
                          device=pred.device)
        with torch.cuda.device(pred.device):
            check(lib.ssdsb_loc_loss(ptr(pred), ptr(target), B, A, H, W, self._type, float(self.beta),
                                     ptr(out), stream_ptr()), type(self).__name__)
        return out

    def forward_sum(self, pred, target, depth):
        """per image sum(loss * (depth > 0)) — pipeline_anchor_basic.py:84-88."""
        pred = dev_f32(pred)
        device = pred.device
        target = dev_f32(target, device)
        depth = dev_f32(depth, device)
        B, A, four, H, W = target.shape
        pred = pred.reshape(target.shape)
        loss_sum = torch.empty((B,), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            ws = self._ws(B, A, H, W, device)
            check(lib.ssdsb_loc_loss_sum(ptr(pred), ptr(target), ptr(depth), B, A, H, W, self._type,
                                         float(self.beta), ptr(loss_sum), ptr(ws), ws.numel(),
                                         stream_ptr()), type(self).__name__ + ".forward_sum")
        return loss_sum


class SmoothL1Loss(_LocLoss):
    """reference criterion.py:111-151 (beta 0.11)."""

    def __init__(self, beta=0.11):
        super().__init__()
        self.beta = beta
        self._type = 0


class IOULoss(_LocLoss):
    """reference criterion.py:154-239; loss_type in {iou, giou, diou, ciou} on delta-format boxes."""
    TYPES = {"iou": 1, "giou": 2, "diou": 3, "ciou": 4}

    def __init__(self, loss_type="iou"):
        super().__init__()
        if loss_type not in self.TYPES:
            raise NotImplementedError(loss_type)        # criterion.py:231
        self.loss_type = loss_type
        self._type = self.TYPES[loss_type]
