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
