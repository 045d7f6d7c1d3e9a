"""`MultiBoxLoss` — reference ssds/core/criterion.py:8-71, on the sm_100a kernels.

forward(pred_logits [B,A,C,H,W], target [B,A,C,H,W], depth [B,A,1,H,W]) -> unreduced loss
[B,A,C,H,W], with hard negatives mined per image (the reference's expand_as at criterion.py:66 only
runs for B == 1; per-image is the intended meaning).  `forward_sum` is the fused variant
that never materialises the one-hot target; it is differentiable (torch.autograd.Function over the
backward kernels: dL/dlogits, dL/dloc — the mined mask and ciou's alpha are constants, as in the
reference).  The unreduced drop-in `forward` outputs are not connected to autograd.
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
        depth = dev_f32(depth, logits.device)
        if logits.requires_grad:
            return _MultiBoxSumFn.apply(logits, depth, int(self.negpos_ratio))
        return _multibox_sum(logits, depth, int(self.negpos_ratio))


def _multibox_sum(logits, depth, negpos_ratio):
    device = logits.device
    B, A, C, H, W = logits.shape
    loss_sum = torch.empty((B,), dtype=torch.float32, device=device)
    num_pos = torch.empty((B,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        need = lib.ssdsb_multibox_loss_workspace_bytes(B, A, C, H, W)
        ws = _lib.workspace(need, device)
        check(lib.ssdsb_multibox_loss_sum(ptr(logits), ptr(depth), B, A, C, H, W, negpos_ratio,
                                          ptr(loss_sum), ptr(num_pos), ptr(ws), ws.numel(), stream_ptr()),
              "MultiBoxLoss.forward_sum")
    return loss_sum, num_pos


class _MultiBoxSumFn(torch.autograd.Function):
    """(loss_sum [B], num_pos [B]) = f(logits); backward: ssdsb_multibox_loss_sum_backward."""

    @staticmethod
    def forward(ctx, logits, depth, negpos_ratio):
        loss_sum, num_pos = _multibox_sum(logits.detach(), depth, negpos_ratio)
        ctx.save_for_backward(logits.detach(), depth)
        ctx.negpos_ratio = negpos_ratio
        ctx.mark_non_differentiable(num_pos)
        return loss_sum, num_pos

    @staticmethod
    def backward(ctx, g_sum, _g_npos):
        logits, depth = ctx.saved_tensors
        B, A, C, H, W = logits.shape
        grad = torch.empty_like(logits)
        scale = g_sum.to(torch.float32).contiguous()
        with torch.cuda.device(logits.device):
            need = lib.ssdsb_multibox_loss_workspace_bytes(B, A, C, H, W)
            ws = _lib.workspace(need, logits.device)
            check(lib.ssdsb_multibox_loss_sum_backward(ptr(logits), ptr(depth), B, A, C, H, W,
                                                       ctx.negpos_ratio, ptr(scale), ptr(grad), ptr(ws),
                                                       ws.numel(), stream_ptr()), "MultiBoxLoss backward")
        return grad, None, None


def _focal_sum(logits, depth, alpha, gamma):
    device = logits.device
    B, A, C, H, W = logits.shape
    loss_sum = torch.empty((B,), dtype=torch.float32, device=device)
    num_pos = torch.empty((B,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        ws = _lib.workspace(lib.ssdsb_loss_sum_workspace_bytes(B, A, H, W), device)
        check(lib.ssdsb_focal_loss_sum(ptr(logits), ptr(depth), B, A, C, H, W, alpha, gamma, ptr(loss_sum),
                                       ptr(num_pos), ptr(ws), ws.numel(), stream_ptr()),
              "FocalLoss.forward_sum")
    return loss_sum, num_pos


class _FocalSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, depth, alpha, gamma):
        loss_sum, num_pos = _focal_sum(logits.detach(), depth, alpha, gamma)
        ctx.save_for_backward(logits.detach(), depth)
        ctx.hp = (alpha, gamma)
        ctx.mark_non_differentiable(num_pos)
        return loss_sum, num_pos

    @staticmethod
    def backward(ctx, g_sum, _g_npos):
        logits, depth = ctx.saved_tensors
        B, A, C, H, W = logits.shape
        grad = torch.empty_like(logits)
        scale = g_sum.to(torch.float32).contiguous()
        with torch.cuda.device(logits.device):
            check(lib.ssdsb_focal_loss_sum_backward(ptr(logits), ptr(depth), B, A, C, H, W, ctx.hp[0],
                                                    ctx.hp[1], ptr(scale), ptr(grad), stream_ptr()),
                  "FocalLoss backward")
        return grad, None, None, None


def _loc_sum(pred, target, depth, loss_type, beta):
    device = pred.device
    B, A, _, H, W = target.shape
    loss_sum = torch.empty((B,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        ws = _lib.workspace(lib.ssdsb_loss_sum_workspace_bytes(B, A, H, W), device)
        check(lib.ssdsb_loc_loss_sum(ptr(pred), ptr(target), ptr(depth), B, A, H, W, loss_type, beta,
                                     ptr(loss_sum), ptr(ws), ws.numel(), stream_ptr()), "loc loss forward_sum")
    return loss_sum


class _LocSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, depth, loss_type, beta):
        ctx.save_for_backward(pred.detach(), target, depth)
        ctx.hp = (loss_type, beta)
        return _loc_sum(pred.detach(), target, depth, loss_type, beta)

    @staticmethod
    def backward(ctx, g_sum):
        pred, target, depth = ctx.saved_tensors
        B, A, _, H, W = target.shape
        grad = torch.empty_like(pred)
        scale = g_sum.to(torch.float32).contiguous()
        with torch.cuda.device(pred.device):
            check(lib.ssdsb_loc_loss_sum_backward(ptr(pred), ptr(target), ptr(depth), B, A, H, W, ctx.hp[0],
                                                  ctx.hp[1], ptr(scale), ptr(grad), stream_ptr()),
                  "loc loss backward")
        return grad, None, None, None, None


class FocalLoss(torch.nn.Module):
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
        depth = dev_f32(depth, logits.device)
        if logits.requires_grad:
            return _FocalSumFn.apply(logits, depth, float(self.alpha), float(self.gamma))
        return _focal_sum(logits, depth, float(self.alpha), float(self.gamma))


class _LocLoss(torch.nn.Module):
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
        pred = pred.reshape(target.shape)
        if pred.requires_grad:
            return _LocSumFn.apply(pred, target, depth, self._type, float(self.beta))
        return _loc_sum(pred, target, depth, self._type, float(self.beta))


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


def GIOULoss(**kwargs):
    """reference criterion.py factory names: `getattr(criterion, cfg.MATCHER.LOCATE_LOSS)()` must resolve."""
    return IOULoss(loss_type="giou", **kwargs)


def DIOULoss(**kwargs):
    return IOULoss(loss_type="diou", **kwargs)


def CIOULoss(**kwargs):
    return IOULoss(loss_type="ciou", **kwargs)
