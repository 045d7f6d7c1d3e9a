"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32 + autograd) of the loss path's backward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product (ssds_pytorch_b200/) never does.

Restates, in plain torch ops, the forward of the reference criteria
    MultiBoxLoss  ssds/core/criterion.py:43-71   (hard-negative mask taken from oracle/box_oracle.py)
    FocalLoss     :95-108
    SmoothL1Loss  :138-151
    IOULoss       :175-239   (iou / giou / diou / ciou; ciou's alpha under no_grad, :221-223)
and the caller's masking + normalisation (ssds/pipeline/pipeline_anchor_basic.py:76-97), and lets
autograd differentiate them.  Pinned: tests/test_oracle_golden.py compares these gradients with the ones
autograd gives on the REFERENCE modules themselves (tests/golden/box_ops.npz, keys ls*_g_* / mbl*_grad,
written by tests/golden/make_golden.py).
"""
import math

import numpy as np
import torch

from . import box_oracle as O


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def focal_loss_t(x, target, alpha=0.25, gamma=2):
    p = torch.sigmoid(x)
    ce = (1 - target) * x - torch.nn.functional.logsigmoid(x)        # BCEWithLogits, reduction none
    a = target * alpha + (1 - target) * (1 - alpha)
    pt = torch.where(target == 1, p, 1 - p)
    return a * (1.0 - pt) ** gamma * ce


def smooth_l1_t(pred, target, beta=0.11):
    x = (pred - target).abs()
    return torch.where(x >= beta, x - 0.5 * beta, 0.5 * x ** 2 / beta)


def iou_loss_t(pred, target, loss_type):
    def ltrb(d):
        wh = torch.exp(d[:, :, 2:])
        return d[:, :, :2] - 0.5 * wh, d[:, :, :2] + 0.5 * wh, wh

    plt, prb, pwh = ltrb(pred)
    tlt, trb, twh = ltrb(target)
    lt, rb = torch.max(plt, tlt), torch.min(prb, trb)
    area_i = torch.prod(rb - lt, dim=2) * (lt < rb).all(dim=2)
    area_u = torch.prod(pwh, dim=2) + torch.prod(twh, dim=2) - area_i
    iou = (area_i + 1e-7) / (area_u + 1e-7)
    if loss_type == "iou":
        return 1 - torch.clamp(iou, min=0, max=1.0).unsqueeze(2)
    olt, orb = torch.min(plt, tlt), torch.max(prb, trb)
    if loss_type == "giou":
        area_o = torch.prod(orb - olt, dim=2) * (olt < orb).all(dim=2) + 1e-7
        return 1 - torch.clamp(iou - (area_o - area_u) / area_o, min=-1.0, max=1.0).unsqueeze(2)
    inter = ((pred[:, :, :2] - target[:, :, :2]) ** 2).sum(dim=2)
    outer = ((orb - olt) ** 2).sum(dim=2) + 1e-7
    if loss_type == "diou":
        return 1 - torch.clamp(iou - inter / outer, min=-1.0, max=1.0).unsqueeze(2)
    if loss_type == "ciou":
        v = (4 / math.pi ** 2) * (torch.atan(twh[:, :, 0] / twh[:, :, 1]) - torch.atan(pwh[:, :, 0] / pwh[:, :, 1])) ** 2
        with torch.no_grad():
            alpha = v / ((1 - iou) + v)
        return 1 - torch.clamp(iou - (inter / outer + alpha * v), min=-1.0, max=1.0).unsqueeze(2)
    raise ValueError(loss_type)


def focal_sum_grad(logits, target, depth, scale, alpha=0.25, gamma=2):
    """d/dlogits of sum_b scale[b] * sum(focal * (depth >= 0))[b]."""
    x = _t(logits).requires_grad_(True)
    d = _t(depth)
    loss = focal_loss_t(x, _t(target), alpha, gamma) * (d >= 0).float()
    (loss.flatten(1).sum(1) * _t(scale)).sum().backward()
    return x.grad.numpy()


def loc_sum_grad(pred, target, depth, scale, loss_type="smoothl1", beta=0.11):
    """d/dpred of sum_b scale[b] * sum(loc_loss * (depth > 0))[b]."""
    p = _t(pred).requires_grad_(True)
    d = _t(depth)
    l = smooth_l1_t(p, _t(target), beta) if loss_type == "smoothl1" else iou_loss_t(p, _t(target), loss_type)
    ((l * (d > 0).float()).flatten(1).sum(1) * _t(scale)).sum().backward()
    return p.grad.numpy()


def multibox_sum_grad(logits, target, depth, scale, negpos_ratio=3):
    """d/dlogits of sum_b scale[b] * sum(MultiBoxLoss * (depth >= 0))[b]; the mined mask is a constant."""
    sel = (O.multibox_loss(logits, target, depth, negpos_ratio) != 0) | \
          (np.broadcast_to(np.asarray(depth) > 0, np.asarray(logits).shape))
    # (ce is > 0 for finite logits, so "!= 0" is the selection; positives are selected by definition)
    x = _t(logits).requires_grad_(True)
    ce = (1 - _t(target)) * x - torch.nn.functional.logsigmoid(x)
    loss = ce * torch.from_numpy(sel.astype(np.float32)) * (_t(depth) >= 0).float()
    (loss.flatten(1).sum(1) * _t(scale)).sum().backward()
    return x.grad.numpy()
