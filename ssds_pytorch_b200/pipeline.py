"""The loss half of the training step, as the reference sequences it
(ssds/pipeline/pipeline_anchor_basic.py:62-97, same body in pipeline_anchor_apex.py:37-72), on the
fused kernels: per level `extract_targets` (match + encode + depth, no one-hot target) followed by the
fused `MultiBoxLoss` reduction.  The returned losses are differentiable w.r.t. the head outputs passed
in (dL/dlogits, dL/dloc through the backward kernels); the conv stack itself has no backward yet.

    cls_loss, parts = multibox_cls_loss_step(conf_logits, targets, anchors, num_classes)

conf_logits: tuple of per-level raw logits [B, A*C, H, W] (the model in training mode, ssd.py:72-73);
targets [B, T, 5] = (x, y, w, h, label) padded with -1 rows; anchors OrderedDict{stride: [A,4]}.
Returns the reference's `cls_loss = sum_l sum(loss_l * (depth_l >= 0)) / sum_l max(#(depth_l > 0), 1)`
(pipeline_anchor_basic.py:76-97) and, per level, (loss_sum [B], num_pos [B], box_target, depth) so that
a localisation criterion can consume box_target / depth.
"""
import torch

from .box import extract_targets
from .criterion import MultiBoxLoss


def multibox_cls_loss_step(conf_logits, targets, anchors, num_classes, match=(0.5, 0.4),
                           negpos_ratio=3, center_sampling_radius=0):
    if center_sampling_radius > 0:
        raise NotImplementedError("the fused loss takes the class from depth, which ATSS centre sampling "
                                  "decouples (box.py:184-191); use extract_targets + MultiBoxLoss.forward")
    crit = MultiBoxLoss(negpos_ratio)
    total = None
    fg_total = None
    parts = []
    for c, (stride, anchor) in zip(conf_logits, anchors.items()):
        B, AC, H, W = c.shape
        A = anchor.shape[0]
        _, box_t, depth = extract_targets(targets, anchors, num_classes, stride, (H, W), list(match),
                                          center_sampling_radius, with_cls_target=False)
        loss_sum, num_pos = crit.forward_sum(c.view(B, A, AC // A, H, W), depth)
        fg = num_pos.sum().clamp(min=1)                       # pipeline_anchor_basic.py:76
        total = loss_sum.sum() if total is None else total + loss_sum.sum()
        fg_total = fg if fg_total is None else fg_total + fg
        parts.append((loss_sum, num_pos, box_t, depth))
    return total / fg_total, parts


def detection_loss_step(loc, conf, targets, anchors, num_classes, cls_criterion=None, loc_criterion=None,
                        match=(0.5, 0.4), center_sampling_radius=0):
    """Both halves of pipeline_anchor_basic.py:62-97 on the fused reductions.

    loc / conf: per-level raw head outputs [B, A*4, H, W] / [B, A*C, H, W] (model in training mode);
    cls_criterion: `FocalLoss` (the reference default, config.py:151) or `MultiBoxLoss`;
    loc_criterion: `SmoothL1Loss` (default, config.py:152) or `IOULoss`.
    Returns (cls_loss, loc_loss, fg_targets) — the two scalars the reference sums into `loss`.
    One divergence, on purpose: a NaN produced at an anchor the mask removes (ciou of two identical
    boxes at depth <= 0) does not reach the sum, whereas `mask * loss` would carry it.
    """
    from .criterion import FocalLoss, SmoothL1Loss
    if center_sampling_radius > 0:
        raise NotImplementedError("fused sums take the class from depth; ATSS centre sampling decouples "
                                  "them (box.py:184-191) — use extract_targets + the unreduced criteria")
    cls_criterion = cls_criterion or FocalLoss()
    loc_criterion = loc_criterion or SmoothL1Loss()
    cls_sum, loc_sum, fg = [], [], []
    for l, c, (stride, anchor) in zip(loc, conf, anchors.items()):
        B, AC, H, W = c.shape
        A = anchor.shape[0]
        _, box_t, depth = extract_targets(targets, anchors, num_classes, stride, (H, W), list(match),
                                          center_sampling_radius, with_cls_target=False)
        cs, npos = cls_criterion.forward_sum(c.view(B, A, AC // A, H, W), depth)
        ls = loc_criterion.forward_sum(l.view(B, A, 4, H, W), box_t, depth)
        cls_sum.append(cs.sum())
        loc_sum.append(ls.sum())
        fg.append(npos.sum().clamp(min=1))
    fg_targets = torch.stack(fg).sum()
    return torch.stack(cls_sum).sum() / fg_targets, torch.stack(loc_sum).sum() / fg_targets, fg_targets
