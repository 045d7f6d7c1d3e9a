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


_LOC_KINDS = {None: -1, "none": -1, "SmoothL1Loss": 0, "IOULoss": 1, "GIOULoss": 2, "DIOULoss": 3, "CIOULoss": 4}
_CLS_KINDS = {"MultiBoxLoss": 0, "FocalLoss": 1}


def fused_loss_step(loc, conf, targets, anchors, num_classes, cls_criterion="MultiBoxLoss",
                    loc_criterion=None, match=(0.5, 0.4), negpos_ratio=3, alpha=0.25, gamma=2.0, beta=0.11,
                    with_targets=False, out=None):
    """pipeline_anchor_basic.py:62-97 — match + classification criterion + localisation criterion + masking +
    normalisation for ALL levels and images in ONE kernel launch (csrc/loss_step.cu, `ssdsb_detection_loss`).

    loc / conf: per-level raw head outputs [B, A*4, H, W] / [B, A*C, H, W] (model in training mode);
    targets [B,T,5]; anchors OrderedDict{stride: [A,4]}; criteria by the reference's config names
    (MATCHER.CLASSIFY_LOSS in {MultiBoxLoss, FocalLoss}; MATCHER.LOCATE_LOSS in {SmoothL1Loss, IOULoss, GIOULoss,
    DIOULoss, CIOULoss} or None to skip).  Returns (scalars [3] = cls_loss, loc_loss, fg_targets — a device tensor,
    no sync —, per-pair dict with 'cls_sum' / 'loc_sum' / 'num_pos' [L,B], and with_targets: 'depth' /
    'box_target' lists as extract_targets returns them).  Forward only (not connected to autograd)."""
    import ctypes as C
    from . import _lib
    from ._lib import lib, check, ptr, dev_f32, stream_ptr
    if cls_criterion not in _CLS_KINDS or loc_criterion not in _LOC_KINDS:
        raise ValueError(f"fused_loss_step: unknown criterion {cls_criterion!r} / {loc_criterion!r}")
    L = len(conf)
    conf = [dev_f32(c) for c in conf]
    device = conf[0].device
    B = conf[0].shape[0]
    loc_kind = _LOC_KINDS[loc_criterion]
    loc = [dev_f32(l, device) for l in loc] if loc_kind >= 0 else [None] * L
    targets = dev_f32(targets, device)
    T = targets.shape[1]
    levels = (_lib.LossLevel * L)()
    keep, depths, boxes = [], [], []
    for i, (c, l, (stride, anchor)) in enumerate(zip(conf, loc, anchors.items())):
        a = dev_f32(anchor, device)
        keep.append(a)
        A = a.shape[0]
        H, W = c.shape[-2:]
        if c.shape[1] != A * num_classes or (l is not None and (l.shape[1] != A * 4 or l.shape[-2:] != c.shape[-2:])):
            raise ValueError(f"fused_loss_step: level {i} shapes {tuple(c.shape)} do not match {A} anchors x "
                             f"{num_classes} classes")
        d = torch.empty((B, A, 1, H, W), dtype=torch.float32, device=device) if with_targets else None
        bt = torch.empty((B, A, 4, H, W), dtype=torch.float32, device=device) if with_targets else None
        depths.append(d)
        boxes.append(bt)
        levels[i] = _lib.LossLevel(c.data_ptr(), l.data_ptr() if l is not None else None, a.data_ptr(),
                                   d.data_ptr() if d is not None else None,
                                   bt.data_ptr() if bt is not None else None, A, num_classes, H, W, int(stride))
    scalars = out if out is not None else torch.empty(3, dtype=torch.float32, device=device)
    per_pair = torch.empty((3, L, B), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        need = lib.ssdsb_detection_loss_workspace_bytes(levels, L, B)
        ws = _lib.workspace(need, device)
        check(lib.ssdsb_detection_loss(levels, L, B, ptr(targets), T, float(match[0]), float(match[1]),
                                       _CLS_KINDS[cls_criterion], int(negpos_ratio), float(alpha), float(gamma),
                                       loc_kind, float(beta), ptr(scalars), ptr(per_pair[0]), ptr(per_pair[1]),
                                       ptr(per_pair[2]), ptr(ws), ws.numel(), stream_ptr()), "fused_loss_step")
    parts = {"cls_sum": per_pair[0], "loc_sum": per_pair[1], "num_pos": per_pair[2]}
    if with_targets:
        parts["depth"], parts["box_target"] = depths, boxes
    return scalars, parts


class LossStep(object):
    """Host-facing training-step front half (pipeline_anchor_basic.py:56-97 up to the loss scalars): a pinned
    host batch (uint8 NHWC images + [B,T,5] targets) -> H2D -> conv stack in training mode (raw logits) ->
    match + hard-negative-mined MultiBoxLoss (or Focal + SmoothL1) -> the loss scalars back on the host.

        step = LossStep(cfg, state_dict)                 # cfg: the reference's yml keys (see ssds.load_cfg)
        out = step.loss_host(images, targets)            # pinned [3] (cls_loss, loc_loss, fg_targets), after step.sync()

    One process per GPU; ranks normalise by their local foreground count like the reference
    (pipeline_anchor_apex.py:69-71), so there is no collective on this path."""

    def __init__(self, cfg, state_dict, device=None, use_graph=True, cls_criterion="MultiBoxLoss",
                 loc_criterion=None):
        from .ssds import load_cfg
        from .model import engine_for, create_anchors, number_box_from_cfg
        cfg = load_cfg(cfg)
        m = cfg["MODEL"]
        self.device = (torch.device(device) if device is not None
                       else torch.device("cuda", torch.cuda.current_device()))
        self.model = engine_for(m["SSDS"], m["NETS"])(
            state_dict, m["FEATURE_LAYER"], m["NUM_CLASSES"], number_box_from_cfg(m), device=self.device,
            mean=float(cfg["DATASET"]["PREPROC"]["MEAN"]), std=float(cfg["DATASET"]["PREPROC"]["STD"]))
        self.model.eval()
        self.anchors = create_anchors(m, self.model, m["IMAGE_SIZE"])
        self.model.train()
        self.num_classes = m["NUM_CLASSES"]
        self.use_graph = use_graph
        self.cls_criterion, self.loc_criterion = cls_criterion, loc_criterion
        self._stage = {}

    def loss_device(self, images, targets):
        """-> device tensor [3] = (cls_loss, loc_loss, fg_targets); loc_loss is 0 when loc_criterion is None."""
        loc, conf = self.model(images, use_graph=self.use_graph)
        scalars, _ = fused_loss_step(loc, conf, targets, self.anchors, self.num_classes, self.cls_criterion,
                                     self.loc_criterion)
        return scalars

    def loss_host(self, images, targets):
        """Asynchronous: returns a pinned [3] tensor that is valid after `sync()`.  Two device staging sets and a copy
        stream: the host->device copy of this call's batch runs while the previous call's step is still computing
        (the copy of a 16 x 640 x 640 batch is ~0.45 ms of a 3.9 ms step when serialised behind it)."""
        images, targets = torch.as_tensor(images), torch.as_tensor(targets)
        key = (tuple(images.shape), images.dtype, tuple(targets.shape))
        st = self._stage.get(key)
        if st is None:
            st = {"img": [torch.empty(images.shape, dtype=images.dtype, device=self.device) for _ in range(2)],
                  "tg": [torch.empty(targets.shape, dtype=torch.float32, device=self.device) for _ in range(2)],
                  "out": [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)],
                  "copied": [torch.cuda.Event() for _ in range(2)], "done": [None, None],
                  "stream": torch.cuda.Stream(device=self.device), "n": 0}
            self._stage[key] = st
        i = st["n"] & 1
        st["n"] += 1
        main = torch.cuda.current_stream(self.device)
        pinned = images.is_pinned() and targets.is_pinned()
        with torch.cuda.stream(st["stream"]):
            if st["done"][i] is not None:
                st["stream"].wait_event(st["done"][i])       # the step that last read this staging set has finished
            st["img"][i].copy_(images, non_blocking=True)
            st["tg"][i].copy_(targets, non_blocking=True)
            st["copied"][i].record(st["stream"])
        if not pinned:
            # a pageable source is staged by the driver: the caller may reuse its buffer as soon as we return only if
            # the copy has completed (same guard as SSDDetector.detect_host)
            st["copied"][i].synchronize()
        main.wait_event(st["copied"][i])
        st["out"][i].copy_(self.loss_device(st["img"][i], st["tg"][i]), non_blocking=True)
        st["done"][i] = torch.cuda.Event()
        st["done"][i].record(main)
        return st["out"][i]

    def sync(self):
        torch.cuda.current_stream(self.device).synchronize()
