"""Host-side mirror of the reference's ssds/modeling/layers/box.py operator surface, bound to the
sm_100a kernels through the C ABI.  Same names, argument meaning and error behaviour:

    configure_ratio_scale  box.py:8-43     (host logic, raises the same ValueErrors)
    generate_anchors       box.py:46-58
    box2delta / delta2box  box.py:61-87
    extract_targets        box.py:362-405  (IoU matcher; the by-scale matcher is out of scope)
    decode                 box.py:408-477
    nms                    box.py:480-546

Inputs may live on any device; results are CUDA fp32 tensors (the reference returns tensors on the
input's device; the input of this path is the GPU model's output).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import lib, check, ptr, dev_f32, stream_ptr


def configure_ratio_scale(num_featmaps, ratios, scales):
    """reference box.py:8-43 (pure host logic, same exceptions)."""
    if len(scales) != num_featmaps:
        raise ValueError(
            "cfg.SIZES is not correct,"
            "the len of cfg.SIZES should equal to num layers({}) or 2, but it is {}".format(
                num_featmaps, len(scales)))
    scales = list(scales)
    for i in range(num_featmaps):
        if not isinstance(scales[i], list):
            scales[i] = [scales[i]]
    if isinstance(ratios[0], list):
        if len(ratios) != num_featmaps:
            raise ValueError(
                "When cfg.ASPECT_RATIOS contains list for each layer,"
                "Len of cfg.ASPECT_RATIOS should equal to num layers({}), but it is {}".format(
                    num_featmaps, len(ratios)))
    else:
        ratios = [ratios for _ in range(num_featmaps)]
    return ratios, scales


def _device(device=None):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device())


def generate_anchors(stride, ratio_vals, scales_vals, device=None):
    """Base anchors [A,4] on the GPU — reference box.py:46-58."""
    device = _device(device)
    nr, ns = len(ratio_vals), len(scales_vals)
    out = torch.empty((nr * ns, 4), dtype=torch.float32, device=device)
    r = (C.c_float * nr)(*[float(v) for v in ratio_vals])
    s = (C.c_float * ns)(*[float(v) for v in scales_vals])
    with torch.cuda.device(device):
        check(lib.ssdsb_generate_anchors(int(stride), r, nr, s, ns, ptr(out), stream_ptr()),
              "generate_anchors")
    return out


def anchor_grid(base_anchors, stride, width, height):
    """Materialised grid [A, W, H, 4] in the reference's x-major order (box.py:151-159)."""
    base = dev_f32(base_anchors)
    A = base.shape[0]
    out = torch.empty((A, width, height, 4), dtype=torch.float32, device=base.device)
    with torch.cuda.device(base.device):
        check(lib.ssdsb_anchor_grid(ptr(base), A, int(stride), int(width), int(height), ptr(out),
                                    stream_ptr()), "anchor_grid")
    return out


def box2delta(boxes, anchors):
    """reference box.py:61-71."""
    boxes = dev_f32(boxes)
    anchors = dev_f32(anchors, boxes.device)
    out = torch.empty_like(boxes)
    with torch.cuda.device(boxes.device):
        check(lib.ssdsb_box2delta(ptr(boxes), ptr(anchors), boxes.shape[0], ptr(out), stream_ptr()),
              "box2delta")
    return out


def delta2box(deltas, anchors, size, stride):
    """reference box.py:74-87; size = [W_feat, H_feat]."""
    deltas = dev_f32(deltas)
    anchors = dev_f32(anchors, deltas.device)
    out = torch.empty_like(deltas)
    with torch.cuda.device(deltas.device):
        check(lib.ssdsb_delta2box(ptr(deltas), ptr(anchors), deltas.shape[0], int(size[0]),
                                  int(size[1]), int(stride), ptr(out), stream_ptr()), "delta2box")
    return out


def decode_levels(conf, loc, anchors_items, threshold, top_n, rescore=True, return_indices=False):
    """All levels of all images in two launches (decoder.py:36-48 without the python loops).

    conf/loc: sequences of [B, A*C, H, W] / [B, A*4, H, W]; anchors_items: sequence of
    (stride, base_anchors[A,4]).  Returns tensors already concatenated along dim 1.
    """
    L = len(conf)
    if L < 1 or L > _lib.SSDSB_MAX_LEVELS:
        raise ValueError(f"decode: {L} levels outside [1,{_lib.SSDSB_MAX_LEVELS}]")
    conf = [dev_f32(c) for c in conf]
    device = conf[0].device
    loc = [dev_f32(l, device) for l in loc]
    B = conf[0].shape[0]
    levels = (_lib.Level * L)()
    keep = []
    for i, (c, l, (stride, anchor)) in enumerate(zip(conf, loc, anchors_items)):
        a = dev_f32(anchor, device)
        keep.append(a)
        A = a.shape[0]
        H, W = c.shape[-2:]
        if c.shape[1] % A != 0 or l.shape[1] != A * 4 or l.shape[-2:] != c.shape[-2:]:
            raise ValueError(f"decode: level {i} shapes {tuple(c.shape)} / {tuple(l.shape)} do not "
                             f"match {A} anchors")
        levels[i] = _lib.Level(c.data_ptr(), l.data_ptr(), a.data_ptr(), A, c.shape[1] // A,
                               H, W, int(stride))
    top_n = int(top_n)
    scores = torch.empty((B, L * top_n), dtype=torch.float32, device=device)
    boxes = torch.empty((B, L * top_n, 4), dtype=torch.float32, device=device)
    classes = torch.empty((B, L * top_n), dtype=torch.float32, device=device)
    index = torch.empty((B, L * top_n), dtype=torch.int32, device=device) if return_indices else None
    with torch.cuda.device(device):
        need = lib.ssdsb_decode_workspace_bytes(levels, L, B, top_n)
        ws = _lib.workspace(need, device)
        check(lib.ssdsb_decode(levels, L, B, float(threshold), top_n, int(bool(rescore)),
                               ptr(scores), ptr(boxes), ptr(classes), ptr(index), ptr(ws),
                               ws.numel(), stream_ptr()), "decode")
    if return_indices:
        return scores, boxes, classes, index
    return scores, boxes, classes


def decode(all_cls_head, all_box_head, stride=1, threshold=0.05, top_n=1000, anchors=None,
           rescore=True, return_indices=False):
    """Box decoding and filtering for one level — reference box.py:408-477."""
    if anchors is None:
        raise ValueError("decode: anchors are required")
    return decode_levels([all_cls_head], [all_box_head], [(stride, anchors)], threshold, top_n,
                         rescore, return_indices)


def nms(all_scores, all_boxes, all_classes, nms=0.5, ndetections=100, using_diou=True,
        return_indices=False, packed_out=None):
    """Non maximum suppression — reference box.py:480-546.
    packed_out: optional [B,D,6] fp32 CUDA tensor that additionally receives (score, x1, y1, x2, y2, class)
    per detection straight from the kernel (no packing launches); `packed_out=True` allocates it and returns
    ONLY that tensor."""
    scores = dev_f32(all_scores)
    device = scores.device
    boxes = dev_f32(all_boxes, device)
    classes = dev_f32(all_classes, device)
    B, N = scores.shape
    D = int(ndetections)
    only_packed = packed_out is True
    if only_packed:
        packed_out = torch.empty((B, D, 6), dtype=torch.float32, device=device)
    if packed_out is not None and (tuple(packed_out.shape) != (B, D, 6) or packed_out.dtype != torch.float32 or
                                   not packed_out.is_contiguous()):
        raise ValueError("nms: packed_out must be a contiguous fp32 [B, ndetections, 6] tensor")
    out_s = out_b = out_c = None
    if not only_packed:
        out_s = torch.empty((B, D), dtype=torch.float32, device=device)
        out_b = torch.empty((B, D, 4), dtype=torch.float32, device=device)
        out_c = torch.empty((B, D), dtype=torch.float32, device=device)
    out_i = torch.empty((B, D), dtype=torch.int32, device=device) if return_indices else None
    with torch.cuda.device(device):
        need = lib.ssdsb_nms_workspace_bytes(B, N, D)          # > 0 for long rows (multi-CTA pre-selection)
        ws = _lib.workspace(need, device) if need else None
        check(lib.ssdsb_nms(ptr(scores), ptr(boxes), ptr(classes), B, N, float(nms), D,
                            int(bool(using_diou)), ptr(out_s), ptr(out_b), ptr(out_c), ptr(out_i),
                            ptr(packed_out), ptr(ws), ws.numel() if ws is not None else 0, stream_ptr()), "nms")
    if only_packed:
        return packed_out
    if return_indices:
        return out_s, out_b, out_c, out_i
    return out_s, out_b, out_c


def extract_targets(targets, anchors, classes, stride, size, match=[0.5, 0.4],
                    center_sampling_radius=0, is_centerness=False, with_cls_target=True):
    """Snap the targets to anchors — reference box.py:362-405 (IoU matcher).

    targets [B,T,5] = (x,y,w,h,label) padded with -1 rows; size = (H_feat, W_feat).
    Returns (cls_target [B,A,C,H,W], box_target [B,A,4,H,W], depth [B,A,1,H,W]).
    """
    if isinstance(match[0], list):
        raise NotImplementedError("snap_to_anchors_by_scale (box.py:229-359) is out of scope")
    if not isinstance(match[0], float):
        raise ValueError("unvalidate match param")            # box.py:402
    if is_centerness:
        raise NotImplementedError("is_centerness targets are out of scope")
    targets = dev_f32(targets)
    device = targets.device
    base = dev_f32(anchors[stride], device)
    B, T = targets.shape[0], targets.shape[1]
    A = base.shape[0]
    H, W = int(size[0]), int(size[1])
    cls_t = (torch.empty((B, A, classes, H, W), dtype=torch.float32, device=device)
             if with_cls_target else None)
    box_t = torch.empty((B, A, 4, H, W), dtype=torch.float32, device=device)
    depth = torch.empty((B, A, 1, H, W), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib.ssdsb_match_iou(ptr(targets), B, T, ptr(base), A, int(classes), int(stride), H, W,
                                  float(match[0]), float(match[1]), float(center_sampling_radius),
                                  ptr(cls_t), ptr(box_t), ptr(depth), stream_ptr()),
              "extract_targets")
    return cls_t, box_t, depth
