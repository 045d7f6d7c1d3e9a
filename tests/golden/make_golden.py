"""Generate golden vectors by RUNNING THE REFERENCE's own functions (CPU, fp32).

Run in the authoring container only (needs /root/reference, which does not
travel to the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/box_ops.npz.  Inputs are stored next to the outputs so the
replaying tests (tests/test_oracle_golden.py on CPU, tests/test_gpu_*.py on the
B200) need neither the reference nor RNG reproducibility.

Reference functions executed (ssds/modeling/layers/box.py @ b5ec682):
generate_anchors :46-58, box2delta :61-71, delta2box :74-87, decode :408-477,
nms :480-546, extract_targets :362-405; decoder.py:25-49 Decoder;
ssds/core/criterion.py:43-71 MultiBoxLoss (called per image, B=1 slices),
:95-108 FocalLoss, :138-151 SmoothL1Loss, :175-239 IOULoss (iou/giou/diou/ciou).
All random inputs are tie-free by construction (distinct float32 scores).
"""
import os
import sys
import warnings
from collections import OrderedDict

import numpy as np
import torch

REF = os.environ.get("SSDS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from ssds.modeling.layers import box as rbox            # noqa: E402
from ssds.modeling.layers.decoder import Decoder        # noqa: E402
from ssds.core.criterion import MultiBoxLoss, FocalLoss, SmoothL1Loss, IOULoss   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "box_ops.npz")
G = {}


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def distinct_scores(rng, shape, lo=0.0, hi=1.0):
    # a jittered, randomly permuted grid: distinct float32 values by construction
    n = int(np.prod(shape))
    v = (rng.permutation(n) + rng.uniform(0.25, 0.75, size=n)) / n
    v = (lo + (hi - lo) * v).astype(np.float32)
    assert np.unique(v).size == n
    return v.reshape(shape)


def gen_anchors():
    cases = [
        (8, [1, 2, 0.5], [4.0, 5.04, 6.35]),
        (8, [1, 2], [2.0]),
        (15, [1, 2, 0.5], [2.0]),
        (15, [1, 2, 0.5], [2.0, 2.828]),
        (30, [1, 2, 0.5], [2.0, 2.828]),
        (100, [1, 2, 0.5], [2.0, 2.828]),
        (256, [1, 2, 0.5], [2.0, 2.828]),
        (128, [1, 2, 0.5], [4.0]),
        (32, [1, 2, 0.5], [2.0, 4.0, 8.0]),
        (7, [0.3, 1, 3.7], [1.5, 2.5]),
    ]
    for i, (s, r, sc) in enumerate(cases):
        G[f"anc{i}_stride"] = np.int64(s)
        G[f"anc{i}_ratios"] = np.asarray(r, np.float64)
        G[f"anc{i}_scales"] = np.asarray(sc, np.float64)
        G[f"anc{i}_out"] = rbox.generate_anchors(s, list(r), list(sc)).numpy()
    G["anc_n"] = np.int64(len(cases))


def gen_codec(rng):
    n = 257
    anchors = np.stack([rng.uniform(-50, 400, n), rng.uniform(-50, 400, n)], 1)
    anchors = np.concatenate([anchors, anchors + rng.uniform(3, 200, (n, 2))], 1).astype(np.float32)
    xy = rng.uniform(0, 400, (n, 2))
    boxes = np.concatenate([xy, xy + rng.uniform(1, 300, (n, 2))], 1).astype(np.float32)
    deltas = rng.normal(0, 0.5, (n, 4)).astype(np.float32)
    G["codec_anchors"], G["codec_boxes"], G["codec_deltas"] = anchors, boxes, deltas
    G["codec_box2delta"] = rbox.box2delta(t(boxes), t(anchors)).numpy()
    G["codec_delta2box"] = rbox.delta2box(t(deltas), t(anchors), [40, 30], 16).numpy()


def gen_decode(rng):
    cases = [
        # B, A, C, H, W, stride, thr, top_n, rescore
        (2, 3, 5, 6, 7, 8, 0.30, 20, True),
        (2, 3, 5, 6, 7, 8, 0.30, 20, False),
        (3, 6, 80, 10, 10, 30, 0.01, 300, True),     # cfg1b level-1 geometry
        (1, 2, 4, 3, 5, 16, 0.90, 50, True),         # fewer passing than top_n
        (2, 6, 7, 1, 1, 300, 0.01, 300, True),       # 1x1 level
        (1, 9, 3, 5, 5, 128, 0.50, 10, True),
        (2, 1, 3, 4, 4, 8, 2.0, 10, True),           # nothing passes -> zero rows
    ]
    for i, (B, A, C, H, W, stride, thr, top_n, rescore) in enumerate(cases):
        conf = distinct_scores(rng, (B, A * C, H, W))
        loc = rng.normal(0, 0.5, (B, A * 4, H, W)).astype(np.float32)
        ratios = [1, 2, 0.5, 3, 1.5, 0.7, 0.4, 2.5, 0.3][:A] if A not in (6, 9) else [1, 2, 0.5]
        scales = [2.0] if A not in (6, 9) else ([2.0, 2.828] if A == 6 else [4.0, 5.04, 6.35])
        anchors = rbox.generate_anchors(stride, ratios, scales)
        assert anchors.shape[0] == A
        s, b, c = rbox.decode(t(conf), t(loc), stride, thr, top_n, anchors, rescore)
        p = f"dec{i}_"
        G[p + "conf"], G[p + "loc"], G[p + "anchors"] = conf, loc, anchors.numpy()
        G[p + "params"] = np.asarray([stride, thr, top_n, int(rescore)], np.float64)
        G[p + "scores"], G[p + "boxes"], G[p + "classes"] = s.numpy(), b.numpy(), c.numpy()
    G["dec_n"] = np.int64(len(cases))


def rand_dets(rng, B, N, ncls, img=300.0, frac_zero=0.1, cluster=True):
    scores = distinct_scores(rng, (B, N), 0.01, 1.0)
    zero = rng.uniform(size=(B, N)) < frac_zero
    scores[zero] = 0.0
    if cluster:  # boxes around a few centres so that suppression actually happens
        ctr = rng.uniform(30, img - 30, (B, 6, 2))
        pick = rng.integers(0, 6, (B, N))
        c = np.take_along_axis(ctr, pick[..., None].repeat(2, -1), 1) + rng.normal(0, 6, (B, N, 2))
    else:
        c = rng.uniform(0, img, (B, N, 2))
    wh = rng.uniform(8, 90, (B, N, 2))
    boxes = np.concatenate([c - wh / 2, c + wh / 2], -1)
    boxes = np.clip(boxes, 0, img - 1).astype(np.float32)
    classes = rng.integers(0, ncls, (B, N)).astype(np.float32)
    return scores, boxes, classes


def gen_nms(rng):
    cases = [
        # B, N, ncls, thr, D, diou, cluster
        (3, 200, 3, 0.6, 100, True, True),
        (3, 200, 3, 0.6, 100, False, True),
        (2, 500, 1, 0.5, 30, True, True),
        (2, 64, 80, 0.6, 100, True, False),      # fewer survivors than D
        (1, 1800, 4, 0.6, 100, True, True),      # default candidate count
        (2, 40, 2, 0.05, 100, True, True),       # aggressive threshold
        (1, 10, 2, 0.6, 100, True, True),
    ]
    for i, (B, N, ncls, thr, D, diou, cl) in enumerate(cases):
        s, b, c = rand_dets(rng, B, N, ncls, cluster=cl)
        if i == 6:
            s[:] = 0.0                            # nothing valid -> zero rows
        os_, ob, oc = rbox.nms(t(s), t(b), t(c), thr, D, diou)
        p = f"nms{i}_"
        G[p + "scores"], G[p + "boxes"], G[p + "classes"] = s, b, c
        G[p + "params"] = np.asarray([thr, D, int(diou)], np.float64)
        G[p + "out_scores"], G[p + "out_boxes"], G[p + "out_classes"] = (
            os_.numpy(), ob.numpy(), oc.numpy())
    G["nms_n"] = np.int64(len(cases))


def gen_decoder(rng):
    # three-level Decoder, cfg1b-like geometry (strides 15/30/60 at 300 px), B=2
    B, C = 2, 80
    levels = [(15, 19), (30, 10), (60, 5)]
    anchors = OrderedDict()
    loc, conf = [], []
    for li, (stride, hw) in enumerate(levels):
        a = rbox.generate_anchors(stride, [1, 2, 0.5], [2.0, 2.828])
        anchors[stride] = a
        A = a.shape[0]
        # scores straddling the 0.01 threshold so ~half pass, like random-init heads
        conf.append(distinct_scores(rng, (B, A * C, hw, hw), 0.001, 0.02))
        loc.append(rng.normal(0, 0.5, (B, A * 4, hw, hw)).astype(np.float32))
        G[f"dcr_conf{li}"], G[f"dcr_loc{li}"], G[f"dcr_anchors{li}"] = conf[-1], loc[-1], a.numpy()
    G["dcr_strides"] = np.asarray([s for s, _ in levels], np.int64)
    dec = Decoder(0.01, 0.6, 100, 300, True, True)
    s, b, c = dec([t(x) for x in loc], [t(x) for x in conf], anchors)
    G["dcr_scores"], G["dcr_boxes"], G["dcr_classes"] = s.numpy(), b.numpy(), c.numpy()


def rand_targets(rng, B, T, ncls, img):
    tg = np.full((B, T, 5), -1, np.float32)
    for b in range(B):
        n = int(rng.integers(1, T + 1))
        xy = rng.uniform(0, img * 0.7, (n, 2))
        wh = rng.uniform(img * 0.05, img * 0.5, (n, 2))
        tg[b, :n, :2], tg[b, :n, 2:4] = xy, wh
        tg[b, :n, 4] = rng.integers(0, ncls, n)
    return tg


def gen_match(rng):
    cases = [
        # B, T, ncls, stride, H, W, ratios, scales, radius
        (3, 6, 5, 8, 8, 8, [1, 2, 0.5], [4.0], 0),
        (2, 8, 80, 16, 10, 12, [1, 2, 0.5], [2.0, 2.828], 0),     # non-square map
        (2, 4, 3, 32, 5, 5, [1, 2, 0.5], [4.0, 5.04, 6.35], 0),
        (2, 6, 5, 8, 8, 8, [1, 2, 0.5], [4.0], 1.5),                # ATSS centre sampling on
    ]
    for i, (B, T, ncls, stride, H, W, ratios, scales, radius) in enumerate(cases):
        a = rbox.generate_anchors(stride, ratios, scales)
        anchors = OrderedDict([(stride, a)])
        tg = rand_targets(rng, B, T, ncls, min(H, W) * stride)
        for b in range(B):   # make target 0 a jittered grid anchor so every case has positives
            k = int(rng.integers(0, a.shape[0]))
            gx, gy = int(rng.integers(0, W)) * stride, int(rng.integers(0, H)) * stride
            x1, y1, x2, y2 = (a[k].numpy() + np.asarray([gx, gy, gx, gy])).tolist()
            j = rng.uniform(-0.04, 0.04, 4) * (x2 - x1)
            tg[b, 0, :4] = [x1 + j[0], y1 + j[1], x2 - x1 + 1 + j[2], y2 - y1 + 1 + j[3]]
        if i == 0:
            tg[2] = -1                                              # image with no targets
        cls_t, box_t, dep = rbox.extract_targets(
            t(tg), anchors, ncls, stride, (H, W), [0.5, 0.4], radius)
        p = f"mat{i}_"
        G[p + "targets"], G[p + "anchors"] = tg, a.numpy()
        G[p + "params"] = np.asarray([ncls, stride, H, W, radius], np.float64)
        G[p + "cls"], G[p + "box"], G[p + "depth"] = cls_t.numpy(), box_t.numpy(), dep.numpy()
    G["mat_n"] = np.int64(len(cases))


def gen_loss(rng):
    crit = MultiBoxLoss(negpos_ratio=3)
    cases = [(3, 3, 5, 8, 8), (2, 6, 80, 5, 5), (2, 9, 4, 10, 10)]
    for i, (B, A, C, H, W) in enumerate(cases):
        logits = rng.normal(-2.0, 2.0, (B, A, C, H, W)).astype(np.float32)
        depth = np.zeros((B, A, 1, H, W), np.float32)
        u = rng.uniform(size=depth.shape)
        cls = rng.integers(0, C, depth.shape)
        depth[u < 0.04] = (cls[u < 0.04] + 1).astype(np.float32)     # positives
        depth[(u >= 0.04) & (u < 0.10)] = -1                          # ignore
        if i == 0:
            depth[2] = 0                                              # image with no positives
        target = np.zeros_like(logits)
        for b, a, y, x in zip(*np.nonzero(depth[:, :, 0] > 0)):
            target[b, a, int(depth[b, a, 0, y, x]) - 1, y, x] = 1
        outs = [crit(t(logits[b:b + 1]), t(target[b:b + 1]), t(depth[b:b + 1])).numpy()
                for b in range(B)]                                    # B=1 slices
        p = f"mbl{i}_"
        G[p + "logits"], G[p + "target"], G[p + "depth"] = logits, target, depth
        G[p + "out"] = np.concatenate(outs, 0)
    G["mbl_n"] = np.int64(len(cases))


def gen_loss2():
    """Focal / SmoothL1 / IoU-family (own RNG so the earlier arrays stay byte-identical)."""
    rng = np.random.default_rng(20260924)
    cases = [(2, 3, 5, 8, 8), (2, 6, 80, 5, 5), (1, 9, 4, 10, 10)]
    for i, (B, A, C, H, W) in enumerate(cases):
        logits = rng.normal(-2.0, 2.5, (B, A, C, H, W)).astype(np.float32)
        depth = np.zeros((B, A, 1, H, W), np.float32)
        u = rng.uniform(size=depth.shape)
        cls = rng.integers(0, C, depth.shape)
        depth[u < 0.06] = (cls[u < 0.06] + 1).astype(np.float32)
        depth[(u >= 0.06) & (u < 0.12)] = -1
        target = np.zeros_like(logits)
        for b, a, y, x in zip(*np.nonzero(depth[:, :, 0] > 0)):
            target[b, a, int(depth[b, a, 0, y, x]) - 1, y, x] = 1
        # deltas: mostly small (overlapping boxes), some far apart (disjoint), some exactly equal
        bt = rng.normal(0, 0.6, (B, A, 4, H, W)).astype(np.float32)
        bp = (bt + rng.normal(0, 0.35, bt.shape)).astype(np.float32)
        far = rng.uniform(size=(B, A, 1, H, W)) < 0.15
        bp[:, :, :2] += np.where(far, 6.0, 0.0).astype(np.float32)
        same = rng.uniform(size=(B, A, 1, H, W)) < 0.05
        bp = np.where(same, bt, bp).astype(np.float32)
        p = f"ls{i}_"
        G[p + "logits"], G[p + "target"], G[p + "depth"] = logits, target, depth
        G[p + "box_pred"], G[p + "box_target"] = bp, bt
        G[p + "focal"] = FocalLoss(alpha=0.25, gamma=2)(t(logits), t(target), t(depth)).numpy()
        G[p + "smoothl1"] = SmoothL1Loss(beta=0.11)(t(bp), t(bt)).numpy()
        for ty in ("iou", "giou", "diou", "ciou"):
            G[p + ty] = IOULoss(loss_type=ty)(t(bp), t(bt)).numpy()
    G["ls_n"] = np.int64(len(cases))


def gen_loss_grads():
    """Gradients autograd gives on the reference modules for the masked, normalised sums of
    pipeline_anchor_basic.py:76-97 (inputs: the ls*/mbl* arrays generated above)."""
    for i in range(int(G["ls_n"])):
        p = f"ls{i}_"
        depth = t(G[p + "depth"])
        fg = (depth > 0).sum().float().clamp(min=1)
        G[p + "fg"] = np.float32(fg.item())
        x = t(G[p + "logits"]).requires_grad_(True)
        loss = FocalLoss(alpha=0.25, gamma=2)(x, t(G[p + "target"]), depth)
        ((loss * (depth >= 0).expand_as(loss).float()).sum() / fg).backward()
        G[p + "g_focal"] = x.grad.numpy().copy()
        for ty in ("smoothl1", "iou", "giou", "diou", "ciou"):
            bp = t(G[p + "box_pred"]).requires_grad_(True)
            crit = SmoothL1Loss(beta=0.11) if ty == "smoothl1" else IOULoss(loss_type=ty)
            l = crit(bp, t(G[p + "box_target"]))
            ((l * (depth > 0).expand_as(l).float()).sum() / fg).backward()
            G[p + "g_" + ty] = bp.grad.numpy().copy()
    crit = MultiBoxLoss(negpos_ratio=3)
    for i in range(int(G["mbl_n"])):
        p = f"mbl{i}_"
        grads = []
        for b in range(G[p + "logits"].shape[0]):                    # B=1 slices
            x = t(G[p + "logits"][b:b + 1]).requires_grad_(True)
            d = t(G[p + "depth"][b:b + 1])
            out = crit(x, t(G[p + "target"][b:b + 1]), d)
            (out * (d >= 0).expand_as(out).float()).sum().backward()
            grads.append(x.grad.numpy().copy())
        G[p + "grad"] = np.concatenate(grads, 0)


def main():
    rng = np.random.default_rng(20260923)
    gen_anchors()
    gen_codec(rng)
    gen_decode(rng)
    gen_nms(rng)
    gen_decoder(rng)
    gen_match(rng)
    gen_loss(rng)
    gen_loss2()
    gen_loss_grads()
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB,", len(G), "arrays")


if __name__ == "__main__":
    main()
