"""Parity of the CUDA box-op kernels (through the C ABI) against
  (1) golden vectors produced by the reference itself (tests/golden/box_ops.npz), and
  (2) the numpy oracle on fresh seeded inputs, plus size-independent properties at the
      BASELINE.json configuration sizes.
Bar: bit-exact for anchors, indices/classes, NMS outputs, depth and class targets; boxes 1e-4
absolute / 1e-5 relative; rescored scores 2e-6 absolute / 1e-5 relative (the centerness factor
sqrt(min/max * min/max) amplifies a 1-ulp exp difference when a box edge is close to the anchor
centre; the north-star tolerance is 1e-4); deltas and BCE 1e-5 relative.
"""
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import ssds_pytorch_b200 as b
    return b


def cpu(t):
    return t.detach().cpu().numpy()


def distinct(rng, shape, lo=0.0, hi=1.0):
    n = int(np.prod(shape))
    v = (rng.permutation(n) + rng.uniform(0.25, 0.75, size=n)) / n
    return (lo + (hi - lo) * v).astype(np.float32).reshape(shape)


# ----------------------------------------------------------------------------- anchors / codec
def test_generate_anchors_golden(B, golden):
    for i in range(int(golden["anc_n"])):
        out = B.generate_anchors(int(golden[f"anc{i}_stride"]), list(golden[f"anc{i}_ratios"]),
                                 list(golden[f"anc{i}_scales"]))
        np.testing.assert_array_equal(cpu(out), golden[f"anc{i}_out"])


def test_anchor_grid_vs_oracle(B):
    from oracle import box_oracle as O
    for stride, W, H in [(8, 64, 64), (15, 19, 19), (32, 7, 5), (300, 1, 1)]:
        base = O.generate_anchors(stride, [1, 2, 0.5], [2.0, 2.828])
        g = B.anchor_grid(torch.from_numpy(base), stride, W, H)
        np.testing.assert_array_equal(cpu(g), O.anchor_grid(base, stride, W, H))


def test_codec_golden(B, golden):
    d = B.box2delta(torch.from_numpy(golden["codec_boxes"]), torch.from_numpy(golden["codec_anchors"]))
    np.testing.assert_allclose(cpu(d), golden["codec_box2delta"], rtol=1e-5, atol=1e-6)
    b = B.delta2box(torch.from_numpy(golden["codec_deltas"]), torch.from_numpy(golden["codec_anchors"]),
                    [40, 30], 16)
    np.testing.assert_allclose(cpu(b), golden["codec_delta2box"], rtol=1e-5, atol=1e-4)


# ----------------------------------------------------------------------------- decode
def test_decode_golden(B, golden):
    for i in range(int(golden["dec_n"])):
        p = f"dec{i}_"
        stride, thr, top_n, rescore = golden[p + "params"]
        s, b, c = B.decode(torch.from_numpy(golden[p + "conf"]), torch.from_numpy(golden[p + "loc"]),
                           int(stride), float(thr), int(top_n), torch.from_numpy(golden[p + "anchors"]),
                           bool(rescore))
        np.testing.assert_array_equal(cpu(c), golden[p + "classes"], err_msg=p)
        np.testing.assert_allclose(cpu(s), golden[p + "scores"], rtol=1e-5, atol=2e-6, err_msg=p)
        np.testing.assert_allclose(cpu(b), golden[p + "boxes"], rtol=1e-5, atol=1e-4, err_msg=p)


def test_decode_vs_oracle_multislice_and_ties(B):
    """A level larger than one 64Ki slice (exercises the cross-CTA bound) and an all-equal map
    (tie rule: ascending flat index)."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(7)
    A, C, H, W, stride = 6, 80, 32, 32, 16          # 491 520 scores / image -> 8 slices
    anc = O.generate_anchors(stride, [1, 2, 0.5], [2.0, 2.828])
    conf = distinct(rng, (2, A * C, H, W), 0.001, 0.02)
    loc = rng.normal(0, 0.5, (2, A * 4, H, W)).astype(np.float32)
    s, b, c, idx = B.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.01, 300,
                            torch.from_numpy(anc), True, return_indices=True)
    os_, ob, oc, oi = O.decode(conf, loc, stride, 0.01, 300, anc, True, return_indices=True)
    np.testing.assert_array_equal(cpu(idx).astype(np.int64), oi)
    np.testing.assert_array_equal(cpu(c), oc)
    np.testing.assert_allclose(cpu(s), os_, rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(cpu(b), ob, rtol=1e-5, atol=1e-4)
    # ties: constant map -> the first top_n flat indices, in order
    conf[:] = 0.5
    s, b, c, idx = B.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.01, 300,
                            torch.from_numpy(anc), True, return_indices=True)
    np.testing.assert_array_equal(cpu(idx), np.tile(np.arange(300, dtype=np.int32), (2, 1)))
    _, _, _, oi = O.decode(conf, loc, stride, 0.01, 300, anc, True, return_indices=True)
    np.testing.assert_array_equal(cpu(idx).astype(np.int64), oi)


def test_decode_large_top_n_rounds(B):
    """top_n > 1024 (the 20 000-per-level NMS stress of SURVEY 8d cfg 5) runs in rounds of 1024."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(21)
    A, C, H, W, stride = 3, 20, 24, 24, 8            # 34 560 scores / image
    anc = O.generate_anchors(stride, [1, 2, 0.5], [4.0])
    conf = distinct(rng, (2, A * C, H, W), 0.0, 1.0)
    conf[1, :, :, :] *= (rng.uniform(size=conf[1].shape) < 0.05)   # few candidates: exhausted early
    loc = rng.normal(0, 0.3, (2, A * 4, H, W)).astype(np.float32)
    for top_n in (2500, 5000):
        got = B.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.3, top_n,
                       torch.from_numpy(anc), True, return_indices=True)
        exp = O.decode(conf, loc, stride, 0.3, top_n, anc, True, return_indices=True)
        np.testing.assert_array_equal(cpu(got[3]).astype(np.int64), exp[3])
        np.testing.assert_array_equal(cpu(got[2]), exp[2])
        np.testing.assert_allclose(cpu(got[0]), exp[0], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(cpu(got[1]), exp[1], rtol=1e-5, atol=1e-4)


def test_decode_odd_sizes_and_edge_values(B):
    from oracle import box_oracle as O
    rng = np.random.default_rng(11)
    A, C, H, W, stride = 3, 7, 5, 3, 100            # 315 scores: scalar (unaligned) path
    anc = O.generate_anchors(stride, [1, 2, 0.5], [2.0])
    conf = distinct(rng, (3, A * C, H, W), -0.5, 1.0)   # negative scores must never pass
    conf[1] = 0.0                                        # an image where nothing passes
    conf[2, 0, 0, 0] = np.nan                            # NaN >= thr is False
    loc = rng.normal(0, 1.0, (3, A * 4, H, W)).astype(np.float32)
    for top_n in (1, 10, 400, 1024):
        got = B.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.05, top_n,
                       torch.from_numpy(anc), True, return_indices=True)
        exp = O.decode(np.nan_to_num(conf, nan=-1.0), loc, stride, 0.05, top_n, anc, True,
                       return_indices=True)
        np.testing.assert_array_equal(cpu(got[3]).astype(np.int64), exp[3])
        np.testing.assert_array_equal(cpu(got[2]), exp[2])
        np.testing.assert_allclose(cpu(got[0]), exp[0], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(cpu(got[1]), exp[1], rtol=1e-5, atol=1e-4)
    assert (cpu(got[0])[1] == 0).all()


# ----------------------------------------------------------------------------- nms
def test_nms_golden_bit_exact(B, golden):
    for i in range(int(golden["nms_n"])):
        p = f"nms{i}_"
        thr, D, diou = golden[p + "params"]
        s, b, c = B.nms(torch.from_numpy(golden[p + "scores"]), torch.from_numpy(golden[p + "boxes"]),
                        torch.from_numpy(golden[p + "classes"]), float(thr), int(D), bool(diou))
        np.testing.assert_array_equal(cpu(s), golden[p + "out_scores"], err_msg=p)
        np.testing.assert_array_equal(cpu(b), golden[p + "out_boxes"], err_msg=p)
        np.testing.assert_array_equal(cpu(c), golden[p + "out_classes"], err_msg=p)


def clustered(rng, B_, N, ncls, img=512.0, nclusters=8, frac_zero=0.05):
    scores = distinct(rng, (B_, N), 0.01, 1.0)
    scores[rng.uniform(size=(B_, N)) < frac_zero] = 0.0
    ctr = rng.uniform(30, img - 30, (B_, nclusters, 2))
    pick = rng.integers(0, nclusters, (B_, N))
    c = np.take_along_axis(ctr, pick[..., None].repeat(2, -1), 1) + rng.normal(0, 8, (B_, N, 2))
    wh = rng.uniform(8, 120, (B_, N, 2))
    boxes = np.clip(np.concatenate([c - wh / 2, c + wh / 2], -1), 0, img - 1).astype(np.float32)
    classes = rng.integers(0, ncls, (B_, N)).astype(np.float32)
    return scores, boxes, classes


@pytest.mark.parametrize("N,ncls,D,diou", [(1800, 3, 100, True), (1800, 80, 100, True),
                                           (900, 1, 100, False), (5000, 2, 100, True),
                                           (3000, 1, 300, True), (130, 2, 100, True)])
def test_nms_vs_oracle_indices_bit_exact(B, N, ncls, D, diou):
    """Heavy-suppression inputs: several chunks, several selection rounds (N > 2048)."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(N + ncls)
    s, b, c = clustered(rng, 4, N, ncls)
    got = B.nms(torch.from_numpy(s), torch.from_numpy(b), torch.from_numpy(c), 0.6, D, diou,
                return_indices=True)
    exp = O.nms(s, b, c, 0.6, D, diou, return_indices=True)
    np.testing.assert_array_equal(cpu(got[3]).astype(np.int64), exp[3])
    np.testing.assert_array_equal(cpu(got[0]), exp[0])
    np.testing.assert_array_equal(cpu(got[1]), exp[1])
    np.testing.assert_array_equal(cpu(got[2]), exp[2])


def test_nms_edge_cases(B):
    from oracle import box_oracle as O
    rng = np.random.default_rng(5)
    s, b, c = clustered(rng, 3, 64, 2)
    s[0] = 0.0                       # empty image -> zero row
    s[1, :60] = 0.0                  # only 4 valid
    s[2, 5] = np.nan                 # NaN score is dropped (score > 0 is False)
    for D in (1, 3, 100):
        got = B.nms(torch.from_numpy(s), torch.from_numpy(b), torch.from_numpy(c), 0.5, D, True,
                    return_indices=True)
        exp = O.nms(np.nan_to_num(s, nan=0.0), b, c, 0.5, D, True, return_indices=True)
        for g, e in zip(got[:3], exp[:3]):
            np.testing.assert_array_equal(cpu(g), e)
        np.testing.assert_array_equal(cpu(got[3]).astype(np.int64), exp[3])
    # ties: identical scores keep input order (stable sort), never suppress across classes,
    # DIoU term vanishes for boxes sharing a top-left corner (SURVEY 8c KATs)
    s = np.full((1, 6), 0.5, np.float32)
    b = np.tile(np.asarray([[10, 10, 50, 50]], np.float32), (1, 6, 1))
    c = np.asarray([[0, 1, 2, 0, 1, 2]], np.float32)
    got = B.nms(torch.from_numpy(s), torch.from_numpy(b), torch.from_numpy(c), 0.6, 100, True,
                return_indices=True)
    np.testing.assert_array_equal(cpu(got[3])[0, :4], [0, 1, 2, -1])
    # N == 0
    z = B.nms(torch.zeros(2, 0), torch.zeros(2, 0, 4), torch.zeros(2, 0), 0.5, 10, True)
    assert cpu(z[0]).shape == (2, 10) and (cpu(z[0]) == 0).all()


def test_nms_stress_properties_100k(B):
    """RegNet-BiFPN 1280 stress shape (SURVEY 8d cfg 5): N = 100 000, D = 100, 80 classes.
    Checked by properties + the oracle on the reduced problem that provably decides the answer."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(99)
    N = 100_000
    scores = distinct(rng, (2, N), 0.01, 1.0)
    xy = rng.uniform(0, 1180, (2, N, 2))
    wh = rng.uniform(8, 400, (2, N, 2))
    boxes = np.clip(np.concatenate([xy, xy + wh], -1), 0, 1279).astype(np.float32)
    classes = rng.integers(0, 80, (2, N)).astype(np.float32)
    got = B.nms(torch.from_numpy(scores), torch.from_numpy(boxes), torch.from_numpy(classes), 0.6,
                100, True, return_indices=True)
    gs, gb, gc, gi = [cpu(g) for g in got]
    assert (np.diff(gs, axis=1) <= 0).all() and (gs > 0).all()            # sorted, full rows
    for b_ in range(2):
        np.testing.assert_array_equal(scores[b_, gi[b_]], gs[b_])
        np.testing.assert_array_equal(boxes[b_, gi[b_]], gb[b_])
        # everything above the last kept score decides the result: run the oracle on that prefix
        cut = gs[b_, -1]
        sel = np.nonzero(scores[b_] >= cut)[0]
        es, eb, ec, ei = O.nms(scores[b_:b_ + 1, sel], boxes[b_:b_ + 1, sel], classes[b_:b_ + 1, sel],
                               0.6, 100, True, return_indices=True)
        np.testing.assert_array_equal(sel[ei[0]], gi[b_])


# ----------------------------------------------------------------------------- Decoder
def test_decoder_golden(B, golden):
    strides = [int(s) for s in golden["dcr_strides"]]
    anchors = OrderedDict((s, torch.from_numpy(golden[f"dcr_anchors{i}"])) for i, s in enumerate(strides))
    loc = [torch.from_numpy(golden[f"dcr_loc{i}"]) for i in range(len(strides))]
    conf = [torch.from_numpy(golden[f"dcr_conf{i}"]) for i in range(len(strides))]
    dec = B.Decoder(0.01, 0.6, 100, 300, True, True)
    s, b, c = dec(loc, conf, anchors)
    np.testing.assert_array_equal(cpu(c), golden["dcr_classes"])
    np.testing.assert_allclose(cpu(s), golden["dcr_scores"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(cpu(b), golden["dcr_boxes"], rtol=1e-5, atol=1e-4)


def test_legacy_C_signatures(B, golden):
    """ssds._C.decode / ssds._C.nms with the ODTK call shapes of box.py:419-421, :483-485."""
    p = "dec0_"
    stride, thr, top_n, rescore = golden[p + "params"]
    anchors = torch.from_numpy(golden[p + "anchors"])
    out = B._C.decode(torch.from_numpy(golden[p + "conf"]).cuda().float(),
                      torch.from_numpy(golden[p + "loc"]).cuda().float(),
                      anchors.view(-1).tolist(), int(stride), float(thr), int(top_n))
    np.testing.assert_array_equal(cpu(out[2]), golden[p + "classes"])
    p = "nms0_"
    thr, D, diou = golden[p + "params"]
    out = B._C.nms(torch.from_numpy(golden[p + "scores"]).cuda(), torch.from_numpy(golden[p + "boxes"]).cuda(),
                   torch.from_numpy(golden[p + "classes"]).cuda(), float(thr), int(D))
    np.testing.assert_array_equal(cpu(out[0]), golden[p + "out_scores"])


# ----------------------------------------------------------------------------- match
def test_extract_targets_golden(B, golden):
    for i in range(int(golden["mat_n"])):
        p = f"mat{i}_"
        ncls, stride, H, W, radius = golden[p + "params"]
        anchors = {int(stride): torch.from_numpy(golden[p + "anchors"])}
        cls_t, box_t, dep = B.extract_targets(torch.from_numpy(golden[p + "targets"]), anchors, int(ncls),
                                              int(stride), (int(H), int(W)), [0.5, 0.4], float(radius))
        np.testing.assert_array_equal(cpu(dep), golden[p + "depth"], err_msg=p)
        np.testing.assert_array_equal(cpu(cls_t), golden[p + "cls"], err_msg=p)
        np.testing.assert_allclose(cpu(box_t), golden[p + "box"], rtol=1e-5, atol=1e-6, err_msg=p)


def make_targets(rng, B_, T, ncls, img):
    tg = np.full((B_, T, 5), -1, np.float32)
    for b_ in range(B_):
        n = int(rng.integers(1, T + 1))
        tg[b_, :n, :2] = rng.uniform(0, img * 0.75, (n, 2))
        tg[b_, :n, 2:4] = rng.uniform(16, 256, (n, 2))
        tg[b_, :n, 4] = rng.integers(0, ncls, n)
    return tg


def test_extract_targets_cfg4_level_vs_oracle(B):
    """SSDFPN-ResNet50 640^2 geometry (SURVEY 8a cfg 4): stride 8 80x80 A=9, and stride 128 5x5,
    T = 32 padded rows, plus > 128 targets (second staging chunk)."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(4321)
    for stride, hw, T in [(8, 80, 32), (128, 5, 32), (16, 40, 150)]:
        anc = O.generate_anchors(stride, [1, 2, 0.5], [4.0, 5.04, 6.35])
        tg = make_targets(rng, 3, T, 80, 640)
        tg[1, 3:] = -1
        got = B.extract_targets(torch.from_numpy(tg), {stride: torch.from_numpy(anc)}, 80, stride,
                                (hw, hw), [0.5, 0.4])
        exp = O.extract_targets(tg, {stride: anc}, 80, stride, (hw, hw), [0.5, 0.4])
        np.testing.assert_array_equal(cpu(got[2]), exp[2])
        np.testing.assert_array_equal(cpu(got[0]), exp[0])
        np.testing.assert_allclose(cpu(got[1]), exp[1], rtol=1e-5, atol=1e-6)


# ----------------------------------------------------------------------------- MultiBoxLoss
def test_multibox_loss_golden(B, golden):
    crit = B.MultiBoxLoss(negpos_ratio=3)
    for i in range(int(golden["mbl_n"])):
        p = f"mbl{i}_"
        out = cpu(crit(torch.from_numpy(golden[p + "logits"]), torch.from_numpy(golden[p + "target"]),
                       torch.from_numpy(golden[p + "depth"])))
        ref = golden[p + "out"]
        np.testing.assert_array_equal(out != 0, ref != 0, err_msg=p)   # identical hard negatives
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6, err_msg=p)


def synth_loss_inputs(rng, B_, A, C, H, W):
    logits = rng.normal(-4.6, 1.0, (B_, A, C, H, W)).astype(np.float32)
    depth = np.zeros((B_, A, 1, H, W), np.float32)
    u = rng.uniform(size=depth.shape)
    cls = rng.integers(0, C, depth.shape)
    depth[u < 0.01] = (cls[u < 0.01] + 1).astype(np.float32)
    depth[(u >= 0.01) & (u < 0.03)] = -1
    target = np.zeros_like(logits)
    for b_, a, y, x in zip(*np.nonzero(depth[:, :, 0] > 0)):
        target[b_, a, int(depth[b_, a, 0, y, x]) - 1, y, x] = 1
    return logits, target, depth


def test_multibox_loss_cfg4_level_vs_oracle(B):
    """cfg 4 level geometry (A=9, C=80, 40x40) — selection checked tie-aware, sums to 1e-5."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(31)
    logits, target, depth = synth_loss_inputs(rng, 3, 9, 80, 40, 40)
    depth[2][depth[2] > 0] = 0                      # an image without positives -> no negatives
    crit = B.MultiBoxLoss(3)
    out = cpu(crit(torch.from_numpy(logits), torch.from_numpy(target), torch.from_numpy(depth)))
    ref = O.multibox_loss(logits, target, depth, 3)
    sel_g = (out != 0).any(axis=2)
    sel_r = (ref != 0).any(axis=2)
    # GPU expf/log1pf differ from numpy by an ulp, so the anchor AT the cut may swap with a
    # neighbour of (nearly) equal max_ce; everything else must agree exactly.
    assert (sel_g != sel_r).sum() <= 2 * out.shape[0]
    agree = (sel_g == sel_r)[:, :, None].repeat(out.shape[2], 2)
    np.testing.assert_allclose(out[agree], ref[agree], rtol=1e-5, atol=1e-6)
    assert sel_g.reshape(3, -1).sum(1)[2] == 0
    # fused sum variant == reduction of the drop-in output (pipeline_anchor_basic.py:76-82)
    ls, npos = crit.forward_sum(torch.from_numpy(logits), torch.from_numpy(depth))
    esum, enpos = O.multibox_loss_reduced(logits, target, depth, 3)
    np.testing.assert_array_equal(cpu(npos), enpos.astype(np.float32))
    np.testing.assert_allclose(cpu(ls), esum, rtol=2e-4)


def test_cls_loss_step_like_reference_pipeline(B):
    """pipeline_anchor_basic.py:62-97 (classification part): extract_targets per level + MultiBoxLoss
    + (depth >= 0) mask + normalisation by the foreground count, on an SSDFPN-like 3-level geometry."""
    from collections import OrderedDict
    from oracle import box_oracle as O
    from ssds_pytorch_b200.pipeline import multibox_cls_loss_step
    rng = np.random.default_rng(77)
    Bn, C = 3, 20
    levels = [(8, 20), (16, 10), (32, 5)]
    anchors = OrderedDict((s, O.generate_anchors(s, [1, 2, 0.5], [4.0, 5.04, 6.35])) for s, _ in levels)
    tg = make_targets(rng, Bn, 12, C, 160)
    logits = [rng.normal(-4.6, 1.0, (Bn, 9 * C, hw, hw)).astype(np.float32) for _, hw in levels]
    got, parts = multibox_cls_loss_step([torch.from_numpy(l) for l in logits], torch.from_numpy(tg),
                                        OrderedDict((s, torch.from_numpy(a)) for s, a in anchors.items()), C)
    tot, fg = 0.0, 0
    for (s, hw), lg in zip(levels, logits):
        cls_t, box_t, dep = O.extract_targets(tg, anchors, C, s, (hw, hw), [0.5, 0.4])
        sums, npos = O.multibox_loss_reduced(lg.reshape(Bn, 9, C, hw, hw), cls_t, dep, 3)
        tot += sums.sum()
        fg += max(int(npos.sum()), 1)
    assert fg > 3
    np.testing.assert_allclose(got.item(), tot / fg, rtol=3e-4)
    np.testing.assert_array_equal(cpu(parts[0][3]), O.extract_targets(tg, anchors, C, 8, (20, 20), [0.5, 0.4])[2])


# ----------------------------------------------------------------------------- Focal / SmoothL1 / IoU family
LOC_TYPES = ("smoothl1", "iou", "giou", "diou", "ciou")


def _loc_crit(B, ty):
    return B.SmoothL1Loss(0.11) if ty == "smoothl1" else B.IOULoss(ty)


def test_focal_smoothl1_iou_golden(B, golden):
    """drop-in unreduced outputs vs the reference's (tests/golden, criterion.py:95-239)."""
    for i in range(int(golden["ls_n"])):
        p = f"ls{i}_"
        out = cpu(B.FocalLoss(0.25, 2)(torch.from_numpy(golden[p + "logits"]), torch.from_numpy(golden[p + "target"]),
                                       torch.from_numpy(golden[p + "depth"])))
        np.testing.assert_allclose(out, golden[p + "focal"], rtol=2e-5, atol=1e-7, err_msg=p)
        for ty in LOC_TYPES:
            got = cpu(_loc_crit(B, ty)(torch.from_numpy(golden[p + "box_pred"]), torch.from_numpy(golden[p + "box_target"])))
            if ty == "smoothl1":
                np.testing.assert_allclose(got, golden[p + ty], rtol=1e-5, atol=2e-6, err_msg=p + ty)
            else:
                from oracle import box_oracle as O
                O.assert_iou_loss_close(got, golden[p + ty], golden[p + "box_pred"], golden[p + "box_target"],
                                        msg=p + ty)


def test_fused_loss_sums_vs_golden_reduction(B, golden):
    """forward_sum == the caller's mask + sum (pipeline_anchor_basic.py:76-97) of the golden outputs."""
    from oracle import box_oracle as O
    for i in range(int(golden["ls_n"])):
        p = f"ls{i}_"
        depth = golden[p + "depth"]
        cs, npos = B.FocalLoss().forward_sum(torch.from_numpy(golden[p + "logits"]), torch.from_numpy(depth))
        for ty in LOC_TYPES:
            ecs, els, enpos = O.masked_loss_sums(golden[p + "focal"], np.nan_to_num(golden[p + ty], nan=0.0), depth)
            ls = _loc_crit(B, ty).forward_sum(torch.from_numpy(golden[p + "box_pred"].reshape(depth.shape[0], -1, *depth.shape[-2:])),
                                              torch.from_numpy(golden[p + "box_target"]), torch.from_numpy(depth))
            got = cpu(ls)
            if ty == "ciou":       # identical boxes: NaN or ~1e-7 (see O.assert_iou_loss_close)
                same = (golden[p + "box_pred"] == golden[p + "box_target"]).all(axis=2, keepdims=True)
                poisoned = (same & (depth > 0)).reshape(depth.shape[0], -1).any(axis=1)
                assert (np.isnan(got) <= poisoned).all()
                got, els = got[~poisoned], els[~poisoned]
            np.testing.assert_allclose(got, els, rtol=2e-5, atol=1e-6, err_msg=p + ty)
        np.testing.assert_allclose(cpu(cs), ecs, rtol=2e-5, atol=1e-6, err_msg=p)
        np.testing.assert_array_equal(cpu(npos), enpos.astype(np.float32))


def test_detection_loss_step_like_reference_pipeline(B):
    """pipeline_anchor_basic.py:62-97 with the reference's default criteria (FocalLoss + SmoothL1Loss,
    config.py:151-152) on a 3-level geometry, vs the oracle run the long way."""
    from collections import OrderedDict
    from oracle import box_oracle as O
    from ssds_pytorch_b200.pipeline import detection_loss_step
    rng = np.random.default_rng(78)
    Bn, C = 3, 20
    levels = [(8, 20), (16, 10), (32, 5)]
    anchors = OrderedDict((s, O.generate_anchors(s, [1, 2, 0.5], [4.0, 5.04, 6.35])) for s, _ in levels)
    tg = make_targets(rng, Bn, 12, C, 160)
    conf = [rng.normal(-4.6, 1.0, (Bn, 9 * C, hw, hw)).astype(np.float32) for _, hw in levels]
    loc = [rng.normal(0, 0.5, (Bn, 9 * 4, hw, hw)).astype(np.float32) for _, hw in levels]
    tanc = OrderedDict((s, torch.from_numpy(a)) for s, a in anchors.items())
    for ty in ("smoothl1", "giou"):
        cl, ll, fg = detection_loss_step([torch.from_numpy(x) for x in loc], [torch.from_numpy(x) for x in conf],
                                         torch.from_numpy(tg), tanc, C, None, _loc_crit(B, ty))
        ecs = els = 0.0
        efg = 0
        for (s, hw), c, l in zip(levels, conf, loc):
            cls_t, box_t, dep = O.extract_targets(tg, anchors, C, s, (hw, hw), [0.5, 0.4])
            f = O.focal_loss(c.reshape(Bn, 9, C, hw, hw), cls_t)
            lv = O.loc_loss(l.reshape(Bn, 9, 4, hw, hw), box_t, ty)
            a, b_, n = O.masked_loss_sums(f, lv, dep)
            ecs += a.sum()
            els += b_.sum()
            efg += max(int(n.sum()), 1)
        assert efg > 3 and fg.item() == efg
        np.testing.assert_allclose(cl.item(), ecs / efg, rtol=1e-4)
        np.testing.assert_allclose(ll.item(), els / efg, rtol=1e-4)


def test_focal_loss_sum_cfg4_level(B):
    """cfg 4 level geometry (A=9, C=80, 40x40, B=8): fused sum vs oracle; gamma != 2 takes the powf path."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(32)
    logits, target, depth = synth_loss_inputs(rng, 8, 9, 80, 40, 40)
    for gamma in (2, 1.5):
        cs, npos = B.FocalLoss(0.25, gamma).forward_sum(torch.from_numpy(logits), torch.from_numpy(depth))
        ecs, _, enpos = O.masked_loss_sums(O.focal_loss(logits, target, 0.25, gamma), np.zeros_like(depth), depth)
        np.testing.assert_allclose(cpu(cs), ecs, rtol=1e-4)
        np.testing.assert_array_equal(cpu(npos), enpos.astype(np.float32))


# ----------------------------------------------------------------------------- loss backward (8f rank 1)
def test_loss_backward_vs_reference_autograd_golden(B, golden):
    """forward_sum(...) is differentiable: dL/dlogits, dL/dloc from the backward kernels vs the gradients
    autograd gives on the REFERENCE modules (tests/golden: ls*_g_*, mbl*_grad)."""
    for i in range(int(golden["ls_n"])):
        p = f"ls{i}_"
        depth = torch.from_numpy(golden[p + "depth"]).cuda()
        fg = float(golden[p + "fg"])
        x = torch.from_numpy(golden[p + "logits"]).cuda().requires_grad_(True)
        ls, npos = B.FocalLoss(0.25, 2).forward_sum(x, depth)
        assert not npos.requires_grad
        (ls.sum() / fg).backward()
        np.testing.assert_allclose(cpu(x.grad), golden[p + "g_focal"], rtol=2e-5, atol=1e-8, err_msg=p)
        same = (golden[p + "box_pred"] == golden[p + "box_target"]).all(axis=2, keepdims=True)
        for ty in LOC_TYPES:
            bp = torch.from_numpy(golden[p + "box_pred"]).cuda().requires_grad_(True)
            l = _loc_crit(B, ty).forward_sum(bp, torch.from_numpy(golden[p + "box_target"]), depth)
            (l.sum() / fg).backward()
            ref = golden[p + "g_" + ty]
            ok = ~np.broadcast_to(same, ref.shape) if ty == "ciou" else np.ones(ref.shape, bool)
            np.testing.assert_allclose(cpu(bp.grad)[ok], ref[ok], rtol=5e-5, atol=2e-8, err_msg=p + ty)
    for i in range(int(golden["mbl_n"])):
        p = f"mbl{i}_"
        x = torch.from_numpy(golden[p + "logits"]).cuda().requires_grad_(True)
        ls, _ = B.MultiBoxLoss(3).forward_sum(x, torch.from_numpy(golden[p + "depth"]))
        ls.sum().backward()
        np.testing.assert_allclose(cpu(x.grad), golden[p + "grad"], rtol=2e-5, atol=1e-7, err_msg=p)


def test_detection_loss_step_backward_vs_oracle(B):
    """the whole loss step (match + FocalLoss + SmoothL1 / GIoU, masks, 1/fg) differentiated w.r.t. the raw
    head outputs, vs the oracle (torch restatement + autograd, pinned by test_oracle_golden.py)."""
    from collections import OrderedDict
    from oracle import box_oracle as O
    from oracle import loss_grad_oracle as LG
    from ssds_pytorch_b200.pipeline import detection_loss_step
    rng = np.random.default_rng(79)
    Bn, C = 2, 20
    levels = [(8, 20), (16, 10)]
    anchors = OrderedDict((s, O.generate_anchors(s, [1, 2, 0.5], [4.0, 5.04, 6.35])) for s, _ in levels)
    tg = make_targets(rng, Bn, 10, C, 160)
    conf = [rng.normal(-3.0, 1.5, (Bn, 9 * C, hw, hw)).astype(np.float32) for _, hw in levels]
    loc = [rng.normal(0, 0.5, (Bn, 9 * 4, hw, hw)).astype(np.float32) for _, hw in levels]
    tanc = OrderedDict((s, torch.from_numpy(a)) for s, a in anchors.items())
    for ty in ("smoothl1", "giou"):
        tc = [torch.from_numpy(x).cuda().requires_grad_(True) for x in conf]
        tl = [torch.from_numpy(x).cuda().requires_grad_(True) for x in loc]
        cl, ll, fg = detection_loss_step(tl, tc, torch.from_numpy(tg), tanc, C, None, _loc_crit(B, ty))
        (cl + ll).backward()
        scale = np.full((Bn,), 1.0 / fg.item(), np.float32)
        for (s, hw), c, l, gc, gl in zip(levels, conf, loc, tc, tl):
            cls_t, box_t, dep = O.extract_targets(tg, anchors, C, s, (hw, hw), [0.5, 0.4])
            eg = LG.focal_sum_grad(c.reshape(Bn, 9, C, hw, hw), cls_t, dep, scale)
            np.testing.assert_allclose(cpu(gc.grad).reshape(eg.shape), eg, rtol=1e-4, atol=1e-8)
            el = LG.loc_sum_grad(l.reshape(Bn, 9, 4, hw, hw), box_t, dep, scale, ty)
            np.testing.assert_allclose(cpu(gl.grad).reshape(el.shape), el, rtol=1e-4, atol=2e-8)
            assert np.abs(el).max() > 0


# ----------------------------------------------------------------------------- decode, large top_n (decode_large.cu)
def _cmp_decode(B, O, conf, loc, stride, thr, top_n, anc):
    """indices / classes bit-exact; boxes 1e-4; raw scores (rescore off) bit-exact.  Rescored scores: the centerness
    ratio min(l,r)/max(l,r) amplifies the 1-ulp difference between the GPU's and numpy's exp() when a box edge sits
    next to the anchor centre, so over tens of thousands of results a few exceed the 2e-6 of the small tests: 2e-5
    absolute / 1e-3 relative here (observed worst case 8.3e-6 absolute, 2.2e-4 relative)."""
    tc, tl, ta = torch.from_numpy(conf), torch.from_numpy(loc), torch.from_numpy(anc)
    got = B.decode(tc, tl, stride, thr, top_n, ta, True, return_indices=True)
    exp = O.decode(conf, loc, stride, thr, top_n, anc, True, return_indices=True)
    np.testing.assert_array_equal(cpu(got[3]).astype(np.int64), exp[3])
    np.testing.assert_array_equal(cpu(got[2]), exp[2])
    np.testing.assert_allclose(cpu(got[0]), exp[0], rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(cpu(got[1]), exp[1], rtol=1e-5, atol=1e-4)
    raw = B.decode(tc, tl, stride, thr, top_n, ta, False, return_indices=True)
    rexp = O.decode(conf, loc, stride, thr, top_n, anc, False, return_indices=True)
    np.testing.assert_array_equal(cpu(raw[3]).astype(np.int64), rexp[3])
    np.testing.assert_array_equal(cpu(raw[0]), rexp[0])
    return got


def test_decode_large_top_n_stress_geometry(B):
    """SURVEY 8d cfg-5 stress: 20 000 candidates per level out of a 160x160x3x80 map whose scores are CLUSTERED like a
    random-init head's (sigmoid(-4.6 + small): nearly all within two exponents), several 64Ki slices; indices bit-exact."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(99)
    A, C, H, W, stride = 3, 80, 80, 80, 16           # 1 536 000 scores / image (24 slices)
    anc = O.generate_anchors(stride, [1, 2, 0.5], [4.0])
    logits = rng.normal(-4.6, 0.25, (2, A * C, H, W)).astype(np.float32)
    conf = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
    loc = rng.normal(0, 0.3, (2, A * 4, H, W)).astype(np.float32)
    for top_n in (20000, 4097, 40000):
        _cmp_decode(B, O, conf, loc, stride, 0.01, top_n, anc)
    # a random-init head on a real image: every score within 0.5 % of sigmoid(bias) = 0.01 (all 1.5 M keys share their
    # top ~17 score bits, thousands share all 32) — the selection must resolve down to the last bit and the index
    narrow = (1.0 / (1.0 + np.exp(-rng.normal(-4.595, 0.002, (2, A * C, H, W))))).astype(np.float32)
    _cmp_decode(B, O, narrow, loc, stride, 0.01, 20000, anc)


def test_decode_large_top_n_edge_cases(B, monkeypatch):
    """fewer survivors than top_n; ties at the cut (quantised scores -> big groups of equal keys, lowest flat index
    first); a constant map (the degenerate single-CTA re-scan path); large path == rounds path bit for bit."""
    from oracle import box_oracle as O
    rng = np.random.default_rng(5)
    A, C, H, W, stride = 3, 20, 48, 48, 8            # 138 240 scores / image
    anc = O.generate_anchors(stride, [1, 2, 0.5], [4.0])
    loc = rng.normal(0, 0.3, (2, A * 4, H, W)).astype(np.float32)
    conf = distinct(rng, (2, A * C, H, W), 0.0, 1.0)
    conf[1] *= (rng.uniform(size=conf[1].shape) < 0.01)          # ~1400 survivors < top_n
    _cmp_decode(B, O, conf, loc, stride, 0.3, 3000, anc)
    q = (np.floor(rng.uniform(0, 1, (2, A * C, H, W)) * 64) / 64).astype(np.float32)      # 64 distinct values
    _cmp_decode(B, O, q, loc, stride, 0.05, 5000, anc)
    const = np.full((2, A * C, H, W), 0.5, np.float32)          # every key equal: first top_n flat indices
    const[1, :, :, 7] = 0.75
    got = _cmp_decode(B, O, const, loc, stride, 0.01, 20000, anc)
    np.testing.assert_array_equal(cpu(got[3])[0], np.arange(20000, dtype=np.int32))
    # rounds path (SSDSB_DECODE_ROUNDS=1) == large path
    a = B.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.05, 6000, torch.from_numpy(anc), True,
                 return_indices=True)
    monkeypatch.setenv("SSDSB_DECODE_ROUNDS", "1")
    b_ = B.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.05, 6000, torch.from_numpy(anc), True,
                  return_indices=True)
    monkeypatch.delenv("SSDSB_DECODE_ROUNDS")
    for x, y in zip(a, b_):
        assert torch.equal(x, y)


def test_nms_packed_output_equals_separate_outputs(B):
    """The [B,D,6] block the NMS kernel can write directly (what SSDDetector ships / all-gathers) equals the three
    separate outputs, with and without them being requested."""
    rng = np.random.default_rng(3)
    scores, boxes, classes = clustered(rng, 3, 900, 5)
    s, b, c = B.nms(torch.from_numpy(scores), torch.from_numpy(boxes), torch.from_numpy(classes), 0.5, 100, True)
    pk = torch.full((3, 100, 6), float("nan"), device="cuda")
    s2, b2, c2 = B.nms(torch.from_numpy(scores), torch.from_numpy(boxes), torch.from_numpy(classes), 0.5, 100, True,
                       packed_out=pk)
    only = B.nms(torch.from_numpy(scores), torch.from_numpy(boxes), torch.from_numpy(classes), 0.5, 100, True,
                 packed_out=True)
    ref = torch.cat([s[..., None], b, c[..., None]], -1)
    assert torch.equal(s, s2) and torch.equal(b, b2) and torch.equal(c, c2)
    assert torch.equal(pk, ref) and torch.equal(only, ref)


def test_nms_long_rows_preselection_equals_streaming_selection(B, monkeypatch):
    """N > 8192: the first 2048 candidates come from the multi-CTA exact selection (decode_large.cu topk_rows); the
    outputs (scores, boxes, classes, keep indices) must equal the single-CTA streaming selection bit for bit — also
    when 2048 candidates are not enough and the kernel falls back to streaming rounds (one class, one tight cluster),
    with tied scores (ascending input position), and with fewer than 2048 positive scores."""
    rng = np.random.default_rng(17)
    cases = []
    s, b, c = clustered(rng, 2, 100000, 80, img=1280.0, nclusters=40)
    cases.append((s, b, c, 100))
    s, b, c = clustered(rng, 2, 20000, 1, img=200.0, nclusters=1)                # > 2048 candidates per pivot set
    cases.append((s, b, c, 300))
    s, b, c = clustered(rng, 2, 30000, 3, img=900.0, nclusters=12)
    s = (np.floor(s * 50) / 50).astype(np.float32)                              # heavy score ties
    cases.append((s, b, c, 100))
    s, b, c = clustered(rng, 2, 9000, 5, img=700.0, nclusters=9, frac_zero=0.9)  # ~900 positive scores
    cases.append((s, b, c, 100))
    for s, b, c, D in cases:
        ts, tb, tc = torch.from_numpy(s), torch.from_numpy(b), torch.from_numpy(c)
        got = B.nms(ts, tb, tc, 0.6, D, True, return_indices=True)
        monkeypatch.setenv("SSDSB_NMS_NO_PRESEL", "1")
        ref = B.nms(ts, tb, tc, 0.6, D, True, return_indices=True)
        monkeypatch.delenv("SSDSB_NMS_NO_PRESEL")
        for x, y in zip(got, ref):
            assert torch.equal(x, y)
        assert (got[0] > 0).sum() > 0
