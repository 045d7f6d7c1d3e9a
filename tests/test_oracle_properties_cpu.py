"""Size-independent properties of the hot path's functions, checked on the CPU oracle (the GPU suite checks
the same properties on the CUDA path at full BASELINE sizes): the oracle is pinned to the reference by the
golden vectors (test_oracle_golden.py); these tests make sure it also behaves like the algorithm it claims
to be on inputs the goldens do not cover."""
import numpy as np
import pytest

from oracle import box_oracle as O

f32 = np.float32


def _dets(rng, B, N, ncls, img=300.0):
    ctr = rng.uniform(20, img - 20, (B, N, 2))
    wh = rng.uniform(8, 80, (B, N, 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2], -1).astype(f32)
    scores = ((rng.permutation(B * N).reshape(B, N) + 0.5) / (B * N)).astype(f32)     # distinct
    scores[rng.uniform(size=scores.shape) < 0.1] = 0
    classes = rng.integers(0, ncls, (B, N)).astype(f32)
    return scores, boxes, classes


def _iou_plus1(a, b):
    x1, y1 = np.maximum(a[0], b[0]), np.maximum(a[1], b[1])
    x2, y2 = np.minimum(a[2], b[2]), np.minimum(a[3], b[3])
    inter = max(x2 - x1 + 1, 0) * max(y2 - y1 + 1, 0)
    ua = (a[2] - a[0] + 1) * (a[3] - a[1] + 1) + (b[2] - b[0] + 1) * (b[3] - b[1] + 1) - inter
    return inter / ua


@pytest.mark.parametrize("using_diou", [False, True])
def test_nms_properties(using_diou):
    rng = np.random.default_rng(5)
    s, b, c = _dets(rng, 3, 400, 4)
    os_, ob, oc, oi = O.nms(s, b, c, 0.5, 60, using_diou, return_indices=True)
    for i in range(3):
        n = int((oi[i] >= 0).sum())
        assert (np.diff(os_[i, :n]) <= 0).all() and (os_[i, n:] == 0).all()            # sorted, zero padded
        np.testing.assert_array_equal(os_[i, :n], s[i, oi[i, :n]])                      # outputs are inputs
        np.testing.assert_array_equal(ob[i, :n], b[i, oi[i, :n]])
        assert len(set(oi[i, :n].tolist())) == n
        if not using_diou:                                                              # plain IoU: kept boxes of
            for p in range(n):                                                          # one class do not overlap
                for q in range(p):
                    if oc[i, p] == oc[i, q]:
                        assert _iou_plus1(ob[i, q], ob[i, p]) <= 0.5 + 1e-6
    # idempotence: NMS of the NMS output keeps everything
    s2, b2, c2 = O.nms(os_, ob, oc, 0.5, 60, using_diou)
    np.testing.assert_array_equal(s2, os_)
    np.testing.assert_array_equal(b2, ob)
    # permuting the candidates does not change the result (scores are distinct)
    perm = rng.permutation(400)
    s3, b3, c3 = O.nms(s[:, perm], b[:, perm], c[:, perm], 0.5, 60, using_diou)
    np.testing.assert_array_equal(s3, os_)
    np.testing.assert_array_equal(b3, ob)


def test_decode_properties():
    rng = np.random.default_rng(6)
    B, A, C, H, W, stride, top_n = 2, 6, 5, 9, 7, 16, 40
    anc = O.generate_anchors(stride, [1, 2, 0.5], [2.0, 2.828])
    n = B * A * C * H * W
    conf = ((rng.permutation(n) + 0.5) / n).astype(f32).reshape(B, A * C, H, W)
    loc = rng.normal(0, 0.3, (B, A * 4, H, W)).astype(f32)
    sc, bx, cl, idx = O.decode(conf, loc, stride, 0.3, top_n, anc, rescore=False, return_indices=True)
    for b in range(B):
        flat = conf[b].reshape(-1)
        assert (np.diff(sc[b]) <= 0).all()                                   # descending
        np.testing.assert_array_equal(sc[b], flat[idx[b]])                   # rescore off: raw scores
        kth = sc[b, -1]
        assert (flat > kth).sum() == top_n - 1                               # exactly the top_n largest
        np.testing.assert_array_equal(cl[b], ((idx[b] // (H * W)) % C).astype(f32))     # box.py:448
        assert (bx[b, :, 0] >= 0).all() and (bx[b, :, 2] <= W * stride - 1).all()       # clamped (box.py:83)
        assert (bx[b, :, 1] >= 0).all() and (bx[b, :, 3] <= H * stride - 1).all()
    # a threshold above every score -> all-zero outputs (box.py:441-442)
    z = O.decode(conf, loc, stride, 2.0, top_n, anc)
    assert not z[0].any() and not z[1].any() and not z[2].any()
    # rescoring never raises a score (centerness factor in [0, 1], box.py:464-471)
    sr = O.decode(conf, loc, stride, 0.3, top_n, anc, rescore=True)[0]
    assert (sr <= sc + 1e-7).all()


def test_codec_round_trip_and_anchor_grid():
    rng = np.random.default_rng(7)
    anc = O.anchor_grid(O.generate_anchors(8, [1, 2, 0.5], [4.0, 5.04]), 8, 5, 4)
    anc = anc.reshape(-1, 4)
    ctr = (anc[:, :2] + anc[:, 2:]) / 2 + rng.normal(0, 3, (len(anc), 2))
    wh = (anc[:, 2:] - anc[:, :2] + 1) * rng.uniform(0.6, 1.6, (len(anc), 2))
    boxes = np.concatenate([ctr - wh / 2, ctr + wh / 2 - 1], 1).astype(f32)
    d = O.box2delta(boxes, anc)
    assert d.shape == (len(anc), 4) and np.isfinite(d).all()
    # delta2box inverts box2delta up to the clamp (boxes here are kept inside a large canvas)
    back = O.delta2box(d, anc, [10000, 10000], 1)
    np.testing.assert_allclose(back, np.maximum(boxes, 0), atol=2e-3)
    # zero deltas decode to the anchors themselves
    same = O.delta2box(np.zeros_like(d), anc, [10000, 10000], 1)
    np.testing.assert_allclose(same, np.maximum(anc, 0), atol=1e-4)


def test_multibox_loss_selection_counts():
    rng = np.random.default_rng(8)
    B, A, C, H, W = 3, 4, 6, 7, 5
    logits = rng.normal(-2, 2, (B, A, C, H, W)).astype(f32)
    depth = np.zeros((B, A, 1, H, W), f32)
    u = rng.uniform(size=depth.shape)
    depth[u < 0.05] = rng.integers(1, C + 1, depth.shape)[u < 0.05]
    depth[(u >= 0.05) & (u < 0.1)] = -1
    depth[1] = np.where(depth[1] > 0, 0, depth[1])                            # image without positives
    target = np.zeros_like(logits)
    for b, a, y, x in zip(*np.nonzero(depth[:, :, 0] > 0)):
        target[b, a, int(depth[b, a, 0, y, x]) - 1, y, x] = 1
    out = O.multibox_loss(logits, target, depth, 3)
    N = A * H * W
    for b in range(B):
        sel = (out[b] != 0).any(axis=1).reshape(-1)
        pos = (depth[b].reshape(-1) > 0)
        npos = int(pos.sum())
        assert (sel[pos]).all()                                               # every positive contributes
        assert (sel & ~pos).sum() == min(3 * npos, N - 1)                     # criterion.py:64-65
        assert not sel[depth[b].reshape(-1) < 0].any() or npos > 0           # ignored anchors are never mined...
        neg_sel = sel & ~pos
        assert not (neg_sel & (depth[b].reshape(-1) != 0)).any()             # ...nor anything with depth != 0
    assert not (out[1] != 0).any()


def test_spp_cascade_identity():
    """The SPP block of YOLOv4 (reference yolo.py:161-184) pools with kernels 5, 9 and 13 (stride 1, -inf padding);
    the engine runs three cascaded 5x5 pools instead.  max-pooling composes exactly: 5 o 5 = 9 and 5 o 9 = 13 —
    checked here on ragged map sizes, including maps smaller than the windows."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    for (h, w) in [(1, 1), (2, 3), (4, 5), (7, 9), (10, 12), (20, 20)]:
        x = torch.randn((2, 8, h, w), generator=g)
        p5 = F.max_pool2d(x, 5, 1, 2)
        p9 = F.max_pool2d(p5, 5, 1, 2)
        p13 = F.max_pool2d(p9, 5, 1, 2)
        assert torch.equal(p9, F.max_pool2d(x, 9, 1, 4))
        assert torch.equal(p13, F.max_pool2d(x, 13, 1, 6))
