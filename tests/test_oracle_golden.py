"""The numpy oracle (oracle/box_oracle.py) replayed against golden vectors that
were produced by running the reference's own functions (tests/golden/make_golden.py).

Integer/index/keep results must be identical; float results that only use
+,-,*,/,sqrt must be bit-identical; exp/log paths are held to 1e-5 relative.
"""
from collections import OrderedDict

import numpy as np

from oracle import box_oracle as O


def test_generate_anchors_exact(golden):
    for i in range(int(golden["anc_n"])):
        out = O.generate_anchors(int(golden[f"anc{i}_stride"]), list(golden[f"anc{i}_ratios"]),
                                 list(golden[f"anc{i}_scales"]))
        np.testing.assert_array_equal(out, golden[f"anc{i}_out"])


def test_survey_kats():
    # hand-checked micro-KATs from SURVEY.md section 8c
    a = O.generate_anchors(8, [1, 2, 0.5], [4.0, 5.04, 6.35])[:3]
    np.testing.assert_array_equal(a, [[-12, -12, 19, 19], [-8, -20, 15, 27], [-18, -8, 25, 15]])
    np.testing.assert_array_equal(O.generate_anchors(8, [1, 2], [2.0]),
                                  [[-4, -4, 11, 11], [-2, -8, 9, 15]])
    np.testing.assert_array_equal(
        O.generate_anchors(15, [1, 2, 0.5], [2.0]),
        [[-7.5, -7.5, 21.5, 21.5], [-3.5, -14.5, 17.5, 28.5], [-13.5, -2.5, 27.5, 16.5]])
    # decode of a single 0.7 at (a=1,c=2,y=1,x=3) on a 2x4 map, stride 8, zero deltas
    A, C, H, W = 2, 3, 2, 4
    conf = np.zeros((1, A * C, H, W), np.float32)
    conf[0, 1 * C + 2, 1, 3] = 0.7
    loc = np.zeros((1, A * 4, H, W), np.float32)
    anc = O.generate_anchors(8, [1, 2], [2.0])
    s, b, c = O.decode(conf, loc, 8, 0.05, 5, anc, True)
    np.testing.assert_array_equal(b[0, 0], [22, 0, 31, 15])
    assert c[0, 0] == 2 and abs(s[0, 0] - 0.3081) < 1e-4
    assert (s[0, 1:] == 0).all()


def test_codec(golden):
    np.testing.assert_allclose(O.box2delta(golden["codec_boxes"], golden["codec_anchors"]),
                               golden["codec_box2delta"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(
        O.delta2box(golden["codec_deltas"], golden["codec_anchors"], [40, 30], 16),
        golden["codec_delta2box"], rtol=1e-5, atol=1e-4)


def test_decode(golden):
    for i in range(int(golden["dec_n"])):
        p = f"dec{i}_"
        stride, thr, top_n, rescore = golden[p + "params"]
        s, b, c = O.decode(golden[p + "conf"], golden[p + "loc"], int(stride), float(thr),
                           int(top_n), golden[p + "anchors"], bool(rescore))
        np.testing.assert_array_equal(c, golden[p + "classes"])
        np.testing.assert_allclose(s, golden[p + "scores"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(b, golden[p + "boxes"], rtol=1e-5, atol=1e-4)


def test_nms_bit_exact(golden):
    for i in range(int(golden["nms_n"])):
        p = f"nms{i}_"
        thr, D, diou = golden[p + "params"]
        s, b, c = O.nms(golden[p + "scores"], golden[p + "boxes"], golden[p + "classes"],
                        float(thr), int(D), bool(diou))
        np.testing.assert_array_equal(s, golden[p + "out_scores"])
        np.testing.assert_array_equal(b, golden[p + "out_boxes"])
        np.testing.assert_array_equal(c, golden[p + "out_classes"])


def test_decoder(golden):
    strides = [int(s) for s in golden["dcr_strides"]]
    anchors = OrderedDict((s, golden[f"dcr_anchors{i}"]) for i, s in enumerate(strides))
    loc = [golden[f"dcr_loc{i}"] for i in range(len(strides))]
    conf = [golden[f"dcr_conf{i}"] for i in range(len(strides))]
    s, b, c = O.decoder_call(loc, conf, anchors, 0.01, 0.6, 100, 300, True, True)
    np.testing.assert_array_equal(c, golden["dcr_classes"])
    np.testing.assert_allclose(s, golden["dcr_scores"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(b, golden["dcr_boxes"], rtol=1e-5, atol=1e-4)
    assert (golden["dcr_scores"] > 0).sum() > 50


def test_extract_targets(golden):
    for i in range(int(golden["mat_n"])):
        p = f"mat{i}_"
        ncls, stride, H, W, radius = golden[p + "params"]
        anchors = {int(stride): golden[p + "anchors"]}
        cls_t, box_t, dep = O.extract_targets(golden[p + "targets"], anchors, int(ncls),
                                              int(stride), (int(H), int(W)), [0.5, 0.4],
                                              float(radius))
        np.testing.assert_array_equal(dep, golden[p + "depth"])
        np.testing.assert_array_equal(cls_t, golden[p + "cls"])
        np.testing.assert_allclose(box_t, golden[p + "box"], rtol=1e-5, atol=1e-6)
        assert (golden[p + "depth"] > 0).sum() > 0


def test_multibox_loss(golden):
    for i in range(int(golden["mbl_n"])):
        p = f"mbl{i}_"
        out = O.multibox_loss(golden[p + "logits"], golden[p + "target"], golden[p + "depth"])
        ref = golden[p + "out"]
        np.testing.assert_array_equal(out != 0, ref != 0)      # identical hard-negative selection
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_focal_smoothl1_iou_losses(golden):
    """oracle restatements of criterion.py:95-108 / :138-151 / :175-239 vs the reference's outputs
    (NaNs included: ciou of identical boxes is 0/0 in the reference)."""
    assert int(golden["ls_n"]) >= 3
    for i in range(int(golden["ls_n"])):
        p = f"ls{i}_"
        np.testing.assert_allclose(O.focal_loss(golden[p + "logits"], golden[p + "target"]),
                                   golden[p + "focal"], rtol=2e-5, atol=1e-7)
        np.testing.assert_allclose(O.smooth_l1_loss(golden[p + "box_pred"], golden[p + "box_target"]),
                                   golden[p + "smoothl1"], rtol=1e-6, atol=1e-7)
        for ty in ("iou", "giou", "diou", "ciou"):
            O.assert_iou_loss_close(O.iou_loss(golden[p + "box_pred"], golden[p + "box_target"], ty),
                                    golden[p + ty], golden[p + "box_pred"], golden[p + "box_target"], msg=p + ty)
    assert np.isnan(golden["ls0_ciou"]).sum() > 0


def test_loss_gradients(golden):
    """oracle backward (torch restatement + autograd) vs autograd on the reference modules."""
    from oracle import loss_grad_oracle as LG
    for i in range(int(golden["ls_n"])):
        p = f"ls{i}_"
        depth = golden[p + "depth"]
        B = depth.shape[0]
        scale = np.full((B,), 1.0 / float(golden[p + "fg"]), np.float32)
        np.testing.assert_allclose(LG.focal_sum_grad(golden[p + "logits"], golden[p + "target"], depth, scale),
                                   golden[p + "g_focal"], rtol=1e-5, atol=1e-9)
        same = (golden[p + "box_pred"] == golden[p + "box_target"]).all(axis=2, keepdims=True)
        for ty in ("smoothl1", "iou", "giou", "diou", "ciou"):
            g = LG.loc_sum_grad(golden[p + "box_pred"], golden[p + "box_target"], depth, scale, ty)
            ref = golden[p + "g_" + ty]
            ok = ~np.broadcast_to(same, ref.shape) if ty == "ciou" else np.ones(ref.shape, bool)
            np.testing.assert_allclose(g[ok], ref[ok], rtol=1e-5, atol=1e-8, err_msg=p + ty)
    for i in range(int(golden["mbl_n"])):
        p = f"mbl{i}_"
        B = golden[p + "depth"].shape[0]
        g = LG.multibox_sum_grad(golden[p + "logits"], golden[p + "target"], golden[p + "depth"],
                                 np.ones((B,), np.float32))
        np.testing.assert_allclose(g, golden[p + "grad"], rtol=1e-5, atol=1e-7, err_msg=p)
