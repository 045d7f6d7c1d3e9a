"""oracle/model_oracle.py (functional restatement of SSD.forward / ResNet.forward on a state_dict)
pinned to outputs of the reference's own nn.Module (tests/golden/model_small.npz), and the synthetic
state_dict key/shape scheme + anchor strides pinned to the reference's create_model/create_anchors."""
import os

import numpy as np
import pytest
import torch

from oracle import box_oracle as O
from oracle import model_oracle as M
from ssds_pytorch_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_small.npz")

CASES = {
    "r50": ("ResNet50", [[3, 4, 5, "Conv:S", "Conv:S", "Conv:S"], [512, 1024, 2048, 512, 256, 256]], 80, 1),
    "r18": ("ResNet18", [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]], 20, 3),
    "fpn50": ("ResNet50", [[3, 4, 5, "Conv:S", "Conv:S"], [512, 1024, 2048, 2048, 256]], 20, 1),
    "mbv2": ("MobileNetV2", [[5, 7, "Conv:S", "Conv:S", "Conv:S", "Conv:S"], [96, 320, 512, 256, 256, 128]], 20, 1),
    "bifpn": ("RegNetX032", [[2, 3, 4, "Conv:S", "Conv:S"], [192, 432, 1008, 1008, 256]], 20, 1),
}
SSDS = {"fpn50": "SSDFPN", "bifpn": "SSDBiFPN"}

FWD = {"r18": M.ssd_resnet_forward, "r50": M.ssd_resnet_forward, "fpn50": M.ssdfpn_resnet_forward,
       "mbv2": M.ssd_mobilenetv2_forward, "bifpn": M.ssdbifpn_forward}


def case_inputs(tag, gold):
    nets, fl, ncls, B = CASES[tag]
    L = len(fl[0])
    sd = synth.synthetic_state_dict(nets, fl, [6] * L, ncls, seed=11, style="test", ssds=SSDS.get(tag, "SSD"))
    image = [int(v) for v in gold[tag + "_image"]]
    x = torch.rand((B, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    return sd, fl, x, image, ncls


@pytest.mark.parametrize("tag", ["r18", "r50", "fpn50", "mbv2", "bifpn"])
def test_model_oracle_matches_reference_module(tag):
    gold = np.load(GOLD)
    sd, fl, x, image, ncls = case_inputs(tag, gold)
    np.testing.assert_array_equal(x.numpy().astype(np.float16), gold[tag + "_x"])
    torch.set_num_threads(8)
    with torch.no_grad():
        loc, conf = FWD[tag](sd, x, fl, training=False, policy="fp32")
    for i, (l, c) in enumerate(zip(loc, conf)):
        np.testing.assert_allclose(l.numpy(), gold[f"{tag}_loc{i}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(c.numpy()[:, ::7], gold[f"{tag}_conf{i}"], rtol=1e-4, atol=1e-6)
    # anchors / strides like model_builder.create_anchors (strides = W_in // W_feat)
    strides = [image[1] // c.shape[-1] for c in conf]
    np.testing.assert_array_equal(strides, gold[tag + "_strides"])
    for i, s in enumerate(strides):
        np.testing.assert_array_equal(O.generate_anchors(s, [1, 2, 0.5], [2.0, 2.828]),
                                      gold[f"{tag}_anchors{i}"])


def test_end_to_end_oracle_detections_r18():
    """model oracle + box oracle == reference model + reference Decoder, end to end."""
    from collections import OrderedDict
    gold = np.load(GOLD)
    tag = "r18"
    sd, fl, x, image, ncls = case_inputs(tag, gold)
    with torch.no_grad():
        loc, conf = M.ssd_resnet_forward(sd, x, fl, training=False, policy="fp32")
    loc, conf = [l.numpy() for l in loc], [c.numpy() for c in conf]
    anchors = OrderedDict((int(s), gold[f"{tag}_anchors{i}"]) for i, s in enumerate(gold[tag + "_strides"]))
    s, b, c = O.decoder_call(loc, conf, anchors, 0.01, 0.6, 100, 300, True, True)
    np.testing.assert_array_equal(c, gold[tag + "_det_classes"])
    np.testing.assert_allclose(s, gold[tag + "_det_scores"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(b, gold[tag + "_det_boxes"], rtol=1e-4, atol=1e-3)


YOLO_FL = [[3, 4, 5], [128, 256, 512]]
YOLO_SIZES = [[2.0, 2.828], [2.0, 2.828], [2.0, 4.0, 8.0]]


def test_yolov3_oracle_matches_reference_module():
    """oracle.yolov3_resnet_forward (yolo.py:44-87 restated) vs the reference's own YOLOV3 module outputs for the
    model of experiments/cfgs/tests/test.yml (tests/golden/model_yolo.npz, make_golden_model.py --yolo), and the
    box oracle's Decoder on them vs the reference's detections."""
    from collections import OrderedDict
    gold = np.load(os.path.join(os.path.dirname(GOLD), "model_yolo.npz"))
    sd = synth.synthetic_state_dict("ResNet18", YOLO_FL, [6, 6, 9], 80, seed=11, style="test", ssds="YOLOV3")
    image = [int(v) for v in gold["yolo_image"]]
    x = torch.rand((2, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    np.testing.assert_array_equal(x.numpy().astype(np.float16), gold["yolo_x"])
    with torch.no_grad():
        loc, conf = M.yolov3_resnet_forward(sd, x, YOLO_FL, training=False, policy="fp32")
    for i, (l, c) in enumerate(zip(loc, conf)):
        np.testing.assert_allclose(l.numpy(), gold[f"yolo_loc{i}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(c.numpy()[:, ::7], gold[f"yolo_conf{i}"], rtol=1e-4, atol=1e-6)
    strides = [image[1] // c.shape[-1] for c in conf]
    np.testing.assert_array_equal(strides, gold["yolo_strides"])
    anchors = OrderedDict()
    for i, s in enumerate(strides):
        a = O.generate_anchors(s, [1, 2, 0.5], YOLO_SIZES[i])
        np.testing.assert_array_equal(a, gold[f"yolo_anchors{i}"])
        anchors[s] = a
    s_, b_, c_ = O.decoder_call([l.numpy() for l in loc], [c.numpy() for c in conf], anchors, 0.01, 0.6, 100, 300,
                                True, True)
    np.testing.assert_array_equal(c_, gold["yolo_det_classes"])
    np.testing.assert_allclose(s_, gold["yolo_det_scores"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(b_, gold["yolo_det_boxes"], rtol=1e-4, atol=1e-3)


YOLO4_FL = [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]]


def test_yolov4_oracle_matches_reference_module():
    """[r2] oracle.yolov4_resnet_forward (yolo.py:161-323 restated: SPP block, PANModule, extras, heads) vs the
    reference's own YOLOV4 module outputs (tests/golden/model_yolo4.npz, make_golden_model.py --yolo4), and the box
    oracle's Decoder on them vs the reference's detections."""
    from collections import OrderedDict
    gold = np.load(os.path.join(os.path.dirname(GOLD), "model_yolo4.npz"))
    sd = synth.synthetic_state_dict("ResNet18", YOLO4_FL, [6] * 4, 20, seed=11, style="test", ssds="YOLOV4")
    image = [int(v) for v in gold["yolo4_image"]]
    x = torch.rand((2, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    np.testing.assert_array_equal(x.numpy().astype(np.float16), gold["yolo4_x"])
    with torch.no_grad():
        loc, conf = M.yolov4_resnet_forward(sd, x, YOLO4_FL, training=False, policy="fp32")
    assert len(loc) == 4
    for i, (l, c) in enumerate(zip(loc, conf)):
        np.testing.assert_allclose(l.numpy(), gold[f"yolo4_loc{i}"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(c.numpy()[:, ::7], gold[f"yolo4_conf{i}"], rtol=1e-4, atol=1e-6)
    strides = [image[1] // c.shape[-1] for c in conf]
    np.testing.assert_array_equal(strides, gold["yolo4_strides"])
    anchors = OrderedDict((s, O.generate_anchors(s, [1, 2, 0.5], [2.0, 2.828])) for s in strides)
    s_, b_, c_ = O.decoder_call([l.numpy() for l in loc], [c.numpy() for c in conf], anchors, 0.01, 0.6, 100, 300,
                                True, True)
    np.testing.assert_array_equal(c_, gold["yolo4_det_classes"])
    np.testing.assert_allclose(s_, gold["yolo4_det_scores"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(b_, gold["yolo4_det_boxes"], rtol=1e-4, atol=1e-3)
