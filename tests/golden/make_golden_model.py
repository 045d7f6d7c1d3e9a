"""Golden vectors for the conv stack, produced by the reference's OWN nn.Module (authoring container
only; needs /root/reference + the torchvision-compat shim of SURVEY 8c).

    python tests/golden/make_golden_model.py

Builds SSD+ResNet50 / SSD+ResNet18 through the reference's `create_model`, checks that
ssds_pytorch_b200.synth.ssd_resnet_shapes reproduces its state_dict keys/shapes, loads the synthetic
weights into it, runs `model.eval()(x)` on a small seeded image batch and stores the outputs
(fp16-compressed subsample + full-precision checksums) in tests/golden/model_small.npz, together with
the reference's `create_anchors` strides/anchors and the end-to-end `Decoder` detections.
"""
import os
import sys
import warnings
from collections import defaultdict
from functools import partial

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("SSDS_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
warnings.filterwarnings("ignore")

import torchvision as tv                                      # noqa: E402
tv.models.resnet.model_urls = defaultdict(lambda: None)       # -> url=None: initialize() skips download
tv.models.densenet.model_urls = defaultdict(lambda: None)
from torchvision.models import mobilenetv2 as _mv2            # noqa: E402
tv.models.mobilenet.model_urls = defaultdict(lambda: None)
tv.models.mobilenet._make_divisible = _mv2._make_divisible
tv.models.mobilenet.InvertedResidual = _mv2.InvertedResidual
tv.models.mobilenet.ConvBNReLU = partial(tv.ops.misc.Conv2dNormActivation, norm_layer=torch.nn.BatchNorm2d,
                                         activation_layer=torch.nn.ReLU6)

from ssds.modeling.nets import regnet as _regnet               # noqa: E402
_regnet.RegNet.initialize = lambda self: None                 # hard-coded URLs: no network here
from ssds.core import config as rcfg                          # noqa: E402
from ssds.modeling import model_builder                       # noqa: E402
from ssds_pytorch_b200 import synth                           # noqa: E402  (pure python part only)

CONF_CH_STRIDE = 7
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_small.npz")


def build(nets, feature_layer, sizes, ratios, image, num_classes, ssds="SSD"):
    cfg = rcfg.cfg.MODEL
    cfg.SSDS, cfg.NETS, cfg.IMAGE_SIZE, cfg.NUM_CLASSES = ssds, nets, image, num_classes
    cfg.FEATURE_LAYER, cfg.SIZES, cfg.ASPECT_RATIOS = feature_layer, sizes, ratios
    return cfg, model_builder.create_model(cfg)


def main():
    G = {}
    cases = {
        "r50": ("SSD", "ResNet50", [[3, 4, 5, "Conv:S", "Conv:S", "Conv:S"], [512, 1024, 2048, 512, 256, 256]],
                [256, 256], 80, 1),
        "r18": ("SSD", "ResNet18", [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]], [96, 160], 20, 3),
        "fpn50": ("SSDFPN", "ResNet50", [[3, 4, 5, "Conv:S", "Conv:S"], [512, 1024, 2048, 2048, 256]],
                  [256, 256], 20, 1),
        # BASELINE configs[0]/[2] model: SSD + MobileNetV2 at 300x300 (SURVEY 8a cfg 1b geometry)
        "mbv2": ("SSD", "MobileNetV2", [[5, 7, "Conv:S", "Conv:S", "Conv:S", "Conv:S"],
                                        [96, 320, 512, 256, 256, 128]], [300, 300], 20, 1),
        # BASELINE configs[4] model: SSDBiFPN + RegNetX-3.2GF (SURVEY 8a cfg 5 geometry), small image
        "bifpn": ("SSDBiFPN", "RegNetX032", [[2, 3, 4, "Conv:S", "Conv:S"], [192, 432, 1008, 1008, 256]],
                  [128, 128], 20, 1),
    }
    for tag, (ssds, nets, fl, image, ncls, B) in cases.items():
        L = len(fl[0])
        sizes = [[2.0, 2.828] for _ in range(L)]
        ratios = [[1, 2, 0.5] for _ in range(L)]
        cfg, model = build(nets, fl, [list(s) for s in sizes], ratios, image, ncls, ssds)
        ref_sd = model.state_dict()
        nb = [6] * L
        shapes = synth.model_shapes(ssds, nets, fl, nb, ncls)
        assert [k for k, _ in shapes] == list(ref_sd.keys()), "state_dict key order differs"
        for k, s in shapes:
            assert tuple(ref_sd[k].shape) == tuple(s), (k, ref_sd[k].shape, s)
        sd = synth.synthetic_state_dict(nets, fl, nb, ncls, seed=11, style="test", ssds=ssds)
        model.load_state_dict(sd)
        model.eval()
        anchors = model_builder.create_anchors(cfg, model, image)
        g = torch.Generator().manual_seed(1234)
        x = torch.rand((B, 3, image[0], image[1]), generator=g)
        with torch.no_grad():
            loc, conf = model(x)
        rcfg.cfg.POST_PROCESS.MAX_DETECTIONS_PER_LEVEL = 300
        decoder = model_builder.create_decoder(rcfg.cfg.POST_PROCESS)
        with torch.no_grad():
            det = decoder(loc, conf, anchors)
        G[tag + "_image"] = np.asarray(image)
        G[tag + "_x"] = x.numpy().astype(np.float16)          # tests regenerate x from the seed too
        G[tag + "_strides"] = np.asarray(list(anchors.keys()))
        for i, (s, a) in enumerate(anchors.items()):
            G[f"{tag}_anchors{i}"] = a.numpy()
        for i, (l, c) in enumerate(zip(loc, conf)):
            G[f"{tag}_loc{i}"] = l.numpy()
            G[f"{tag}_conf{i}"] = c.numpy()[:, ::CONF_CH_STRIDE]   # channel subsample keeps the file small
        G[tag + "_det_scores"], G[tag + "_det_boxes"], G[tag + "_det_classes"] = [d.numpy() for d in det]
        print(tag, "levels", [tuple(c.shape) for c in conf], "dets>0:", int((det[0] > 0).sum()))
    np.savez_compressed(OUT, **G)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


def main_yolo():
    """The model of the shipped experiments/cfgs/tests/test.yml (YOLOV3 + ResNet18, its own SIZES / ASPECT_RATIOS,
    80 classes) at a small image, through the reference's create_model -> model_yolo.npz (kept separate so that
    model_small.npz stays byte-identical)."""
    G = {}
    fl = [[3, 4, 5], [128, 256, 512]]
    sizes = [[2.0, 2.828], [2.0, 2.828], [2.0, 4.0, 8.0]]
    ratios = [[1, 2, 0.5]] * 3
    image, ncls, B, nb = [128, 160], 80, 2, [6, 6, 9]
    cfg, model = build("ResNet18", fl, [list(s_) for s_ in sizes], ratios, image, ncls, "YOLOV3")
    ref_sd = model.state_dict()
    shapes = synth.model_shapes("YOLOV3", "ResNet18", fl, nb, ncls)
    assert [k for k, _ in shapes] == list(ref_sd.keys()), "state_dict key order differs"
    for k, s_ in shapes:
        assert tuple(ref_sd[k].shape) == tuple(s_), (k, ref_sd[k].shape, s_)
    sd = synth.synthetic_state_dict("ResNet18", fl, nb, ncls, seed=11, style="test", ssds="YOLOV3")
    model.load_state_dict(sd)
    model.eval()
    anchors = model_builder.create_anchors(cfg, model, image)
    x = torch.rand((B, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        loc, conf = model(x)
    rcfg.cfg.POST_PROCESS.MAX_DETECTIONS_PER_LEVEL = 300
    decoder = model_builder.create_decoder(rcfg.cfg.POST_PROCESS)
    with torch.no_grad():
        det = decoder(loc, conf, anchors)
    tag = "yolo"
    G[tag + "_image"] = np.asarray(image)
    G[tag + "_x"] = x.numpy().astype(np.float16)
    G[tag + "_strides"] = np.asarray(list(anchors.keys()))
    for i, (s_, a) in enumerate(anchors.items()):
        G[f"{tag}_anchors{i}"] = a.numpy()
    for i, (l, c) in enumerate(zip(loc, conf)):
        G[f"{tag}_loc{i}"] = l.numpy()
        G[f"{tag}_conf{i}"] = c.numpy()[:, ::CONF_CH_STRIDE]
    G[tag + "_det_scores"], G[tag + "_det_boxes"], G[tag + "_det_classes"] = [d.numpy() for d in det]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_yolo.npz")
    np.savez_compressed(out, **G)
    print(tag, "levels", [tuple(c.shape) for c in conf], "dets>0:", int((det[0] > 0).sum()))
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


def main_yolo4():
    """YOLOV4 + ResNet18 (SPP + PAN neck, one 'Conv:S' extra level) through the reference's create_model ->
    model_yolo4.npz."""
    G = {}
    fl = [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]]
    sizes = [[2.0, 2.828]] * 4
    ratios = [[1, 2, 0.5]] * 4
    image, ncls, B, nb = [128, 160], 20, 2, [6] * 4
    cfg, model = build("ResNet18", fl, [list(s_) for s_ in sizes], ratios, image, ncls, "YOLOV4")
    ref_sd = model.state_dict()
    shapes = synth.model_shapes("YOLOV4", "ResNet18", fl, nb, ncls)
    assert [k for k, _ in shapes] == list(ref_sd.keys()), "state_dict key order differs"
    for k, s_ in shapes:
        assert tuple(ref_sd[k].shape) == tuple(s_), (k, ref_sd[k].shape, s_)
    sd = synth.synthetic_state_dict("ResNet18", fl, nb, ncls, seed=11, style="test", ssds="YOLOV4")
    model.load_state_dict(sd)
    model.eval()
    anchors = model_builder.create_anchors(cfg, model, image)
    x = torch.rand((B, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    with torch.no_grad():
        loc, conf = model(x)
    rcfg.cfg.POST_PROCESS.MAX_DETECTIONS_PER_LEVEL = 300
    decoder = model_builder.create_decoder(rcfg.cfg.POST_PROCESS)
    with torch.no_grad():
        det = decoder(loc, conf, anchors)
    tag = "yolo4"
    G[tag + "_image"] = np.asarray(image)
    G[tag + "_x"] = x.numpy().astype(np.float16)
    G[tag + "_strides"] = np.asarray(list(anchors.keys()))
    for i, (l, c) in enumerate(zip(loc, conf)):
        G[f"{tag}_loc{i}"] = l.numpy()
        G[f"{tag}_conf{i}"] = c.numpy()[:, ::CONF_CH_STRIDE]
    G[tag + "_det_scores"], G[tag + "_det_boxes"], G[tag + "_det_classes"] = [d.numpy() for d in det]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "model_yolo4.npz")
    np.savez_compressed(out, **G)
    print(tag, "levels", [tuple(c.shape) for c in conf], "dets>0:", int((det[0] > 0).sum()))
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    if "--yolo4" in sys.argv:
        main_yolo4()
    elif "--yolo" in sys.argv:
        main_yolo()
    else:
        main()
