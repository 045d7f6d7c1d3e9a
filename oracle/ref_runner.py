"""Drives the UNMODIFIED reference (`oracle/_ref/ssds`, installed by oracle/build_ref.py) through its own
public API on the host CPU — the `--impl reference` arm / `cpu_baseline` leg of bench.py and the
reference side of the drop-in tests.

TEST / BENCH INFRASTRUCTURE ONLY: imports nothing from `ssds_pytorch_b200` (so the reference arm never
loads libssdsb200.so) and is imported by nothing in the product.

What runs is reference code only:
    create_model / create_anchors / create_decoder   ssds/modeling/model_builder.py:9-74
    model(x)                                           ssds/modeling/ssds/{ssd,fpn,bifpn,yolo}.py forward
    Decoder.__call__ -> decode -> nms                  ssds/modeling/layers/decoder.py:25-49, box.py:408-546
    extract_targets + MultiBoxLoss                     box.py:362-405, ssds/core/criterion.py:43-71
The only non-reference code is the torchvision compatibility shim of SURVEY 8c (`shim()`): the reference's
backbones touch torchvision names that torchvision >= 0.13 removed (model_urls, mobilenet.ConvBNReLU, ...),
and RegNet.initialize would download weights; the shim restores those names / skips the download.  No
reference arithmetic is touched.
"""
import os
import sys
import time
import warnings
from collections import defaultdict
from functools import partial

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# the yml shipped with the reference (experiments/cfgs/tests/test.yml:2-8), as a MODEL dict
YOLOV3_TEST_YML = dict(SSDS="YOLOV3", NETS="ResNet18", IMAGE_SIZE=[320, 320], NUM_CLASSES=80,
                       FEATURE_LAYER=[[3, 4, 5], [128, 256, 512]],
                       SIZES=[[2.0, 2.828], [2.0, 2.828], [2.0, 4.0, 8.0]], ASPECT_RATIOS=[[1, 2, 0.5]] * 3)


def available():
    return os.path.isfile(os.path.join(REF_DIR, "ssds", "modeling", "model_builder.py"))


_ref = None


def shim():
    """Import the reference with the torchvision-compat shim.  Returns a namespace of the
    reference modules used here."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise ImportError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    warnings.filterwarnings("ignore")
    import torch
    import torchvision as tv
    tv.models.resnet.model_urls = defaultdict(lambda: None)        # url=None -> initialize() skips the download
    tv.models.densenet.model_urls = defaultdict(lambda: None)
    from torchvision.models import mobilenetv2 as _mv2
    tv.models.mobilenet.model_urls = defaultdict(lambda: None)
    tv.models.mobilenet._make_divisible = _mv2._make_divisible
    tv.models.mobilenet.InvertedResidual = _mv2.InvertedResidual
    tv.models.mobilenet.ConvBNReLU = partial(tv.ops.misc.Conv2dNormActivation, norm_layer=torch.nn.BatchNorm2d,
                                             activation_layer=torch.nn.ReLU6)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from ssds.modeling.nets import regnet as _regnet
    _regnet.RegNet.initialize = lambda self: None                   # hard-coded URLs: no network
    from ssds.core import config as rcfg
    from ssds.core import criterion
    from ssds.modeling import model_builder
    from ssds.modeling.layers import box

    class NS:
        pass
    _ref = NS()
    _ref.cfg, _ref.model_builder, _ref.box, _ref.criterion = rcfg.cfg, model_builder, box, criterion
    return _ref


def build_reference_model(model_cfg, seed=0, per_level=300):
    """(model, anchors, decoder, cfg.MODEL) through the reference's own builders; `model_cfg` is a dict of the
    reference's MODEL yml keys (SSDS, NETS, IMAGE_SIZE, NUM_CLASSES, FEATURE_LAYER, SIZES, ASPECT_RATIOS);
    random init at `seed` (the reference's own initialisers)."""
    import torch
    R = shim()
    m = R.cfg.MODEL
    for k, v in model_cfg.items():
        setattr(m, k, [list(x) if isinstance(x, (list, tuple)) else x for x in v] if isinstance(v, list) else v)
    torch.manual_seed(seed)
    model = R.model_builder.create_model(m).eval()
    anchors = R.model_builder.create_anchors(m, model, m.IMAGE_SIZE)
    pp = R.cfg.POST_PROCESS
    pp.MAX_DETECTIONS_PER_LEVEL = per_level
    decoder = R.model_builder.create_decoder(pp)
    return model, anchors, decoder, m


def synthetic_targets(B, T=32, seed=4321):
    """SURVEY 8d cfg-4 targets: per image T~U{1..32} boxes, xy~U(0,480), wh~U(16,256), label~U{0..79}, -1 padded."""
    import torch
    g = torch.Generator().manual_seed(seed)
    tg = torch.full((B, T, 5), -1.0)
    for b in range(B):
        n = int(torch.randint(1, T + 1, (1,), generator=g))
        tg[b, :n, :2] = torch.rand((n, 2), generator=g) * 480
        tg[b, :n, 2:4] = torch.rand((n, 2), generator=g) * 240 + 16
        tg[b, :n, 4] = torch.randint(0, 80, (n,), generator=g).float()
    return tg


def reference_loss_step(R, model, anchors, images, targets, num_classes):
    """pipeline_anchor_basic.py:62-97 with CLASSIFY_LOSS=MultiBoxLoss (classification half; BASELINE configs[3]).
    MultiBoxLoss.forward only runs at B == 1 (criterion.py:66-68 expand_as bug), so the criterion
    is called per image, as the oracle does."""
    import torch
    crit = R.criterion.MultiBoxLoss(negpos_ratio=3)
    loc, conf = model(images)
    cls_losses, fg_targets = [], []
    for j, (stride, anchor) in enumerate(anchors.items()):
        size = conf[j].shape[-2:]
        conf_target, loc_target, depth = R.box.extract_targets(targets, anchors, num_classes, stride, size,
                                                               [0.5, 0.4], 0)
        fg_targets.append((depth > 0).sum().float().clamp(min=1))
        c = conf[j].view_as(conf_target).float()
        cls_mask = (depth >= 0).expand_as(conf_target).float()
        per_img = [crit(c[i:i + 1], conf_target[i:i + 1], depth[i:i + 1]) for i in range(c.shape[0])]
        cls_losses.append((cls_mask * torch.cat(per_img, 0)).sum())
    return torch.stack(cls_losses).sum() / torch.stack(fg_targets).sum()


_built = {}


def run(model_cfg, n_images, threads, steps=1, warm=True, seed=0, kind="detect", per_level=300):
    """Time the reference on `n_images` synthetic images of `model_cfg` on `threads` host threads.
    kind "detect": model.eval()(x) + Decoder;  kind "loss": training forward + extract_targets + MultiBoxLoss.
    Returns (seconds per step [list], description)."""
    import torch
    torch.set_num_threads(threads)
    R = shim()
    key = repr((sorted(model_cfg.items()), seed, per_level))
    if key not in _built:
        _built.clear()
        _built[key] = build_reference_model(model_cfg, seed, per_level)
    model, anchors, decoder, m = _built[key]
    model.eval()
    H, W = m.IMAGE_SIZE
    g = torch.Generator().manual_seed(1234)
    x = torch.randint(0, 256, (n_images, H, W, 3), generator=g, dtype=torch.uint8)
    x = (x.float() / 255.0).permute(0, 3, 1, 2).contiguous()           # ssds.py:48-57 normalisation
    if kind == "loss":
        model.train()                                                    # the reference's training forward
        tg = synthetic_targets(n_images)

        def once(xx, tt):
            with torch.no_grad():
                return reference_loss_step(R, model, anchors, xx, tt, m.NUM_CLASSES)
        if warm:
            once(x[:1], tg[:1])
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            once(x, tg)
            times.append(time.perf_counter() - t0)
        what = "reference model (train mode, no_grad) + extract_targets + MultiBoxLoss per image"
    else:
        def once(xx):
            with torch.no_grad():
                loc, conf = model(xx)
                return decoder(loc, conf, anchors)
        if warm:
            once(x[:1])
        times = []
        for _ in range(steps):
            t0 = time.perf_counter()
            once(x)
            times.append(time.perf_counter() - t0)
        what = "reference create_model(...).eval()(x) fp32 + reference Decoder (decode + nms)"
    return times, what


def detect(model_cfg, x, seed=0, state_dict=None, per_level=300):
    """Reference detections for an fp32 NCHW batch (drop-in tests): (scores, boxes, classes) numpy + anchors."""
    import torch
    model, anchors, decoder, m = build_reference_model(model_cfg, seed, per_level)
    if state_dict is not None:
        model.load_state_dict(state_dict)
        model.eval()
    with torch.no_grad():
        loc, conf = model(x)
        s, b, c = decoder(loc, conf, anchors)
    return (s.numpy(), b.numpy(), c.numpy()), anchors, (loc, conf), model
