"""Synthetic (seeded) weights with the reference's `state_dict()` key names and shapes for
SSD + ResNet models — there is no network for pretrained checkpoints, and the reference's own random
init cannot travel to the GPU box.  tests/golden/make_golden_model.py checks that this key/shape set
equals `create_model(cfg).state_dict()` of the reference.

Key scheme: torchvision ResNet under `backbone.` (resnet.py:9-35), `extras.{j}.{0,1,3,4}` for
ConvBNReLUx2 (basic_layers.py:41-57), `loc.{l}` / `conf.{l}` heads (ssd.py:100-103).
"""
import math

import torch

RESNETS = {
    "ResNet18": ("basic", [2, 2, 2, 2]), "ResNet34": ("basic", [3, 4, 6, 3]),
    "ResNet50": ("bottleneck", [3, 4, 6, 3]), "ResNet101": ("bottleneck", [3, 4, 23, 3]),
    "ResNet152": ("bottleneck", [3, 8, 36, 3]),
}


def _bn_keys(out, p, c):
    out.extend([(p + ".weight", (c,)), (p + ".bias", (c,)), (p + ".running_mean", (c,)),
                (p + ".running_var", (c,)), (p + ".num_batches_tracked", ())])


def resnet_backbone_shapes(nets):
    """(key, shape) list of the torchvision ResNet under `backbone.` (resnet.py:9-35)."""
    block, layers = RESNETS[nets]
    exp = 4 if block == "bottleneck" else 1
    out = []

    def bn(p, c):
        _bn_keys(out, p, c)

    out.append(("backbone.conv1.weight", (64, 3, 7, 7)))
    bn("backbone.bn1", 64)
    inplanes = 64
    for li, nblocks in enumerate(layers, start=1):
        planes = 64 * 2 ** (li - 1)
        for bi in range(nblocks):
            p = f"backbone.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            if block == "bottleneck":
                out.append((p + ".conv1.weight", (planes, inplanes, 1, 1))); bn(p + ".bn1", planes)
                out.append((p + ".conv2.weight", (planes, planes, 3, 3))); bn(p + ".bn2", planes)
                out.append((p + ".conv3.weight", (planes * 4, planes, 1, 1))); bn(p + ".bn3", planes * 4)
            else:
                out.append((p + ".conv1.weight", (planes, inplanes, 3, 3))); bn(p + ".bn1", planes)
                out.append((p + ".conv2.weight", (planes, planes, 3, 3))); bn(p + ".bn2", planes)
            if stride != 1 or inplanes != planes * exp:
                out.append((p + ".downsample.0.weight", (planes * exp, inplanes, 1, 1)))
                bn(p + ".downsample.1", planes * exp)
            inplanes = planes * exp
    out.append(("backbone.fc.weight", (1000, 512 * exp)))
    out.append(("backbone.fc.bias", (1000,)))
    return out


MBV2_SETTINGS = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2],
                 [6, 320, 1, 1]]


def mobilenetv2_backbone_shapes():
    """(key, shape) list of the reference MobileNetEx v2 under `backbone.` (mobilenet.py:40-100):
    conv1 = ConvBNReLU(3,32,s2); layer{j}.{i} = torchvision InvertedResidual; head_conv; classifier."""
    out = [("backbone.conv1.0.weight", (32, 3, 3, 3))]
    _bn_keys(out, "backbone.conv1.1", 32)
    inp = 32
    for j, (t, c, n, s) in enumerate(MBV2_SETTINGS, start=1):
        for i in range(n):
            p = f"backbone.layer{j}.{i}.conv"
            hid = inp * t
            k = 0
            if t != 1:
                out.append((f"{p}.0.0.weight", (hid, inp, 1, 1))); _bn_keys(out, f"{p}.0.1", hid)
                k = 1
            out.append((f"{p}.{k}.0.weight", (hid, 1, 3, 3))); _bn_keys(out, f"{p}.{k}.1", hid)
            out.append((f"{p}.{k + 1}.weight", (c, hid, 1, 1))); _bn_keys(out, f"{p}.{k + 2}", c)
            inp = c
    out.append(("backbone.head_conv.0.weight", (1280, 320, 1, 1)))
    _bn_keys(out, "backbone.head_conv.1", 1280)
    out.append(("backbone.classifier.1.weight", (1000, 1280)))
    out.append(("backbone.classifier.1.bias", (1000,)))
    return out


REGNETX032 = {"ws": [96, 192, 432, 1008], "ds": [2, 6, 15, 2], "gw": 48}


def regnetx032_backbone_shapes():
    """(key, shape) list of the reference RegNetX-3.2GF under `backbone.` (regnet.py:28-282, :388-401)."""
    out = [("backbone.stem.conv.weight", (32, 3, 3, 3))]
    _bn_keys(out, "backbone.stem.bn", 32)
    w_in = 32
    for si, (w, d) in enumerate(zip(REGNETX032["ws"], REGNETX032["ds"]), start=1):
        for bi in range(1, d + 1):
            p = f"backbone.s{si}.b{bi}"
            stride = 2 if bi == 1 else 1
            if w_in != w or stride != 1:
                out.append((p + ".proj.weight", (w, w_in, 1, 1))); _bn_keys(out, p + ".bn", w)
            out.append((p + ".f.a.weight", (w, w_in, 1, 1))); _bn_keys(out, p + ".f.a_bn", w)
            out.append((p + ".f.b.weight", (w, REGNETX032["gw"], 3, 3))); _bn_keys(out, p + ".f.b_bn", w)
            out.append((p + ".f.c.weight", (w, w, 1, 1))); _bn_keys(out, p + ".f.c_bn", w)
            w_in = w
    out.append(("backbone.head.fc.weight", (1000, w_in)))
    out.append(("backbone.head.fc.bias", (1000,)))
    return out


def fpn_neck_shapes(feature_layer, number_box, num_classes, bifpn_stacks=0):
    """transforms / extras / [stack_bifpn] / shared towers of SSDFPN (fpn.py:103-146) and SSDBiFPN
    (bifpn.py:144-191)."""
    out = []
    ti = 0
    for layer, depth in zip(feature_layer[0], feature_layer[1]):
        if isinstance(layer, int):
            out.append((f"transforms.{ti}.weight", (256, depth, 1, 1)))
            out.append((f"transforms.{ti}.bias", (256,)))
            ti += 1
    for i, (layer, depth) in enumerate(zip(feature_layer[0], feature_layer[1])):
        cin = 256 if isinstance(layer, int) else depth
        out.append((f"extras.{i}.0.weight", (256, cin, 3, 3)))
        _bn_keys(out, f"extras.{i}.1", 256)
    for s_ in range(bifpn_stacks):
        p = f"stack_bifpn.{s_}"
        out.append((p + ".w1", (2, ti)))
        out.append((p + ".w2", (3, ti - 2)))
        for i in range(ti - 1, 0, -1):
            out.append((f"{p}.top-down-{i - 1}.0.weight", (256, 256, 3, 3))); _bn_keys(out, f"{p}.top-down-{i - 1}.1", 256)
        for i in range(0, ti - 1):
            out.append((f"{p}.bottom-up-{i + 1}.0.weight", (256, 256, 3, 3))); _bn_keys(out, f"{p}.bottom-up-{i + 1}.1", 256)
    for tower, cout in (("loc", number_box[0] * 4), ("conf", number_box[0] * num_classes)):
        for j in range(4):
            out.append((f"{tower}.{j}.0.weight", (256, 256, 3, 3)))
            _bn_keys(out, f"{tower}.{j}.1", 256)
        out.append((f"{tower}.4.weight", (cout, 256, 3, 3)))
        out.append((f"{tower}.4.bias", (cout,)))
    return out


def model_shapes(ssds, nets, feature_layer, number_box, num_classes):
    """(key, shape) list of the reference `create_model(cfg).state_dict()` for the supported models."""
    ssds = ssds.upper()
    if nets == "MobileNetV2":
        back = mobilenetv2_backbone_shapes()
    elif nets == "RegNetX032":
        back = regnetx032_backbone_shapes()
    else:
        back = resnet_backbone_shapes(nets)
    if ssds == "SSD":
        return back + ssd_neck_shapes(feature_layer, number_box, num_classes)
    if ssds == "YOLOV3":
        return back + yolov3_neck_shapes(feature_layer, number_box, num_classes)
    if ssds == "YOLOV4":
        return back + yolov4_neck_shapes(feature_layer, number_box, num_classes)
    stacks = 0
    if ssds == "SSDBIFPN":
        stacks = 1 if len(feature_layer) == 2 else feature_layer[2]
    return back + fpn_neck_shapes(feature_layer, number_box, num_classes, stacks)


def yolov3_neck_shapes(feature_layer, number_box, num_classes):
    """YOLOV3.add_extras (yolo.py:88-160): transforms.{i} = ConvBNReLU 3x3 on every backbone level but the last,
    extras.{i} = ConvBNReLUx2 (levels) / ConvBNReLU stride 2 ('Conv:S'), per-level heads loc.{l} / conf.{l} =
    Sequential(ConvBNReLU(c, c, 3), Conv2d(c, A*4 | A*C, 3)).  A depth is an int d (neck width d/2) or an
    [in, out] pair (backbone channels `in`, neck width `out`; yolo.py:120-131)."""
    layers, depths = feature_layer
    ints = [l for l in layers if isinstance(l, int)]
    d_in = lambda d: d[0] if isinstance(d, list) else d            # backbone channels of the level
    d_out = lambda d: d[1] if isinstance(d, list) else d // 2      # neck channels of the level
    tr, ex, heads = [], [], []
    in_ch = None
    for idx, (layer, depth) in enumerate(zip(layers, depths)):
        if isinstance(layer, int):
            if layer == ints[-1]:
                ex.append(("x2", d_in(depth), d_out(depth)))
            else:
                prev = depths[idx + 1]
                tr.append((d_out(prev), d_in(depth) // 2))
                ex.append(("x2", int(d_in(depth) * 1.5), d_out(depth)))
            head_c = d_out(depth)
            in_ch = d_in(depth)               # yolo.py:157: the 'Conv:S' chain starts from the RAW last backbone map
        elif layer == "Conv:S":
            ex.append(("s2", in_ch, depth))
            head_c = depth
            in_ch = depth
        else:
            raise ValueError(layer + " does not support by YOLO")
        heads.append(head_c)
    out = []
    for i, (cin, cout) in enumerate(tr):
        out.append((f"transforms.{i}.0.weight", (cout, cin, 3, 3))); _bn_keys(out, f"transforms.{i}.1", cout)
    for i, (kind, cin, cout) in enumerate(ex):
        if kind == "x2":
            out.append((f"extras.{i}.0.weight", (cout // 2, cin, 1, 1))); _bn_keys(out, f"extras.{i}.1", cout // 2)
            out.append((f"extras.{i}.3.weight", (cout, cout // 2, 3, 3))); _bn_keys(out, f"extras.{i}.4", cout)
        else:
            out.append((f"extras.{i}.0.weight", (cout, cin, 3, 3))); _bn_keys(out, f"extras.{i}.1", cout)
    for tower, per in (("loc", 4), ("conf", num_classes)):
        for l, (c, nb) in enumerate(zip(heads, number_box)):
            out.append((f"{tower}.{l}.0.0.weight", (c, c, 3, 3))); _bn_keys(out, f"{tower}.{l}.0.1", c)
            out.append((f"{tower}.{l}.1.weight", (nb * per, c, 3, 3)))
            out.append((f"{tower}.{l}.1.bias", (nb * per,)))
    return out


def yolov4_neck_shapes(feature_layer, number_box, num_classes):
    """YOLOV4.add_extras (yolo.py:325-392): transforms.{i} = ConvBNReLU(depth, depth/2, 3) per backbone level, the
    last one Sequential(ConvBNReLU(depth, depth/2, 3), SPPModule(3), ConvBNReLU(2*depth, depth/2, 3)); extras.{j} =
    ConvBNReLU stride 2 per 'Conv:S'; fpn.{s} = PANModule(depths/2) (yolo.py:187-246: top-down-{i}-to-{i-1} 3x3 +
    ConvBNReLUx2, then bottom-up-{i}-to-{i+1} 3x3 stride 2 + ConvBNReLUx2); heads as in YOLOV3.  Key order =
    registration order of the reference modules: transforms, extras, fpn, loc, conf."""
    layers, depths = feature_layer[0], feature_layer[1]
    stacks = 1 if len(feature_layer) == 2 else feature_layer[2]
    ints = [l for l in layers if isinstance(l, int)]
    out, heads, chans = [], [], []
    in_ch, ei = None, 0
    extras = []
    for idx, (layer, depth) in enumerate(zip(layers, depths)):
        if isinstance(layer, int):
            i = len(chans)
            chans.append(depth // 2)
            if layer == ints[-1]:
                out.append((f"transforms.{i}.0.0.weight", (depth // 2, depth, 3, 3))); _bn_keys(out, f"transforms.{i}.0.1", depth // 2)
                out.append((f"transforms.{i}.2.0.weight", (depth // 2, depth * 2, 3, 3))); _bn_keys(out, f"transforms.{i}.2.1", depth // 2)
            else:
                out.append((f"transforms.{i}.0.weight", (depth // 2, depth, 3, 3))); _bn_keys(out, f"transforms.{i}.1", depth // 2)
            in_ch = depth // 2
        elif layer == "Conv:S":
            extras.append((in_ch, depth))
            in_ch = depth
        else:
            raise ValueError(layer + " does not support by YOLO")
        heads.append(in_ch)
    for j, (cin, cout) in enumerate(extras):
        out.append((f"extras.{j}.0.weight", (cout, cin, 3, 3))); _bn_keys(out, f"extras.{j}.1", cout)
    n = len(chans)

    def x2(prefix, cin, cout):
        out.append((prefix + ".0.weight", (cout // 2, cin, 1, 1))); _bn_keys(out, prefix + ".1", cout // 2)
        out.append((prefix + ".3.weight", (cout, cout // 2, 3, 3))); _bn_keys(out, prefix + ".4", cout)

    for st in range(stacks):
        for i in range(n - 1, 0, -1):
            pre = f"fpn.{st}.top-down-{i}-to-{i - 1}"
            out.append((pre + ".0.weight", (chans[i - 1], chans[i], 3, 3))); _bn_keys(out, pre + ".1", chans[i - 1])
            x2(f"fpn.{st}.top-down-{i - 1}", chans[i - 1] * 2, chans[i - 1])
        for i in range(0, n - 1):
            pre = f"fpn.{st}.bottom-up-{i}-to-{i + 1}"
            out.append((pre + ".0.weight", (chans[i + 1], chans[i], 3, 3))); _bn_keys(out, pre + ".1", chans[i + 1])
            x2(f"fpn.{st}.bottom-up-{i + 1}", chans[i + 1] * 2, chans[i + 1])
    for tower, per in (("loc", 4), ("conf", num_classes)):
        for l, (c, nb) in enumerate(zip(heads, number_box)):
            out.append((f"{tower}.{l}.0.0.weight", (c, c, 3, 3))); _bn_keys(out, f"{tower}.{l}.0.1", c)
            out.append((f"{tower}.{l}.1.weight", (nb * per, c, 3, 3)))
            out.append((f"{tower}.{l}.1.bias", (nb * per,)))
    return out


def ssd_neck_shapes(feature_layer, number_box, num_classes):
    out = []
    in_ch, ei = None, 0
    for layer, depth in zip(feature_layer[0], feature_layer[1]):
        if not isinstance(layer, int):
            p = f"extras.{ei}"
            out.append((p + ".0.weight", (depth // 2, in_ch, 1, 1))); _bn_keys(out, p + ".1", depth // 2)
            out.append((p + ".3.weight", (depth, depth // 2, 3, 3))); _bn_keys(out, p + ".4", depth)
            ei += 1
        in_ch = depth
    for l, (depth, nb) in enumerate(zip(feature_layer[1], number_box)):
        out.append((f"loc.{l}.weight", (nb * 4, depth, 3, 3)))
        out.append((f"loc.{l}.bias", (nb * 4,)))
    for l, (depth, nb) in enumerate(zip(feature_layer[1], number_box)):
        out.append((f"conf.{l}.weight", (nb * num_classes, depth, 3, 3)))
        out.append((f"conf.{l}.bias", (nb * num_classes,)))
    return out


def ssd_mobilenetv2_shapes(nets, feature_layer, number_box, num_classes):
    return mobilenetv2_backbone_shapes() + ssd_neck_shapes(feature_layer, number_box, num_classes)


def ssdfpn_resnet_shapes(nets, feature_layer, number_box, num_classes):
    """SSDFPN (fpn.py:36-146): transforms.{i} 1x1 laterals with bias, extras.{i} ConvBNReLU 3x3,
    shared towers loc/conf = 4 x ConvBNReLU(256,256,3) + Conv2d(256, A*4 | A*C, 3)."""
    out = resnet_backbone_shapes(nets)
    ti = 0
    for layer, depth in zip(feature_layer[0], feature_layer[1]):
        if isinstance(layer, int):
            out.append((f"transforms.{ti}.weight", (256, depth, 1, 1)))
            out.append((f"transforms.{ti}.bias", (256,)))
            ti += 1
    for i, (layer, depth) in enumerate(zip(feature_layer[0], feature_layer[1])):
        cin = 256 if isinstance(layer, int) else depth
        out.append((f"extras.{i}.0.weight", (256, cin, 3, 3)))
        _bn_keys(out, f"extras.{i}.1", 256)
    for tower, cout in (("loc", number_box[0] * 4), ("conf", number_box[0] * num_classes)):
        for j in range(4):
            out.append((f"{tower}.{j}.0.weight", (256, 256, 3, 3)))
            _bn_keys(out, f"{tower}.{j}.1", 256)
        out.append((f"{tower}.4.weight", (cout, 256, 3, 3)))
        out.append((f"{tower}.4.bias", (cout,)))
    return out


def ssd_resnet_shapes(nets, feature_layer, number_box, num_classes):
    """OrderedDict-like list of (key, shape) in the reference's state_dict order."""
    block, layers = RESNETS[nets]
    exp = 4 if block == "bottleneck" else 1
    out = []

    def bn(p, c):
        _bn_keys(out, p, c)

    out.append(("backbone.conv1.weight", (64, 3, 7, 7)))
    bn("backbone.bn1", 64)
    inplanes = 64
    for li, nblocks in enumerate(layers, start=1):
        planes = 64 * 2 ** (li - 1)
        for bi in range(nblocks):
            p = f"backbone.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            if block == "bottleneck":
                out.append((p + ".conv1.weight", (planes, inplanes, 1, 1))); bn(p + ".bn1", planes)
                out.append((p + ".conv2.weight", (planes, planes, 3, 3))); bn(p + ".bn2", planes)
                out.append((p + ".conv3.weight", (planes * 4, planes, 1, 1))); bn(p + ".bn3", planes * 4)
            else:
                out.append((p + ".conv1.weight", (planes, inplanes, 3, 3))); bn(p + ".bn1", planes)
                out.append((p + ".conv2.weight", (planes, planes, 3, 3))); bn(p + ".bn2", planes)
            if stride != 1 or inplanes != planes * exp:
                out.append((p + ".downsample.0.weight", (planes * exp, inplanes, 1, 1)))
                bn(p + ".downsample.1", planes * exp)
            inplanes = planes * exp
    out.append(("backbone.fc.weight", (1000, 512 * exp)))
    out.append(("backbone.fc.bias", (1000,)))
    in_ch, ei = None, 0
    for layer, depth in zip(feature_layer[0], feature_layer[1]):
        if not isinstance(layer, int):
            p = f"extras.{ei}"
            out.append((p + ".0.weight", (depth // 2, in_ch, 1, 1))); bn(p + ".1", depth // 2)
            out.append((p + ".3.weight", (depth, depth // 2, 3, 3))); bn(p + ".4", depth)
            ei += 1
        in_ch = depth
    for l, (depth, nb) in enumerate(zip(feature_layer[1], number_box)):
        out.append((f"loc.{l}.weight", (nb * 4, depth, 3, 3)))
        out.append((f"loc.{l}.bias", (nb * 4,)))
    for l, (depth, nb) in enumerate(zip(feature_layer[1], number_box)):
        out.append((f"conf.{l}.weight", (nb * num_classes, depth, 3, 3)))
        out.append((f"conf.{l}.bias", (nb * num_classes,)))
    return out


def synthetic_state_dict(nets, feature_layer, number_box, num_classes, seed=0, style="test",
                         ssds="SSD"):
    """Deterministic weights.
    style "init": the reference's initialisation statistics — kaiming-normal convs, BN (1, 0, 0, 1)
    (torchvision ResNet), xavier extras (ssdsbase.py:27-31), N(0, 0.01) heads with the conf prior bias
    -log(99) (ssdsbase.py:15-25).
    style "test": same, but non-trivial BN statistics so that BN folding is actually exercised, with
    the last BN of each residual block damped to keep activations O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    prior = -math.log((1 - 0.01) / 0.01)
    shapes = model_shapes(ssds, nets, feature_layer, number_box, num_classes)
    bn_prefixes = {k.rsplit(".", 1)[0] for k, _ in shapes if k.endswith("running_mean")}
    fpn_like = ssds.upper() != "SSD"
    if ssds.upper() in ("YOLOV3", "YOLOV4"):
        head_final = {k for k, _ in shapes if k.startswith(("loc.", "conf.")) and k.split(".")[2] == "1"}
    else:
        head_final = {k for k, _ in shapes if k.startswith(("loc.", "conf.")) and
                      k.rsplit(".", 1)[0] not in bn_prefixes and (not fpn_like or k.split(".")[1] == "4")}
    for key, shape in shapes:
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.tensor(0, dtype=torch.long)
        elif key.endswith("running_mean"):
            sd[key] = torch.zeros(shape) if style == "init" else torch.randn(shape, generator=g) * 0.1
        elif key.endswith("running_var"):
            sd[key] = torch.ones(shape) if style == "init" else torch.rand(shape, generator=g) + 0.5
        elif len(shape) == 1 and key.rsplit(".", 1)[0] in bn_prefixes:
            is_w = key.endswith("weight")
            if style == "init":
                sd[key] = torch.ones(shape) if is_w else torch.zeros(shape)
            elif is_w:
                last = key.endswith(("bn3.weight",)) or ("layer" in key and key.endswith("bn2.weight")
                                                         and nets in ("ResNet18", "ResNet34"))
                lo = 0.2 if last else 0.8
                sd[key] = torch.rand(shape, generator=g) * 0.4 + lo
            else:
                sd[key] = torch.randn(shape, generator=g) * 0.1
        elif key.endswith((".w1", ".w2")):           # BiFPN fusion weights (init 0.5, bifpn.py:15-16)
            sd[key] = torch.full(shape, 0.5) if style == "init" else torch.rand(shape, generator=g) + 0.1
        elif key in head_final or (not fpn_like and key.startswith(("loc.", "conf."))):
            if key.endswith("weight"):
                sd[key] = torch.randn(shape, generator=g) * 0.01
            else:
                sd[key] = torch.full(shape, prior if key.startswith("conf.") else 0.0)
        elif key.startswith(("loc.", "conf.")):      # SSDFPN tower convs: N(0, 0.01) (ssdsbase.py:21-25)
            sd[key] = torch.randn(shape, generator=g) * (0.01 if style == "init" else 0.03)
        elif key.startswith("transforms."):
            if key.endswith("weight"):
                k2 = shape[2] * shape[3] if len(shape) == 4 else 1
                fan_in, fan_out = shape[1] * k2, shape[0] * k2
                a = math.sqrt(6.0 / (fan_in + fan_out))
                sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * a
            else:
                sd[key] = torch.zeros(shape) if style == "init" else torch.randn(shape, generator=g) * 0.05
        elif key.startswith(("backbone.fc", "backbone.classifier", "backbone.head.fc")):
            sd[key] = torch.zeros(shape)
        elif key.startswith(("extras.", "stack_bifpn.")):
            fan_in = shape[1] * shape[2] * shape[3]
            fan_out = shape[0] * shape[2] * shape[3]
            a = math.sqrt(6.0 / (fan_in + fan_out))
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * a
        else:  # backbone conv: kaiming normal, fan_out, relu
            fan_out = shape[0] * shape[2] * shape[3]
            sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
    return sd


def synthetic_targets(B, T=32, seed=4321):
    """SURVEY 8d cfg-4 targets [B,T,5] = (x, y, w, h, label): per image n~U{1..T} boxes, xy~U(0,480),
    wh~U(16,256), label~U{0..79}; unused rows are -1 (the reference's padding, dataset_factory.py:29-35)."""
    g = torch.Generator().manual_seed(seed)
    tg = torch.full((B, T, 5), -1.0)
    for b in range(B):
        n = int(torch.randint(1, T + 1, (1,), generator=g))
        tg[b, :n, :2] = torch.rand((n, 2), generator=g) * 480
        tg[b, :n, 2:4] = torch.rand((n, 2), generator=g) * 240 + 16
        tg[b, :n, 4] = torch.randint(0, 80, (n,), generator=g).float()
    return tg
