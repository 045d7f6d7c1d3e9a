"""CPU oracle for the conv-stack half of the hot path (torch functional ops, fp32).

TEST INFRASTRUCTURE ONLY (same rule as oracle/box_oracle.py): not importable from the product.

Restates the reference graph with plain `torch.nn.functional` calls on a `state_dict`:
    ssds/modeling/ssds/ssd.py:42-74       SSD.forward
    ssds/modeling/nets/resnet.py:41-56    ResNet.forward
    torchvision.models.resnet Bottleneck/BasicBlock forward (v1.5, stride on the 3x3)
    ssds/modeling/layers/basic_layers.py:41-57  ConvBNReLUx2
The convolution/BN arithmetic itself lives in PyTorch (third party, unpinned by the reference —
SURVEY 8c); the oracle calls the same library (torch 2.11 CPU) the reference would.
Pinned against the reference's own nn.Module forward by tests/golden/model_small.npz
(tests/golden/make_golden_model.py).

policy="fp32": the reference's arithmetic.  policy="bf16": the rounding policy of the B200 path —
BN folded into the weights, weights and every stored activation rounded to bf16, fp32 accumulation,
fp32 head outputs — so that the CUDA path can be compared within accumulation-order noise.
"""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _r(t, policy):
    return t.to(torch.bfloat16).float() if policy == "bf16" else t


def _conv_bn(x, sd, wkey, bnkey, stride, pad, relu, policy, residual=None, bias=None, groups=1):
    """relu: False none, True ReLU, 6 ReLU6."""
    w = sd[wkey].float()
    if policy == "bf16":
        b = bias.float() if bias is not None else torch.zeros(w.shape[0], device=w.device)
        if bnkey is not None:
            scale = sd[bnkey + ".weight"].float() / torch.sqrt(sd[bnkey + ".running_var"].float() + EPS)
            w = w * scale.view(-1, 1, 1, 1)
            b = (b - sd[bnkey + ".running_mean"].float()) * scale + sd[bnkey + ".bias"].float()
        y = F.conv2d(x, _r(w, policy), b, stride=stride, padding=pad, groups=groups)
    else:
        y = F.conv2d(x, w, bias, stride=stride, padding=pad, groups=groups)
        if bnkey is not None:
            y = F.batch_norm(y, sd[bnkey + ".running_mean"], sd[bnkey + ".running_var"],
                             sd[bnkey + ".weight"], sd[bnkey + ".bias"], False, 0.0, EPS)
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    if relu == 6:
        y = y.clamp(max=6.0)
    return y


MBV2_SETTINGS = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2],
                 [6, 320, 1, 1]]


def mobilenetv2_features(sd, x, outputs, policy):
    """reference mobilenet.py:180-192 (MobileNetEx.forward, v2) with torchvision InvertedResidual."""
    x = _r(x.float(), policy)
    x = _r(_conv_bn(x, sd, "backbone.conv1.0.weight", "backbone.conv1.1", 2, 1, 6, policy), policy)
    feats = []
    inp = 32
    for j, (t, c, n, s_) in enumerate(MBV2_SETTINGS, start=1):
        if j > max(outputs):
            break
        for i in range(n):
            p = f"backbone.layer{j}.{i}.conv"
            stride = s_ if i == 0 else 1
            hid = inp * t
            y, k = x, 0
            if t != 1:
                y = _r(_conv_bn(y, sd, f"{p}.0.0.weight", f"{p}.0.1", 1, 0, 6, policy), policy)
                k = 1
            y = _r(_conv_bn(y, sd, f"{p}.{k}.0.weight", f"{p}.{k}.1", stride, 1, 6, policy, groups=hid), policy)
            res = x if (stride == 1 and inp == c) else None
            x = _r(_conv_bn(y, sd, f"{p}.{k + 1}.weight", f"{p}.{k + 2}", 1, 0, False, policy, residual=res), policy)
            inp = c
        if j in outputs:
            feats.append(x)
    return feats


def resnet_features(sd, x, outputs, policy):
    """reference resnet.py:41-56: list of the backbone feature maps named in `outputs`."""
    x = _r(x.float(), policy)
    x = _r(_conv_bn(x, sd, "backbone.conv1.weight", "backbone.bn1", 2, 3, True, policy), policy)
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li in range(1, 5):
        if li + 1 > max(outputs):
            break
        bi = 0
        while f"backbone.layer{li}.{bi}.conv1.weight" in sd:
            p = f"backbone.layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            identity = x
            if (p + ".downsample.0.weight") in sd:
                identity = _r(_conv_bn(x, sd, p + ".downsample.0.weight", p + ".downsample.1", stride, 0,
                                       False, policy), policy)
            if (p + ".conv3.weight") in sd:
                y = _r(_conv_bn(x, sd, p + ".conv1.weight", p + ".bn1", 1, 0, True, policy), policy)
                y = _r(_conv_bn(y, sd, p + ".conv2.weight", p + ".bn2", stride, 1, True, policy), policy)
                x = _r(_conv_bn(y, sd, p + ".conv3.weight", p + ".bn3", 1, 0, True, policy, identity), policy)
            else:
                y = _r(_conv_bn(x, sd, p + ".conv1.weight", p + ".bn1", stride, 1, True, policy), policy)
                x = _r(_conv_bn(y, sd, p + ".conv2.weight", p + ".bn2", 1, 1, True, policy, identity), policy)
            bi += 1
        if li + 1 in outputs:
            feats.append(x)
    return feats


def ssd_resnet_forward(sd, x, feature_layer, training=False, policy="fp32", backbone="resnet"):
    """x fp32 NCHW (already normalised).  Returns (tuple loc, tuple conf) like ssd.py:42-74."""
    outputs = [l for l in feature_layer[0] if isinstance(l, int)]
    feats = (mobilenetv2_features if backbone == "mobilenetv2" else resnet_features)(sd, x, outputs, policy)
    ei = 0
    for layer in feature_layer[0]:
        if isinstance(layer, int):
            continue
        stride = 2 if layer == "Conv:S" else 1
        p = f"extras.{ei}"
        y = _r(_conv_bn(feats[-1], sd, p + ".0.weight", p + ".1", 1, 0, True, policy), policy)
        y = _r(_conv_bn(y, sd, p + ".3.weight", p + ".4", stride, 1, True, policy), policy)
        feats.append(y)
        ei += 1
    loc, conf = [], []
    for l, f in enumerate(feats):
        loc.append(_conv_bn(f, sd, f"loc.{l}.weight", None, 1, 1, False, policy, bias=sd[f"loc.{l}.bias"]))
        c = _conv_bn(f, sd, f"conf.{l}.weight", None, 1, 1, False, policy, bias=sd[f"conf.{l}.bias"])
        conf.append(c if training else torch.sigmoid(c))
    return tuple(loc), tuple(conf)


def ssdfpn_resnet_forward(sd, x, feature_layer, training=False, policy="fp32"):
    """reference fpn.py:58-101 (SSDFPN.forward) on a state_dict.

    bf16 policy mirrors the B200 path: the lateral conv output is stored (rounded) before the
    upsample-add, the sum is stored again; every tower layer output is stored in bf16."""
    outputs = [l for l in feature_layer[0] if isinstance(l, int)]
    feats = resnet_features(sd, x, outputs, policy)
    n_back = len(feats)
    raw_top = feats[-1]
    pyr = [None] * n_back
    xx = None
    for i in range(n_back - 1, -1, -1):                                   # fpn.py:79-87
        lat = _r(_conv_bn(feats[i], sd, f"transforms.{i}.weight", None, 1, 0, False, policy,
                          bias=sd[f"transforms.{i}.bias"]), policy)
        if i != n_back - 1:
            lat = _r(F.interpolate(xx, scale_factor=2, mode="nearest") + lat, policy)
        xx = lat
        pyr[i] = lat
    loc, conf = [], []
    for i, layer in enumerate(feature_layer[0]):                          # fpn.py:89-97
        stride = 1 if isinstance(layer, int) else 2
        src = pyr[i] if i < n_back else (raw_top if i == n_back else xx)
        xx = _r(_conv_bn(src, sd, f"extras.{i}.0.weight", f"extras.{i}.1", stride, 1, True, policy), policy)
        outs = []
        for tower in ("loc", "conf"):                                     # SharedHead fpn.py:10-18
            t = xx
            for j in range(4):
                t = _r(_conv_bn(t, sd, f"{tower}.{j}.0.weight", f"{tower}.{j}.1", 1, 1, True, policy), policy)
            outs.append(_conv_bn(t, sd, f"{tower}.4.weight", None, 1, 1, False, policy,
                                 bias=sd[f"{tower}.4.bias"]))
        loc.append(outs[0])
        conf.append(outs[1] if training else torch.sigmoid(outs[1]))
    return tuple(loc), tuple(conf)


def yolov3_resnet_forward(sd, x, feature_layer, training=False, policy="fp32"):
    """reference yolo.py:44-87 (YOLOV3.forward) on a state_dict, int depths; ConvBNReLU / ConvBNReLUx2 of
    basic_layers.py:27-57 (plain ReLU).  bf16 policy: every conv output and the concatenated map are stored."""
    layers = feature_layer[0]
    outputs = [l for l in layers if isinstance(l, int)]
    feats = resnet_features(sd, x, outputs, policy)
    n_back = len(feats)
    raw_top = feats[-1]
    xx = feats[-1]
    for i in range(n_back - 1, -1, -1):                                   # yolo.py:67-74
        if i != n_back - 1:
            t = _r(_conv_bn(xx, sd, f"transforms.{i}.0.weight", f"transforms.{i}.1", 1, 1, True, policy), policy)
            xx = torch.cat((feats[i], F.interpolate(t, scale_factor=2)), dim=1)
        else:
            xx = feats[i]
        p = f"extras.{i}"
        xx = _r(_conv_bn(xx, sd, p + ".0.weight", p + ".1", 1, 0, True, policy), policy)
        xx = _r(_conv_bn(xx, sd, p + ".3.weight", p + ".4", 1, 1, True, policy), policy)
        feats[i] = xx
    loc, conf = [], []
    for i in range(len(layers)):                                          # yolo.py:77-84
        if i < n_back:
            xx = feats[i]
        else:
            src = raw_top if i == n_back else xx
            xx = _r(_conv_bn(src, sd, f"extras.{i}.0.weight", f"extras.{i}.1", 2, 1, True, policy), policy)
        outs = []
        for tower in ("loc", "conf"):
            t = _r(_conv_bn(xx, sd, f"{tower}.{i}.0.0.weight", f"{tower}.{i}.0.1", 1, 1, True, policy), policy)
            outs.append(_conv_bn(t, sd, f"{tower}.{i}.1.weight", None, 1, 1, False, policy,
                                 bias=sd[f"{tower}.{i}.1.bias"]))
        loc.append(outs[0])
        conf.append(outs[1] if training else torch.sigmoid(outs[1]))
    return tuple(loc), tuple(conf)


def yolov4_resnet_forward(sd, x, feature_layer, training=False, policy="fp32"):
    """reference yolo.py:249-323 (YOLOV4.forward): transforms (the last one with the SPP block, yolo.py:161-184:
    max-pools of 5 / 9 / 13, stride 1, concatenated after x), PANModule.forward (yolo.py:222-246: top-down with
    nearest-2x upsample + concat + ConvBNReLUx2, then bottom-up with stride-2 conv + concat + ConvBNReLUx2),
    'Conv:S' extras chained from the last PAN level, per-level heads.  bf16 policy: every conv output is stored
    (max-pooling and concatenation of stored values are exact)."""
    layers = feature_layer[0]
    stacks = 1 if len(feature_layer) == 2 else feature_layer[2]
    outputs = [l for l in layers if isinstance(l, int)]
    feats = resnet_features(sd, x, outputs, policy)
    n = len(feats)

    def cbr(t, conv, bn, stride=1, pad=1):
        return _r(_conv_bn(t, sd, conv, bn, stride, pad, True, policy), policy)

    def x2(t, p):
        t = cbr(t, p + ".0.weight", p + ".1", 1, 0)
        return cbr(t, p + ".3.weight", p + ".4", 1, 1)

    xx = []
    for i in range(n):
        if i == n - 1:
            t = cbr(feats[i], f"transforms.{i}.0.0.weight", f"transforms.{i}.0.1")
            pools = [t] + [F.max_pool2d(t, kernel_size=k, stride=1, padding=(k - 1) // 2) for k in (5, 9, 13)]
            t = cbr(torch.cat(pools, dim=1), f"transforms.{i}.2.0.weight", f"transforms.{i}.2.1")
        else:
            t = cbr(feats[i], f"transforms.{i}.0.weight", f"transforms.{i}.1")
        xx.append(t)
    for st in range(stacks):
        for i in range(n - 1, 0, -1):
            t = cbr(xx[i], f"fpn.{st}.top-down-{i}-to-{i - 1}.0.weight", f"fpn.{st}.top-down-{i}-to-{i - 1}.1")
            xx[i - 1] = x2(torch.cat((xx[i - 1], F.interpolate(t, scale_factor=2, mode="nearest")), dim=1),
                           f"fpn.{st}.top-down-{i - 1}")
        for i in range(0, n - 1):
            t = cbr(xx[i], f"fpn.{st}.bottom-up-{i}-to-{i + 1}.0.weight", f"fpn.{st}.bottom-up-{i}-to-{i + 1}.1", 2, 1)
            xx[i + 1] = x2(torch.cat((xx[i + 1], t), dim=1), f"fpn.{st}.bottom-up-{i + 1}")
    t, j = xx[-1], 0
    while f"extras.{j}.0.weight" in sd:
        t = cbr(t, f"extras.{j}.0.weight", f"extras.{j}.1", 2, 1)
        xx.append(t)
        j += 1
    loc, conf = [], []
    for i, f in enumerate(xx):
        outs = []
        for tower in ("loc", "conf"):
            h = cbr(f, f"{tower}.{i}.0.0.weight", f"{tower}.{i}.0.1")
            outs.append(_conv_bn(h, sd, f"{tower}.{i}.1.weight", None, 1, 1, False, policy,
                                 bias=sd[f"{tower}.{i}.1.bias"]))
        loc.append(outs[0])
        conf.append(outs[1] if training else torch.sigmoid(outs[1]))
    return tuple(loc), tuple(conf)


def ssd_mobilenetv2_forward(sd, x, feature_layer, training=False, policy="fp32"):
    """SSD.forward (ssd.py:42-74) over the MobileNetV2 backbone."""
    return ssd_resnet_forward(sd, x, feature_layer, training, policy, backbone="mobilenetv2")


REGNET_GW = 48


def regnetx_features(sd, x, outputs, policy):
    """reference regnet.py:270-282 (RegNet.forward) with SimpleStemIN / ResBottleneckBlock (:28-107)."""
    x = _r(x.float(), policy)
    x = _r(_conv_bn(x, sd, "backbone.stem.conv.weight", "backbone.stem.bn", 2, 1, True, policy), policy)
    feats = []
    for si in range(1, 5):
        if si > max(outputs):
            break
        bi = 1
        while f"backbone.s{si}.b{bi}.f.a.weight" in sd:
            p = f"backbone.s{si}.b{bi}"
            stride = 2 if bi == 1 else 1
            identity = x
            if (p + ".proj.weight") in sd:
                identity = _r(_conv_bn(x, sd, p + ".proj.weight", p + ".bn", stride, 0, False, policy), policy)
            w_b = sd[p + ".f.b.weight"].shape[0]
            y = _r(_conv_bn(x, sd, p + ".f.a.weight", p + ".f.a_bn", 1, 0, True, policy), policy)
            y = _r(_conv_bn(y, sd, p + ".f.b.weight", p + ".f.b_bn", stride, 1, True, policy,
                            groups=w_b // REGNET_GW), policy)
            x = _r(_conv_bn(y, sd, p + ".f.c.weight", p + ".f.c_bn", 1, 0, True, policy, residual=identity), policy)
            bi += 1
        if si in outputs:
            feats.append(x)
    return feats


def _shared_heads(sd, xx, training, policy):
    outs = []
    for tower in ("loc", "conf"):                                     # SharedHead fpn.py:10-18
        t = xx
        for j in range(4):
            t = _r(_conv_bn(t, sd, f"{tower}.{j}.0.weight", f"{tower}.{j}.1", 1, 1, True, policy), policy)
        outs.append(_conv_bn(t, sd, f"{tower}.4.weight", None, 1, 1, False, policy, bias=sd[f"{tower}.4.bias"]))
    return outs[0], (outs[1] if training else torch.sigmoid(outs[1]))


def ssdbifpn_forward(sd, x, feature_layer, training=False, policy="fp32", backbone="regnetx"):
    """reference bifpn.py:104-142 (SSDBiFPN.forward) + BiFPNModule.forward (:30-63) on a state_dict."""
    outputs = [l for l in feature_layer[0] if isinstance(l, int)]
    feats = (regnetx_features if backbone == "regnetx" else resnet_features)(sd, x, outputs, policy)
    n_back = len(feats)
    raw_top = feats[-1]
    xx = [_r(_conv_bn(f, sd, f"transforms.{i}.weight", None, 1, 0, False, policy,
                      bias=sd[f"transforms.{i}.bias"]), policy) for i, f in enumerate(feats)]
    s_ = 0
    while f"stack_bifpn.{s_}.w1" in sd:
        p = f"stack_bifpn.{s_}"
        w1 = F.relu(sd[p + ".w1"].float())
        w1 = w1 / (w1.sum(0) + 1e-6)
        w2 = F.relu(sd[p + ".w2"].float())
        w2 = w2 / (w2.sum(0) + 1e-6)
        L = n_back
        xs = list(xx)
        for i in range(L - 1, 0, -1):
            f_ = _r(w1[0, i - 1] * xx[i - 1] + w1[1, i - 1] * F.interpolate(xx[i], scale_factor=2, mode="nearest"),
                    policy)
            xx[i - 1] = _r(_conv_bn(f_, sd, f"{p}.top-down-{i - 1}.0.weight", f"{p}.top-down-{i - 1}.1", 1, 1, True,
                                    policy), policy)
        for i in range(0, L - 2):
            f_ = _r(w2[0, i] * xx[i + 1] + w2[1, i] * F.max_pool2d(xx[i], kernel_size=2) + w2[2, i] * xs[i + 1],
                    policy)
            xx[i + 1] = _r(_conv_bn(f_, sd, f"{p}.bottom-up-{i + 1}.0.weight", f"{p}.bottom-up-{i + 1}.1", 1, 1,
                                    True, policy), policy)
        f_ = _r(w1[0, L - 1] * xx[L - 1] + w1[1, L - 1] * F.max_pool2d(xx[L - 2], kernel_size=2), policy)
        xx[L - 1] = _r(_conv_bn(f_, sd, f"{p}.bottom-up-{L - 1}.0.weight", f"{p}.bottom-up-{L - 1}.1", 1, 1, True,
                                policy), policy)
        s_ += 1
    loc, conf = [], []
    cur = None
    for i, layer in enumerate(feature_layer[0]):                          # bifpn.py:131-139
        stride = 1 if isinstance(layer, int) else 2
        src = xx[i] if i < n_back else (raw_top if i == n_back else cur)
        cur = _r(_conv_bn(src, sd, f"extras.{i}.0.weight", f"extras.{i}.1", stride, 1, True, policy), policy)
        l, c = _shared_heads(sd, cur, training, policy)
        loc.append(l)
        conf.append(c)
    return tuple(loc), tuple(conf)
