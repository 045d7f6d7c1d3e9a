"""SSD + ResNet conv stack on the tcgen05 implicit-GEMM kernel — the drop-in for `model(x)`.

Mirrors the reference graph
    ssds/modeling/ssds/ssd.py:42-74        SSD.forward (backbone -> extras -> loc/conf heads, eval sigmoid)
    ssds/modeling/nets/resnet.py:41-56     ResNet.forward (conv1/bn1/relu/maxpool, layer1..4, outputs)
    torchvision.models.resnet Bottleneck / BasicBlock (v1.5: stride on the 3x3)
    ssds/modeling/layers/basic_layers.py:41-57  ConvBNReLUx2 ("Conv:S" / "Conv" extras)
and consumes the reference's own `state_dict()` (same key names), folding eval-mode BatchNorm into
bf16 weights once at construction.  `forward(x) -> (tuple loc_l [B,A*4,H,W], tuple conf_l
[B,A*C,H,W])` in fp32 NCHW exactly like the reference; conf is sigmoid-ed iff eval (ssd.py:72-73).

Every conv (+BN +ReLU +residual) is one kernel launch; each level's loc+conf pair is one fused
launch.  For a fixed input shape the launch sequence is recorded once into static buffers and can be
replayed as a CUDA graph (`use_graph=True`).
"""
import math
import os
from collections import OrderedDict

import torch

from . import conv as K
from .box import configure_ratio_scale, generate_anchors

BN_EPS = 1e-5
# conv3 -> next conv1 pair launches (conv_pair.cu) are used from this many M-tiles per SM upwards
# (tests lower it to exercise the pairing on small inputs)
PAIR_MIN_TILES_PER_SM = 8


def _bn(sd, prefix):
    return (sd[prefix + ".weight"], sd[prefix + ".bias"], sd[prefix + ".running_mean"],
            sd[prefix + ".running_var"], BN_EPS)


def _cpad(c, m=32):
    return (c + m - 1) // m * m


class _Conv:
    """One packed conv: bf16 weights [Cout_pad, taps, Cin_pad] + fp32 bias on the device.
    relu: False/0 none, True/1 ReLU, 2 ReLU6.  cin_pad / cout_pad add zero channels so that every
    activation tensor has a channel count the K-block (and the 32-column epilogue) can address."""

    def __init__(self, weight, bn, bias, stride, pad, relu, device, stem=False, cin_pad=None,
                 cout_pad=None):
        w, b = K.fold_bn(weight, bn, bias)
        self.KH, self.KW = (4, 4) if stem else (weight.shape[2], weight.shape[3])
        self.stride, self.pad, self.relu = (1, 2, int(relu)) if stem else (stride, pad, int(relu))
        cout = weight.shape[0]
        self.cout = cout_pad or cout
        wp = K.pack_stem_weight_s2d(w) if stem else K.pack_weight(w, cin_pad)
        if self.cout != cout:
            wp = torch.cat([wp, torch.zeros((self.cout - cout,) + tuple(wp.shape[1:]), dtype=wp.dtype)], 0)
            b = torch.cat([b, torch.zeros(self.cout - cout)], 0)
        self.w = wp.contiguous().to(device)
        self.bias = b.contiguous().to(device)
        self.flops_per_pixel = 2 * weight[0].numel() * cout        # algorithmic (un-padded) FLOPs


_MB_CHOICE = {}      # (block shape) -> True when the fused inverted-residual launch measured faster than three launches


def _graph_time(fn, reps=4):
    """milliseconds of `reps` back-to-back calls of fn replayed from one CUDA graph (no per-launch host time)."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(reps):
                fn()
        gr.replay()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        gr.replay()
        e1.record(st)
        st.synchronize()
    torch.cuda.current_stream().wait_stream(st)
    return e0.elapsed_time(e1)


class _DwConv:
    """Depthwise 3x3 + folded BN + activation: bf16 weights [9, C_pad], fp32 bias."""

    def __init__(self, weight, bn, stride, relu, device, c_pad=None):
        w, b = K.fold_bn(weight, bn, None)
        c = weight.shape[0]
        self.c = c_pad or c
        self.stride, self.relu = stride, int(relu)
        self.w = K.pack_dw_weight(w, self.c).to(device)
        self.bias = torch.cat([b, torch.zeros(self.c - c)], 0).contiguous().to(device)
        self.flops_per_pixel = 2 * 9 * c


class _Steps(list):
    """The recorded launch sequence.  Every step carries the id of the BRANCH it was recorded on (0 = main).
    Branches express independence between chains of small launches (the per-level towers of FPN / BiFPN /
    YOLOv3 heads, SSD extras): under CUDA-graph capture each branch is captured on its own stream, forked from
    the main stream where its first step was recorded (it depends on everything recorded on main before that)
    or from another branch (`fork(new, after=parent)`), and joined back at the end of the graph, so that kernels
    too small to fill the GPU run side by side.  Eager replay runs the steps in recording order on one stream —
    same kernels, same values."""

    def __init__(self):
        super().__init__()
        self.tags = []
        self.parents = {}          # branch id -> (parent branch id, number of steps recorded at fork time)
        self.cur = 0
        self._next = 1

    def append(self, fn):
        super().append(fn)
        self.tags.append(self.cur)

    def fork(self, after=0):
        """Start a new branch that depends on everything recorded so far on branch `after`; returns its id."""
        b = self._next
        self._next += 1
        self.parents[b] = (after, len(self))
        return b

    def on(self, branch):
        steps = self

        class _Ctx:
            def __enter__(self_):
                self_.prev = steps.cur
                steps.cur = branch

            def __exit__(self_, *a):
                steps.cur = self_.prev
        return _Ctx()


class _Engine(torch.nn.Module):
    """Shared machinery: plan recording (static buffers), CUDA-graph replay.  Subclasses provide a
    backbone (`_build_backbone`, `_plan_backbone`) and a neck (`_build_neck`, `_plan_neck`)."""

    def __init__(self, state_dict, feature_layer, num_classes, number_box, device="cuda",
                 mean=0.0, std=1.0):
        super().__init__()
        sd = {k[7:] if k.startswith("module.") else k: v for k, v in state_dict.items()}
        self.device = torch.device(device)
        self.num_classes = num_classes
        self.number_box = list(number_box)
        self.mean, self.std = float(mean), float(std)
        self._plans = {}
        self.training = False
        self._build_backbone(sd, feature_layer)
        self._build_neck(sd, feature_layer)


    # ------------------------------------------------------------------ plan (static buffers)
    def _build_plan(self, images):
        """Record the launch sequence for this input shape/dtype into preallocated buffers."""
        dev = self.device
        if images.dtype == torch.uint8:
            N, H, W, _ = images.shape
        else:
            N, _, H, W = images.shape
        steps = _Steps()    # list of zero-arg callables (+ the branch each one was recorded on)
        info = {}           # step index -> {kind, flops, bytes}: algorithmic work of that launch (profiling tools)
        flops = [0]
        split = [None]
        bf = torch.bfloat16

        def buf(n, h, w, c):
            return torch.empty((n, h, w, c), dtype=bf, device=dev)

        src = torch.empty_like(images, device=dev)
        # s2d image with 2 zero pixels of left padding per row (zeroed once; the pack kernel never
        # touches the padding) so the stem can fetch 4 taps x 16 ch as one 128-byte window
        packed = torch.zeros((N, H // 2, W // 2 + K.STEM_ROW_EXTRA, 16), dtype=bf, device=dev)
        mean, std = self.mean, self.std
        steps.append(lambda: K.pack_image_s2d(src, mean, std, out=packed))

        def add_conv(cv, x, residual=None, relu=None, Ho=0, Wo=0, x_kind=0, x_width=None, out=None):
            n, h, w, _ = x.shape
            ho = Ho or (h + 2 * cv.pad - cv.KH) // cv.stride + 1
            wo = Wo or (w + 2 * cv.pad - cv.KW) // cv.stride + 1
            y = out if out is not None else buf(n, ho, wo, cv.cout)
            r = cv.relu if relu is None else relu
            steps.append(lambda: K.conv2d(x, cv.w, cv.bias, cv.KH, cv.KW, cv.stride, cv.pad, r,
                                          residual, out=y, Ho=ho, Wo=wo, x_kind=x_kind,
                                          x_width=x_width))
            flops[0] += cv.flops_per_pixel * n * ho * wo
            info[len(steps) - 1] = {"kind": f"conv{cv.KH}x{cv.KW}s{cv.stride} {x.shape[3]}->{cv.cout} @{ho}x{wo}",
                                    "flops": cv.flops_per_pixel * n * ho * wo,
                                    # a strided 1x1 only touches every stride-th pixel of x
                                    "bytes": 2 * (x.numel() // (cv.stride ** 2 if cv.KH == 1 else 1) + cv.w.numel() +
                                                  y.numel() + (residual.numel() if residual is not None else 0))}
            return y

        def add_head(f, h, loc=None, conf=None):
            """fused loc+conf head conv -> fp32 NCHW tensors (n_loc channels go to loc)."""
            n, fh, fw, _ = f.shape
            if loc is None:
                loc = torch.empty((n, h.n_loc, fh, fw), dtype=torch.float32, device=dev)
            if conf is None:
                conf = torch.empty((n, h.cout - h.n_loc, fh, fw), dtype=torch.float32, device=dev)
            steps.append(lambda: K.conv2d_head(f, h.w, h.bias, h.n_loc, not self.training,
                                               loc=loc, conf=conf))
            flops[0] += h.flops_per_pixel * n * fh * fw
            info[len(steps) - 1] = {"kind": f"head3x3 {f.shape[3]}->{h.cout} @{fh}x{fw}",
                                    "flops": h.flops_per_pixel * n * fh * fw,
                                    "bytes": 2 * (f.numel() + h.w.numel()) + 4 * (loc.numel() + conf.numel())}
            return loc, conf

        def add_dw(dw, x):
            n, h, w, c = x.shape
            ho, wo = (h - 1) // dw.stride + 1, (w - 1) // dw.stride + 1
            y = buf(n, ho, wo, c)
            steps.append(lambda: K.dwconv3x3(x, dw.w, dw.bias, dw.stride, dw.relu, out=y))
            flops[0] += dw.flops_per_pixel * n * ho * wo
            info[len(steps) - 1] = {"kind": f"dw3x3s{dw.stride} {c} @{ho}x{wo}", "flops": dw.flops_per_pixel * n * ho * wo,
                                    "bytes": 2 * (x.numel() + y.numel() + dw.w.numel())}
            return y

        def add_raw(fn, nflops, kind="raw", nbytes=0):
            steps.append(fn)
            flops[0] += nflops
            info[len(steps) - 1] = {"kind": kind, "flops": nflops, "bytes": nbytes}

        self._add_raw = add_raw
        self._last_info = lambda: info[len(steps) - 1]
        feats = self._plan_backbone(packed, H, W, steps, buf, add_conv, add_dw)
        # everything after the backbone may write the loc/conf outputs: the graph is split here, and the second
        # part waits for the consumer of the previous step's outputs (decode/NMS on another stream)
        split[0] = len(steps)
        locs, confs = self._plan_neck(feats, steps, buf, add_conv, add_head)
        return {"src": src, "steps": steps, "loc": tuple(locs), "conf": tuple(confs),
                "flops": flops[0], "graph": None, "launches": len(steps), "info": info,
                "split": split[0]}

    def plan_for(self, images):
        key = (tuple(images.shape), images.dtype, self.training)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = self._build_plan(images)
        return plan

    def run_plan(self, plan, use_graph=False, outputs_free=None):
        """Replay the recorded launches.  `outputs_free` (a CUDA event) is waited for right before the
        first launch that overwrites the loc/conf outputs, so that a consumer of the PREVIOUS step's
        outputs (decode/NMS on another stream) can overlap this step's backbone."""
        k = plan["split"]
        if use_graph:
            if plan["graph"] is None:
                for s in plan["steps"]:                # warm-up: sets kernel attributes etc.
                    s()
                torch.cuda.synchronize(self.device)
                plan["graph"] = [self._capture(plan["steps"], 0, k), self._capture(plan["steps"], k, len(plan["steps"]))]
            plan["graph"][0].replay()
            if outputs_free is not None:
                torch.cuda.current_stream().wait_event(outputs_free)
            plan["graph"][1].replay()
        else:
            for i, s in enumerate(plan["steps"]):
                if i == k and outputs_free is not None:
                    torch.cuda.current_stream().wait_event(outputs_free)
                s()
        return plan["loc"], plan["conf"]

    def _capture(self, steps, lo, hi):
        """CUDA graph of steps[lo:hi]; branches (see _Steps) are captured on forked side streams."""
        use_branches = os.environ.get("SSDSB_NO_BRANCH", "0") != "1"
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            main = torch.cuda.current_stream()
            streams = {0: main}
            for i in range(lo, hi):
                tag = steps.tags[i] if use_branches else 0
                st = streams.get(tag)
                if st is None:                                  # first step of this branch inside the part: fork
                    st = streams[tag] = torch.cuda.Stream(device=self.device)
                    parent = steps.parents[tag][0]
                    st.wait_stream(streams.get(parent, main))   # depends on what its parent has recorded so far
                if st is main:
                    steps[i]()
                else:
                    with torch.cuda.stream(st):
                        steps[i]()
            for tag, st in streams.items():                     # join
                if st is not main:
                    main.wait_stream(st)
        return g

    def forward(self, x, use_graph=False, outputs_free=None):
        """x: fp32 NCHW [B,3,H,W] (already `(img-mean)/std`-normalised if mean/std were left at
        0/1) or uint8 NHWC [B,H,W,3].  Returns (tuple loc, tuple conf) like ssd.py:42-74."""
        if not x.is_cuda:
            x = x.to(self.device, non_blocking=True)
        x = x.contiguous()
        plan = self.plan_for(x)
        plan["src"].copy_(x, non_blocking=True)
        return self.run_plan(plan, use_graph, outputs_free)

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = bool(mode)
        return self


class _ResNetBackbone:
    """torchvision-style ResNet under `backbone.` (reference nets/resnet.py:9-56)."""

    def _build_backbone(self, sd, feature_layer):
        dev = self.device
        self.stem = _Conv(sd["backbone.conv1.weight"], _bn(sd, "backbone.bn1"), None, 2, 3, True, dev,
                          stem=True)
        # backbone levels: level = layer index + 1 (resnet.py:48-54)
        self.outputs = [l for l in feature_layer[0] if isinstance(l, int)]
        self.layers = []
        for li in range(1, 5):
            if li + 1 > max(self.outputs):
                break
            blocks = []
            bi = 0
            while f"backbone.layer{li}.{bi}.conv1.weight" in sd:
                p = f"backbone.layer{li}.{bi}"
                bottleneck = (p + ".conv3.weight") in sd
                stride = 2 if (li > 1 and bi == 0) else 1
                blk = {}
                if bottleneck:
                    blk["convs"] = [
                        _Conv(sd[p + ".conv1.weight"], _bn(sd, p + ".bn1"), None, 1, 0, True, dev),
                        _Conv(sd[p + ".conv2.weight"], _bn(sd, p + ".bn2"), None, stride, 1, True, dev),
                        _Conv(sd[p + ".conv3.weight"], _bn(sd, p + ".bn3"), None, 1, 0, True, dev),
                    ]
                else:
                    blk["convs"] = [
                        _Conv(sd[p + ".conv1.weight"], _bn(sd, p + ".bn1"), None, stride, 1, True, dev),
                        _Conv(sd[p + ".conv2.weight"], _bn(sd, p + ".bn2"), None, 1, 1, True, dev),
                    ]
                if (p + ".downsample.0.weight") in sd:
                    blk["down"] = _Conv(sd[p + ".downsample.0.weight"], _bn(sd, p + ".downsample.1"),
                                        None, stride, 0, False, dev)
                blocks.append(blk)
                bi += 1
            self.layers.append(blocks)

    def _plan_backbone(self, packed, H, W, steps, buf, add_conv, add_dw):
        x = add_conv(self.stem, packed, Ho=H // 2, Wo=W // 2, x_kind=1, x_width=W // 2)
        N = x.shape[0]
        pooled = buf(N, (x.shape[1] - 1) // 2 + 1, (x.shape[2] - 1) // 2 + 1, x.shape[3])
        xs = x
        steps.append(lambda: K.maxpool3x3s2(xs, out=pooled))
        x = pooled
        feats = []
        # A bottleneck's conv3 (+bn3 +identity +ReLU) and the NEXT bottleneck's conv1 (+bn1 +ReLU) are both
        # pointwise: they run as one launch (conv_pair.cu) in which conv1 reads each block-output tile
        # back from L2 instead of HBM.  SSDSB_NO_PAIR=1 keeps them separate (profiling / A-B runs).
        use_pair = os.environ.get("SSDSB_NO_PAIR", "0") != "1"
        flat = [(li, bi, blk) for li, blocks in enumerate(self.layers) for bi, blk in enumerate(blocks)]

        def pairable(c3, c1, n, h, w):
            # M-tiles are the scheduling unit of the fused kernel: below ~8 tiles per SM the tail
            # quantisation costs more than the saved HBM read (measured: 32x32 and 16x16 stages at B=64)
            chans = (c3.w.shape[-1], c3.cout, c1.cout)
            return (use_pair and c1.w.shape[-1] == c3.cout and c1.KH == 1 and c1.stride == 1 and
                    n * h * w >= 128 * PAIR_MIN_TILES_PER_SM * K.sm_count() and
                    all(c % 64 == 0 and (c <= 256 or c % 256 == 0) for c in chans))

        pre = None                # conv1 output of the current block when the previous launch made it
        for i, (li, bi, blk) in enumerate(flat):
            identity = add_conv(blk["down"], x) if "down" in blk else x
            convs = blk["convs"]
            if len(convs) == 3:
                y = pre if pre is not None else add_conv(convs[0], x)
                y = add_conv(convs[1], y)
                nxt = flat[i + 1][2]["convs"] if i + 1 < len(flat) else None
                n, h, w, _ = y.shape
                if nxt is not None and len(nxt) == 3 and pairable(convs[2], nxt[0], n, h, w):
                    c3, c1 = convs[2], nxt[0]
                    out1, out2 = buf(n, h, w, c3.cout), buf(n, h, w, c1.cout)
                    self._add_raw(lambda x=y, r=identity, c3=c3, c1=c1, o1=out1, o2=out2: K.conv1x1_pair(
                        x, c3.w, c3.bias, True, r, c1.w, c1.bias, c1.relu, out1=o1, out2=o2),
                        (c3.flops_per_pixel + c1.flops_per_pixel) * n * h * w,
                        kind=f"pair1x1 {y.shape[3]}->{c3.cout}->{c1.cout} @{h}x{w}",
                        nbytes=2 * (y.numel() + identity.numel() + out1.numel() + out2.numel() + c3.w.numel() +
                                    c1.w.numel()))
                    x, pre = out1, out2
                else:
                    x, pre = add_conv(convs[2], y, residual=identity, relu=True), None
            else:
                y = add_conv(convs[0], x)
                x = add_conv(convs[1], y, residual=identity, relu=True)
            if bi == len(self.layers[li]) - 1 and li + 2 in self.outputs:
                feats.append(x)
        return feats


class _MobileNetV2Backbone:
    """MobileNetV2 under `backbone.` (reference nets/mobilenet.py:40-212, torchvision InvertedResidual:
    [1x1 expand + ReLU6] -> 3x3 depthwise + ReLU6 -> 1x1 linear project [+ residual]).
    Every activation tensor is stored with its channels zero-padded to a multiple of 32
    (16, 24 -> 32; 144 -> 160) so the 64-byte K-block / 32-column epilogue can address it."""
    SETTINGS = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2],
                [6, 320, 1, 1]]

    def _build_backbone(self, sd, feature_layer):
        dev = self.device
        self.outputs = [l for l in feature_layer[0] if isinstance(l, int)]
        self.stem = _Conv(sd["backbone.conv1.0.weight"], _bn(sd, "backbone.conv1.1"), None, 2, 1, 2, dev,
                          stem=True)
        self.layers = []
        inp = 32
        for j, (t, c, n, s_) in enumerate(self.SETTINGS, start=1):
            if j > max(self.outputs):
                break
            blocks = []
            for i in range(n):
                p = f"backbone.layer{j}.{i}.conv"
                stride = s_ if i == 0 else 1
                hid = inp * t
                blk = {"res": stride == 1 and inp == c}
                k = 0
                if t != 1:
                    blk["expand"] = _Conv(sd[f"{p}.0.0.weight"], _bn(sd, f"{p}.0.1"), None, 1, 0, 2, dev,
                                          cin_pad=_cpad(inp), cout_pad=_cpad(hid))
                    k = 1
                blk["dw"] = _DwConv(sd[f"{p}.{k}.0.weight"], _bn(sd, f"{p}.{k}.1"), stride, 2, dev,
                                    c_pad=_cpad(hid))
                blk["project"] = _Conv(sd[f"{p}.{k + 1}.weight"], _bn(sd, f"{p}.{k + 2}"), None, 1, 0, 0, dev,
                                       cin_pad=_cpad(hid), cout_pad=_cpad(c))
                blocks.append(blk)
                inp = c
            self.layers.append(blocks)

    def _plan_backbone(self, packed, H, W, steps, buf, add_conv, add_dw):
        x = add_conv(self.stem, packed, Ho=H // 2, Wo=W // 2, x_kind=1, x_width=W // 2)
        feats = []
        # [r2] one launch per inverted residual (conv_mbconv.cu): the expanded tensor stays on the SM.  Per block the
        # plan keeps whichever is faster on this GPU — the fused launch or expand / depthwise / project — measured
        # once per shape when the plan is built (both give bit-identical results, so the choice never changes an
        # output).  SSDSB_NO_MBFUSE=1: never fuse (A/B runs, the bit-exact reference of the bench self-check);
        # SSDSB_MBFUSE=1: always fuse where the kernel has a configuration.
        mode = "never" if os.environ.get("SSDSB_NO_MBFUSE", "0") == "1" else (
            "always" if os.environ.get("SSDSB_MBFUSE", "0") == "1" else "auto")

        def mb_args(blk, x):
            dw, pr, ex = blk["dw"], blk["project"], blk.get("expand")
            ew, eb = (ex.w, ex.bias) if ex is not None else (None, None)
            relu = (ex.relu if ex is not None else 0, dw.relu, pr.relu)
            return (x, ew, eb, dw.w, dw.bias, pr.w, pr.bias, dw.stride, blk["res"], relu)

        def fused_is_faster(blk, x):
            if mode != "auto":
                return mode == "always"
            n, h, w, cin = x.shape
            dw, pr, ex = blk["dw"], blk["project"], blk.get("expand")
            key = (n, h, w, cin, dw.c, pr.cout, dw.stride, blk["res"], ex is not None, x.device.index)
            if key not in _MB_CHOICE:
                ho, wo = (h - 1) // dw.stride + 1, (w - 1) // dw.stride + 1
                xs = torch.zeros_like(x)
                hb = buf(n, h, w, dw.c) if ex is not None else None
                db, y = buf(n, ho, wo, dw.c), buf(n, ho, wo, pr.cout)
                args = mb_args(blk, xs)

                def separate():
                    t = K.conv2d(xs, ex.w, ex.bias, 1, 1, 1, 0, ex.relu, out=hb) if ex is not None else xs
                    K.dwconv3x3(t, dw.w, dw.bias, dw.stride, dw.relu, out=db)
                    K.conv2d(db, pr.w, pr.bias, 1, 1, 1, 0, pr.relu, residual=xs if blk["res"] else None, out=y)

                try:
                    t_f = _graph_time(lambda: K.mbconv(*args, out=y))
                    t_s = _graph_time(separate)
                    _MB_CHOICE[key] = t_f < t_s
                except NotImplementedError:
                    _MB_CHOICE[key] = False
                del xs, hb, db, y
            return _MB_CHOICE[key]

        def add_mb(blk, x):
            n, h, w, cin = x.shape
            dw, pr, ex = blk["dw"], blk["project"], blk.get("expand")
            ho, wo = (h - 1) // dw.stride + 1, (w - 1) // dw.stride + 1
            y = buf(n, ho, wo, pr.cout)
            args = mb_args(blk, x)
            nfl = ((ex.flops_per_pixel * n * h * w) if ex is not None else 0) + \
                (dw.flops_per_pixel + pr.flops_per_pixel) * n * ho * wo
            # "bytes" stays the UN-FUSED algorithmic figure of SURVEY 8d (what expand / depthwise / project move as
            # three launches), so the reported roofline fraction is comparable across plans; the fused launch's own
            # floor (block input + output) is kept next to it
            wts = dw.w.numel() + pr.w.numel() + (ex.w.numel() if ex is not None else 0)
            hid_in, hid_out = n * h * w * dw.c, n * ho * wo * dw.c
            nby = 2 * (x.numel() * (2 if blk["res"] else 1) + (2 * hid_in if ex is not None else 0) + 2 * hid_out +
                       y.numel() + wts)
            self._add_raw(lambda: K.mbconv(*args, out=y), nfl,
                          kind=f"mbconv s{dw.stride} {cin}->{dw.c}->{pr.cout} @{ho}x{wo}", nbytes=nby)
            steps_info = self._last_info()
            steps_info["bytes_fused"] = 2 * (x.numel() * (2 if blk["res"] else 1) + y.numel() + wts)
            return y

        fuse = mode != "never"
        for j, blocks in enumerate(self.layers, start=1):
            for blk in blocks:
                if fuse and blk["project"].cout <= 256 and fused_is_faster(blk, x):
                    x = add_mb(blk, x)
                else:
                    y = add_conv(blk["expand"], x) if "expand" in blk else x
                    y = add_dw(blk["dw"], y)
                    x = add_conv(blk["project"], y, residual=x if blk["res"] else None)
            if j in self.outputs:
                feats.append(x)
        return feats


class _RegNetXBackbone:
    """RegNetX under `backbone.` (reference nets/regnet.py:28-282): SimpleStemIN 3x3/s2 (3->32), stages of
    ResBottleneckBlocks: a 1x1 -> b 3x3 grouped (group width 48, stride) -> c 1x1, + identity or 1x1/s
    projection, ReLU.  Channel counts are zero-padded to a multiple of 96 (= lcm(48, 32): two groups per
    block-diagonal chunk of the chunked igemm); padded channels stay exactly 0."""
    GROUP_W, CHUNK = 48, 96

    def _build_backbone(self, sd, feature_layer):
        dev = self.device
        self.outputs = [l for l in feature_layer[0] if isinstance(l, int)]
        self.stem = _Conv(sd["backbone.stem.conv.weight"], _bn(sd, "backbone.stem.bn"), None, 2, 1, 1, dev,
                          stem=True)
        self.layers = []
        pad = lambda c: _cpad(c, self.CHUNK)
        cin_pad = 32
        for si in range(1, 5):
            if si > max(self.outputs):
                break
            blocks = []
            bi = 1
            while f"backbone.s{si}.b{bi}.f.a.weight" in sd:
                p = f"backbone.s{si}.b{bi}"
                stride = 2 if bi == 1 else 1
                w_b = sd[p + ".f.a.weight"].shape[0]
                w_out = sd[p + ".f.c.weight"].shape[0]
                if sd[p + ".f.b.weight"].shape[1] != self.GROUP_W:
                    raise NotImplementedError("only group width 48 (RegNetX-3.2GF family) is packed here")
                blk = {
                    "a": _Conv(sd[p + ".f.a.weight"], _bn(sd, p + ".f.a_bn"), None, 1, 0, 1, dev,
                               cin_pad=cin_pad, cout_pad=pad(w_b)),
                    "b_w": K.pack_grouped_weight(K.fold_bn(sd[p + ".f.b.weight"], _bn(sd, p + ".f.b_bn"), None)[0],
                                                 self.CHUNK, pad(w_b)).to(dev),
                    "b_stride": stride,
                    "c": _Conv(sd[p + ".f.c.weight"], _bn(sd, p + ".f.c_bn"), None, 1, 0, 1, dev,
                               cin_pad=pad(w_b), cout_pad=pad(w_out)),
                }
                bb = K.fold_bn(sd[p + ".f.b.weight"], _bn(sd, p + ".f.b_bn"), None)[1]
                blk["b_bias"] = torch.cat([bb, torch.zeros(pad(w_b) - w_b)], 0).contiguous().to(dev)
                blk["b_flops"] = 2 * 9 * self.GROUP_W * w_b
                if (p + ".proj.weight") in sd:
                    blk["proj"] = _Conv(sd[p + ".proj.weight"], _bn(sd, p + ".bn"), None, stride, 0, 0, dev,
                                        cin_pad=cin_pad, cout_pad=pad(w_out))
                blocks.append(blk)
                cin_pad = pad(w_out)
                bi += 1
            self.layers.append(blocks)

    def _plan_backbone(self, packed, H, W, steps, buf, add_conv, add_dw):
        x = add_conv(self.stem, packed, Ho=H // 2, Wo=W // 2, x_kind=1, x_width=W // 2)
        feats = []
        for si, blocks in enumerate(self.layers, start=1):
            for blk in blocks:
                identity = add_conv(blk["proj"], x) if "proj" in blk else x
                y = add_conv(blk["a"], x)
                n, h, w, c = y.shape
                s_ = blk["b_stride"]
                ho, wo = (h - 1) // s_ + 1, (w - 1) // s_ + 1
                yb = buf(n, ho, wo, c)
                self._add_raw(lambda y=y, blk=blk, yb=yb, s_=s_: K.conv2d(
                    y, blk["b_w"], blk["b_bias"], 3, 3, s_, 1, True, out=yb, chunk=self.CHUNK),
                    blk["b_flops"] * n * ho * wo)
                x = add_conv(blk["c"], yb, residual=identity, relu=True)
            if si in self.outputs:
                feats.append(x)
        return feats


class _SSDNeck:
    """SSD extras (ConvBNReLUx2, layers_parser.py:17-20) + per-level fused loc/conf heads
    (ssd.py:42-104).  Input channel counts may be padded (see _MobileNetV2Backbone)."""

    def _build_neck(self, sd, feature_layer):
        dev = self.device
        # extras: ConvBNReLUx2 (layers_parser.py:17-20)
        self.extras = []
        ei = 0
        for layer in feature_layer[0]:
            if isinstance(layer, int):
                continue
            if layer not in ("Conv:S", "Conv"):
                raise NotImplementedError(f"extra layer {layer!r} is out of scope (SURVEY 2, rows 9-11)")
            stride = 2 if layer == "Conv:S" else 1
            p = f"extras.{ei}"
            self.extras.append([
                _Conv(sd[p + ".0.weight"], _bn(sd, p + ".1"), None, 1, 0, True, dev,
                      cin_pad=_cpad(sd[p + ".0.weight"].shape[1])),
                _Conv(sd[p + ".3.weight"], _bn(sd, p + ".4"), None, stride, 1, True, dev),
            ])
            ei += 1
        # heads: loc + conf fused per level (ssd.py:100-103)
        self.heads = []
        for l, nb in enumerate(self.number_box):
            w = torch.cat([sd[f"loc.{l}.weight"], sd[f"conf.{l}.weight"]], 0)
            b = torch.cat([sd[f"loc.{l}.bias"], sd[f"conf.{l}.bias"]], 0)
            h = _Conv(w, None, b, 1, 1, False, dev, cin_pad=_cpad(w.shape[1]))
            h.n_loc = nb * 4
            self.heads.append(h)

    def _plan_neck(self, feats, steps, buf, add_conv, add_head):
        feats = list(feats)
        n_back = len(feats)
        locs, confs = [None] * len(self.heads), [None] * len(self.heads)
        # the extras chain and the heads of its (small) levels run on side branches next to each other; the
        # heads of the backbone levels follow on main (see _Steps)
        chain = steps.fork() if self.extras else 0
        for j, ex in enumerate(self.extras):           # ssd.py:61-64
            with steps.on(chain):
                y = feats[-1]
                for cv in ex:
                    y = add_conv(cv, y)
                feats.append(y)
            with steps.on(steps.fork(after=chain)):
                locs[n_back + j], confs[n_back + j] = add_head(y, self.heads[n_back + j])
        for l in range(n_back):                        # ssd.py:67-70
            locs[l], confs[l] = add_head(feats[l], self.heads[l])
        return locs, confs


class _FPNNeck:
    """B200 engine for SSDFPN / RetinaNet over ResNet (reference cfg: SSDS='SSDFPN'; fpn.py:58-146).

    transforms: 1x1 lateral convs with bias (no BN/ReLU); top-down nearest-2x upsample + add
    (fpn.py:80-87); extras: ConvBNReLU 3x3 per pyramid level, 'Conv:S' levels are stride-2 convs on the
    raw last backbone map / the previous extra (fpn.py:89-95); loc and conf are SHARED towers of
    4 x ConvBNReLU(256,256,3) + 3x3 conv (fpn.py:10-18).  The first tower layer of loc and conf reads
    the same input, so the two are fused into one 256->512 conv; the later layers read their own
    256-channel half through the channel stride."""

    def _build_neck(self, sd, feature_layer):
        dev = self.device
        n_back = len(self.outputs)
        cpad = getattr(self, "CHUNK", 32)       # the backbone's channel padding (RegNet: 96)
        self.transforms = [_Conv(sd[f"transforms.{i}.weight"], None, sd[f"transforms.{i}.bias"], 1, 0,
                                 False, dev, cin_pad=_cpad(sd[f"transforms.{i}.weight"].shape[1], cpad))
                           for i in range(n_back)]
        self.extras = []
        for i, layer in enumerate(feature_layer[0]):
            stride = 1 if isinstance(layer, int) else 2
            if not isinstance(layer, int) and layer != "Conv:S":
                raise ValueError(layer + " does not support by SSDFPN")       # fpn.py:144
            w = sd[f"extras.{i}.0.weight"]
            self.extras.append(_Conv(w, _bn(sd, f"extras.{i}.1"), None, stride, 1, True, dev,
                                     cin_pad=_cpad(w.shape[1], cpad) if i == n_back else None))
        nb = self.number_box[0]
        if any(b != nb for b in self.number_box):
            raise ValueError("For SSDFPN module, the number of box have to be same in every layer")
        # tower layer 0 of loc and conf fused (same input)
        w0 = torch.cat([sd["loc.0.0.weight"], sd["conf.0.0.weight"]], 0)
        bn0 = tuple(torch.cat([a, b], 0) if isinstance(a, torch.Tensor) else a
                    for a, b in zip(_bn(sd, "loc.0.1"), _bn(sd, "conf.0.1")))
        self.tower0 = _Conv(w0, bn0, None, 1, 1, True, dev)
        self.tower = {t: [_Conv(sd[f"{t}.{j}.0.weight"], _bn(sd, f"{t}.{j}.1"), None, 1, 1, True, dev)
                          for j in (1, 2, 3)] for t in ("loc", "conf")}
        self.head_loc = _Conv(sd["loc.4.weight"], None, sd["loc.4.bias"], 1, 1, False, dev)
        self.head_loc.n_loc = self.head_loc.cout                 # every channel goes to `loc`
        self.head_conf = _Conv(sd["conf.4.weight"], None, sd["conf.4.bias"], 1, 1, False, dev)
        self.head_conf.n_loc = 0                                 # every channel goes to `conf`

    def _plan_pyramid(self, feats, steps, buf, add_conv):
        n_back = len(feats)
        pyr = [None] * n_back
        for i in range(n_back - 1, -1, -1):            # fpn.py:79-87
            lat = add_conv(self.transforms[i], feats[i])
            if i != n_back - 1:
                coarse = pyr[i + 1]
                steps.append(lambda c=coarse, f=lat: K.upsample2x_add(c, f))
            pyr[i] = lat
        return pyr

    def _plan_neck(self, feats, steps, buf, add_conv, add_head):
        n_back = len(feats)
        raw_top = feats[-1]
        pyr = self._plan_pyramid(feats, steps, buf, add_conv)
        # every pyramid level is an independent chain (extra -> fused tower layer 0 -> loc tower | conf tower ->
        # heads): level 0 stays on main, the others run on their own branches, the conf tower of a level on a
        # sub-branch of it (see _Steps) — the 5x5 ... 40x40 levels cannot fill 148 SMs one launch at a time
        # Recording order matters: a branch depends on what its parent had recorded when the branch's first step
        # was recorded, so the branches are recorded first and the main (level-0) chain last.
        L = len(self.extras)
        locs, confs = [None] * L, [None] * L
        dev = self.device
        level_in = {}                                   # level -> (branch, extra output) once recorded

        def record_level(i, br):
            with steps.on(br):
                if i < n_back:
                    f = add_conv(self.extras[i], pyr[i])                       # fpn.py:89-95
                elif i == n_back:
                    f = add_conv(self.extras[i], raw_top)
                else:
                    f = add_conv(self.extras[i], level_in[i - 1][1])
                level_in[i] = (br, f)
                t = add_conv(self.tower0, f)               # [.., 512]: loc half | conf half   (fpn.py:94-95)
                tl, tc = t[..., :256], t[..., 256:]
                n, fh, fw, _ = f.shape
                loc = torch.empty((n, self.head_loc.cout, fh, fw), dtype=torch.float32, device=dev)
                conf = torch.empty((n, self.head_conf.cout, fh, fw), dtype=torch.float32, device=dev)
                dummy_c = torch.empty((1,), dtype=torch.float32, device=dev)
                with steps.on(steps.fork(after=br) if br != 0 else 0):        # conf tower: sub-branch
                    for j in range(3):
                        tc = add_conv(self.tower["conf"][j], tc)
                    add_head(tc, self.head_conf, loc=dummy_c, conf=conf)
                for j in range(3):
                    tl = add_conv(self.tower["loc"][j], tl)
                add_head(tl, self.head_loc, loc=loc, conf=dummy_c)
            locs[i], confs[i] = loc, conf

        for i in range(1, L):
            br = steps.fork() if i <= n_back else steps.fork(after=level_in[i - 1][0])
            record_level(i, br)
        record_level(0, 0)
        return locs, confs


class _BiFPNNeck(_FPNNeck):
    """SSDBiFPN (EfficientDet-style, reference bifpn.py:10-142): 1x1 laterals (no top-down add), then
    `num_stack` BiFPNModules — relu-normalised scalar weights (host side, bifpn.py:35-38), weighted
    top-down (nearest up) and bottom-up (2x2 max-pool) fusions, each followed by ConvBNReLU 3x3 — then the
    same extras and shared towers as SSDFPN."""

    def _build_neck(self, sd, feature_layer):
        super()._build_neck(sd, feature_layer)
        dev = self.device
        self.bifpn = []
        s_ = 0
        while f"stack_bifpn.{s_}.w1" in sd:
            p = f"stack_bifpn.{s_}"
            w1 = torch.relu(sd[p + ".w1"].float())
            w1 = w1 / (w1.sum(0) + 1e-6)
            w2 = torch.relu(sd[p + ".w2"].float())
            w2 = w2 / (w2.sum(0) + 1e-6)
            levels = w1.shape[1]
            mod = {"w1": w1.tolist(), "w2": w2.tolist(), "levels": levels, "td": {}, "bu": {}}
            for i in range(levels - 1):
                mod["td"][i] = _Conv(sd[f"{p}.top-down-{i}.0.weight"], _bn(sd, f"{p}.top-down-{i}.1"), None,
                                     1, 1, True, dev)
                mod["bu"][i + 1] = _Conv(sd[f"{p}.bottom-up-{i + 1}.0.weight"],
                                         _bn(sd, f"{p}.bottom-up-{i + 1}.1"), None, 1, 1, True, dev)
            self.bifpn.append(mod)
            s_ += 1

    def _plan_pyramid(self, feats, steps, buf, add_conv):
        xx = [add_conv(self.transforms[i], f) for i, f in enumerate(feats)]      # bifpn.py:127-128
        for mod in self.bifpn:                                                    # bifpn.py:30-63
            L, w1, w2 = mod["levels"], mod["w1"], mod["w2"]
            xs = list(xx)
            for i in range(L - 1, 0, -1):
                fused = buf(*xx[i - 1].shape)
                steps.append(lambda a=xx[i - 1], b=xx[i], o=fused, wa=w1[0][i - 1], wb=w1[1][i - 1]:
                             K.bifpn_fuse(a, b, wa, wb, mode=0, out=o))
                xx[i - 1] = add_conv(mod["td"][i - 1], fused)
            for i in range(0, L - 2):
                fused = buf(*xx[i + 1].shape)
                steps.append(lambda a=xx[i + 1], b=xx[i], c=xs[i + 1], o=fused, wa=w2[0][i], wb=w2[1][i],
                             wc=w2[2][i]: K.bifpn_fuse(a, b, wa, wb, c=c, w2=wc, mode=1, out=o))
                xx[i + 1] = add_conv(mod["bu"][i + 1], fused)
            fused = buf(*xx[L - 1].shape)
            steps.append(lambda a=xx[L - 1], b=xx[L - 2], o=fused, wa=w1[0][L - 1], wb=w1[1][L - 1]:
                         K.bifpn_fuse(a, b, wa, wb, mode=1, out=o))
            xx[L - 1] = add_conv(mod["bu"][L - 1], fused)
        return xx


class _YOLOV3Neck:
    """YOLOv3 neck + heads (reference ssds/modeling/ssds/yolo.py:44-160; the model of the shipped
    experiments/cfgs/tests/test.yml).  Top-down from the last backbone level: extras[i] = ConvBNReLUx2 (1x1 then
    3x3), transforms[i] = ConvBNReLU 3x3 + nearest-2x upsample, concatenated AFTER the backbone feature
    (torch.cat((features[i], xx), 1), yolo.py:70-72) -> one `upsample2x_concat` launch; 'Conv:S' extras are
    stride-2 ConvBNReLU on the raw last backbone map / the previous extra (yolo.py:77-82).  Heads are per
    level Sequential(ConvBNReLU(c, c, 3), Conv2d(c, A*4 | A*C, 3)): the loc and conf towers read the same map,
    so their first convs run as ONE c -> 2c launch and the final convs read their half through the channel
    stride.  All activations are plain ReLU in this reference (basic_layers.py:27-57)."""

    def _build_neck(self, sd, feature_layer):
        dev = self.device
        layers, depths = feature_layer        # ([in, out] depth pairs only change channel counts: read off the weights)
        n_back = len(self.outputs)
        self.transforms = [_Conv(sd[f"transforms.{i}.0.weight"], _bn(sd, f"transforms.{i}.1"), None, 1, 1, True, dev)
                           for i in range(n_back - 1)]
        self.extras = []
        for i, layer in enumerate(layers):
            p = f"extras.{i}"
            if isinstance(layer, int):
                self.extras.append([_Conv(sd[p + ".0.weight"], _bn(sd, p + ".1"), None, 1, 0, True, dev),
                                    _Conv(sd[p + ".3.weight"], _bn(sd, p + ".4"), None, 1, 1, True, dev)])
            elif layer == "Conv:S":
                self.extras.append([_Conv(sd[p + ".0.weight"], _bn(sd, p + ".1"), None, 2, 1, True, dev)])
            else:
                raise ValueError(layer + " does not support by YOLO")            # yolo.py:136
        self.towers, self.head_loc, self.head_conf = [], [], []
        for l in range(len(layers)):
            w0 = torch.cat([sd[f"loc.{l}.0.0.weight"], sd[f"conf.{l}.0.0.weight"]], 0)
            bn0 = tuple(torch.cat([a, b], 0) if isinstance(a, torch.Tensor) else a
                        for a, b in zip(_bn(sd, f"loc.{l}.0.1"), _bn(sd, f"conf.{l}.0.1")))
            self.towers.append(_Conv(w0, bn0, None, 1, 1, True, dev))
            hl = _Conv(sd[f"loc.{l}.1.weight"], None, sd[f"loc.{l}.1.bias"], 1, 1, False, dev)
            hl.n_loc = hl.cout
            hc = _Conv(sd[f"conf.{l}.1.weight"], None, sd[f"conf.{l}.1.bias"], 1, 1, False, dev)
            hc.n_loc = 0
            self.head_loc.append(hl)
            self.head_conf.append(hc)

    def _plan_neck(self, feats, steps, buf, add_conv, add_head):
        n_back = len(feats)
        feats = list(feats)
        raw_top = feats[-1]
        L = len(self.extras)
        locs, confs = [None] * L, [None] * L
        dev = self.device

        def record_heads(l, f, br):
            with steps.on(br):
                t = add_conv(self.towers[l], f)                                   # [.., 2c]: loc half | conf half
                c = t.shape[3] // 2
                n, fh, fw, _ = f.shape
                loc = torch.empty((n, self.head_loc[l].cout, fh, fw), dtype=torch.float32, device=dev)
                conf = torch.empty((n, self.head_conf[l].cout, fh, fw), dtype=torch.float32, device=dev)
                dummy = torch.empty((1,), dtype=torch.float32, device=dev)
                add_head(t[..., :c], self.head_loc[l], loc=loc, conf=dummy)
                add_head(t[..., c:], self.head_conf[l], loc=dummy, conf=conf)
            locs[l], confs[l] = loc, conf

        # 'Conv:S' levels hang off the raw last backbone map (yolo.py:77-82): their chain runs on a branch of its
        # own, next to the top-down path; every level's heads fork off as soon as its map exists (see _Steps)
        prev, src = 0, raw_top
        for i in range(n_back, L):
            br = steps.fork(after=prev)
            with steps.on(br):
                src = add_conv(self.extras[i][0], src)
            record_heads(i, src, steps.fork(after=br))
            prev = br
        xx = None
        for i in range(n_back - 1, -1, -1):                                       # yolo.py:67-74
            if i != n_back - 1:
                t = add_conv(self.transforms[i], xx)
                n, h, w, cf = feats[i].shape
                cat = buf(n, h, w, cf + t.shape[3])
                steps.append(lambda f=feats[i], c=t, o=cat: K.upsample2x_concat(f, c, out=o))
                xx = cat
            else:
                xx = feats[i]
            for cv in self.extras[i]:
                xx = add_conv(cv, xx)
            feats[i] = xx
            record_heads(i, xx, steps.fork() if i != 0 else 0)
        return locs, confs


class _YOLOV4Neck(_YOLOV3Neck):
    """YOLOv4 neck + heads (reference ssds/modeling/ssds/yolo.py:161-392): per backbone level transforms.{i} =
    ConvBNReLU 3x3 (depth -> depth/2), the last one followed by the SPP block (max-pools of 5 / 9 / 13 concatenated
    after x: three cascaded 5x5 pools, each reading one channel slice of the 4c buffer and writing the next) and a
    second 3x3 conv; a stack of PANModules (top-down: 3x3 conv + nearest-2x upsample concatenated AFTER the finer
    level + ConvBNReLUx2; bottom-up: 3x3 stride-2 conv concatenated AFTER the coarser level + ConvBNReLUx2);
    'Conv:S' extras chained from the last PAN level; per-level heads as in YOLOv3.  Concatenations cost no launch
    except the upsampling one: every producer writes straight into its channel slice of the concatenated buffer
    (conv `out_cstride`), and the convs read channel slices through `x_cstride`."""

    def _build_neck(self, sd, feature_layer):
        dev = self.device
        layers = feature_layer[0]
        self.stacks = 1 if len(feature_layer) == 2 else feature_layer[2]
        if self.stacks != 1:
            raise NotImplementedError("YOLOV4 with more than one PAN stack is not on the tcgen05 conv stack yet")
        n = len(self.outputs)
        cbr = lambda p, q, stride=1, pad=1: _Conv(sd[p + ".weight"], _bn(sd, q), None, stride, pad, True, dev)
        self.transforms = []
        for i in range(n):
            if i == n - 1:
                self.transforms.append([cbr(f"transforms.{i}.0.0", f"transforms.{i}.0.1"),
                                        cbr(f"transforms.{i}.2.0", f"transforms.{i}.2.1")])
            else:
                self.transforms.append([cbr(f"transforms.{i}.0", f"transforms.{i}.1")])
        self.pan = []
        for st in range(self.stacks):
            td, bu = {}, {}
            for i in range(n - 1, 0, -1):
                p = f"fpn.{st}.top-down-{i - 1}"
                td[i] = (cbr(f"fpn.{st}.top-down-{i}-to-{i - 1}.0", f"fpn.{st}.top-down-{i}-to-{i - 1}.1"),
                         cbr(p + ".0", p + ".1", 1, 0), cbr(p + ".3", p + ".4"))
            for i in range(0, n - 1):
                p = f"fpn.{st}.bottom-up-{i + 1}"
                bu[i] = (cbr(f"fpn.{st}.bottom-up-{i}-to-{i + 1}.0", f"fpn.{st}.bottom-up-{i}-to-{i + 1}.1", 2, 1),
                         cbr(p + ".0", p + ".1", 1, 0), cbr(p + ".3", p + ".4"))
            self.pan.append((td, bu))
        self.extras = []
        j = 0
        for layer in layers:
            if isinstance(layer, int):
                continue
            if layer != "Conv:S":
                raise ValueError(layer + " does not support by YOLO")            # yolo.py:360
            self.extras.append(cbr(f"extras.{j}.0", f"extras.{j}.1", 2, 1))
            j += 1
        self.towers, self.head_loc, self.head_conf = [], [], []
        for l in range(len(layers)):
            w0 = torch.cat([sd[f"loc.{l}.0.0.weight"], sd[f"conf.{l}.0.0.weight"]], 0)
            bn0 = tuple(torch.cat([a, b], 0) if isinstance(a, torch.Tensor) else a
                        for a, b in zip(_bn(sd, f"loc.{l}.0.1"), _bn(sd, f"conf.{l}.0.1")))
            self.towers.append(_Conv(w0, bn0, None, 1, 1, True, dev))
            hl = _Conv(sd[f"loc.{l}.1.weight"], None, sd[f"loc.{l}.1.bias"], 1, 1, False, dev)
            hl.n_loc = hl.cout
            hc = _Conv(sd[f"conf.{l}.1.weight"], None, sd[f"conf.{l}.1.bias"], 1, 1, False, dev)
            hc.n_loc = 0
            self.head_loc.append(hl)
            self.head_conf.append(hc)

    def _plan_neck(self, feats, steps, buf, add_conv, add_head):
        n = len(feats)
        L = n + len(self.extras)
        locs, confs = [None] * L, [None] * L
        dev = self.device

        def record_heads(l, f, br):
            with steps.on(br):
                t = add_conv(self.towers[l], f)                                   # [.., 2c]: loc half | conf half
                c = t.shape[3] // 2
                nb, fh, fw, _ = f.shape
                loc = torch.empty((nb, self.head_loc[l].cout, fh, fw), dtype=torch.float32, device=dev)
                conf = torch.empty((nb, self.head_conf[l].cout, fh, fw), dtype=torch.float32, device=dev)
                dummy = torch.empty((1,), dtype=torch.float32, device=dev)
                add_head(t[..., :c], self.head_loc[l], loc=loc, conf=dummy)
                add_head(t[..., c:], self.head_conf[l], loc=dummy, conf=conf)
            locs[l], confs[l] = loc, conf

        # the concatenated input of every bottom-up ConvBNReLUx2 (level j >= 1): [.., xx[j] | stride-2 conv of xx[j-1]];
        # whoever produces xx[j] writes it straight into the first half
        def cat_for(t):
            nb, h, w, c = t.shape
            return buf(nb, h, w, 2 * c)

        xx = [None] * n
        for i in range(n):
            tr = self.transforms[i]
            nb, h, w, _ = feats[i].shape
            c = tr[0].cout
            if i == n - 1:
                spp = buf(nb, h, w, 4 * c)                                       # yolo.py:161-184: x | pool5 | pool9 | pool13
                add_conv(tr[0], feats[i], out=spp[..., :c])
                for k in range(3):
                    steps.append(lambda a=spp[..., k * c:(k + 1) * c], o=spp[..., (k + 1) * c:(k + 2) * c]:
                                 K.maxpool5x5s1(a, o))
                top_cat = buf(nb, h, w, 2 * c) if n > 1 else None
                xx[i] = add_conv(tr[1], spp, out=top_cat[..., :c] if top_cat is not None else None)
                cats = {i: top_cat}
            else:
                xx[i] = add_conv(tr[0], feats[i])
        for st, (td, bu) in enumerate(self.pan):
            if st > 0:
                cats = {}
            for i in range(n - 1, 0, -1):                                         # top-down (yolo.py:225-236)
                t = add_conv(td[i][0], xx[i])
                nb, h, w, cf = xx[i - 1].shape
                cat = buf(nb, h, w, cf + t.shape[3])
                steps.append(lambda f=xx[i - 1], cc=t, o=cat: K.upsample2x_concat(f, cc, out=o))
                y = add_conv(td[i][1], cat)
                if i - 1 >= 1:                                                     # becomes xx[i-1] of a bottom-up concat
                    cats[i - 1] = buf(nb, h, w, 2 * td[i][2].cout)
                    xx[i - 1] = add_conv(td[i][2], y, out=cats[i - 1][..., :td[i][2].cout])
                else:
                    xx[i - 1] = add_conv(td[i][2], y)
            for i in range(0, n - 1):                                             # bottom-up (yolo.py:238-246)
                cat = cats.get(i + 1)
                c = bu[i][0].cout
                if cat is None:                                                    # (only after a previous PAN stack)
                    nb, h, w, _ = xx[i + 1].shape
                    cat = buf(nb, h, w, 2 * c)
                    steps.append(lambda s_=xx[i + 1], o=cat[..., :c]: o.copy_(s_))
                add_conv(bu[i][0], xx[i], out=cat[..., c:])
                y = add_conv(bu[i][1], cat)
                xx[i + 1] = add_conv(bu[i][2], y)
        t = xx[-1]
        for e in self.extras:                                                      # yolo.py:303-306
            t = add_conv(e, t)
            xx.append(t)
        for l, f in enumerate(xx):
            record_heads(l, f, steps.fork() if l != len(xx) - 1 else 0)
        return locs, confs


class SSDResNet(_SSDNeck, _ResNetBackbone, _Engine):
    """SSD + ResNet18/34/50/101/152 (reference cfg SSDS='SSD', NETS='ResNet*')."""


class SSDFPNResNet(_FPNNeck, _ResNetBackbone, _Engine):
    """SSDFPN (RetinaNet) + ResNet (reference cfg SSDS='SSDFPN', NETS='ResNet*')."""


class SSDMobileNetV2(_SSDNeck, _MobileNetV2Backbone, _Engine):
    """SSD + MobileNetV2 (reference cfg SSDS='SSD', NETS='MobileNetV2'; BASELINE configs[0]/[2])."""


class SSDBiFPNResNet(_BiFPNNeck, _ResNetBackbone, _Engine):
    """SSDBiFPN + ResNet."""


class SSDBiFPNRegNetX(_BiFPNNeck, _RegNetXBackbone, _Engine):
    """SSDBiFPN + RegNetX (group width 48: RegNetX-3.2GF; BASELINE configs[4])."""


class SSDFPNRegNetX(_FPNNeck, _RegNetXBackbone, _Engine):
    """SSDFPN + RegNetX."""


class YOLOV3ResNet(_YOLOV3Neck, _ResNetBackbone, _Engine):
    """YOLOV3 + ResNet18/34/50/... (reference cfg SSDS='YOLOV3'; experiments/cfgs/tests/test.yml is ResNet18 @320)."""


class YOLOV4ResNet(_YOLOV4Neck, _ResNetBackbone, _Engine):
    """YOLOV4 + ResNet (reference cfg SSDS='YOLOV4': SPP + PAN neck, yolo.py:161-392)."""


ENGINES = {("YOLOV3", "ResNet"): YOLOV3ResNet, ("YOLOV4", "ResNet"): YOLOV4ResNet, ("SSD", "ResNet"): SSDResNet, ("SSDFPN", "ResNet"): SSDFPNResNet,
           ("SSD", "MobileNetV2"): SSDMobileNetV2, ("SSDBIFPN", "ResNet"): SSDBiFPNResNet,
           ("SSDBIFPN", "RegNetX032"): SSDBiFPNRegNetX, ("SSDFPN", "RegNetX032"): SSDFPNRegNetX}


def engine_for(ssds, nets):
    key = (ssds.upper(), "ResNet" if nets.startswith("ResNet") else nets)
    if key not in ENGINES:
        raise NotImplementedError(f"SSDS={ssds!r} / NETS={nets!r} is not on the tcgen05 conv stack yet "
                                  f"(have: {sorted(ENGINES)})")
    return ENGINES[key]


def create_anchors(model_cfg, model, image_size):
    """reference model_builder.py:30-56: strides = W_in // W_feat of each conf map, then
    generate_anchors per level.  The feature sizes come from the plan (no dummy forward needed)."""
    x = torch.zeros((1, 3, image_size[0], image_size[1]), device=model.device)
    plan = model.plan_for(x)
    strides = [image_size[1] // c.shape[-1] for c in plan["conf"]]
    ratios, scales = configure_ratio_scale(len(strides), model_cfg["ASPECT_RATIOS"], model_cfg["SIZES"])
    return OrderedDict((strides[i], generate_anchors(strides[i], ratios[i], scales[i], model.device))
                       for i in range(len(strides)))


def number_box_from_cfg(model_cfg):
    """reference model_builder.py:14-15."""
    ratios, scales = configure_ratio_scale(len(model_cfg["SIZES"]), model_cfg["ASPECT_RATIOS"],
                                           [s if isinstance(s, list) else s for s in model_cfg["SIZES"]])
    return [len(r) * len(s) for r, s in zip(ratios, scales)]
