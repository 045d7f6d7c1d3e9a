"""Run every NON-conv kernel of the hot path once (after a warm-up) at the BASELINE geometry it is benchmarked
on, so that ONE `ncu --set full` capture of this script has a warm launch of each:

    python tools/profile_misc.py [loss] [match] [decode] [decode_large] [nms] [dw] [layout]

loss          loss_step_kernel (cfg 4: 5 levels, A=9, C=80, 76 725 anchors/img, 16 images, T=32) MultiBox + SmoothL1
match         match_kernel + mbl_ce/mbl_select/mbl_sum + focal_kernel + loc_loss_kernel + masked_sum (cfg-4 level 0)
decode        decode_select + decode_finalize (cfg 2: 6 levels, 64 images)
decode_large  dl_pass<1,2,3> + dl_finalize (cfg-5 stress: 5 levels, A=3, 4 images, 20 000 per level)
nms           nms_kernel at N=1800 x 64 images and N=100 000 x 4 images
dw            dwconv3x3_kernel at the MobileNetV2-SSD 300x300 B=64 block shapes
layout        pack_image_s2d, maxpool3x3s2, upsample2x_add, upsample2x_concat, bifpn_fuse
Without a profiler it prints CUDA-event timings of each section.
"""
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import ssds_pytorch_b200 as S                              # noqa: E402
from ssds_pytorch_b200 import conv as K                   # noqa: E402
from ssds_pytorch_b200 import pipeline as P               # noqa: E402
from ssds_pytorch_b200 import synth                       # noqa: E402


def timed(name, fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / iters * 1e3:.1f} us")


def timed_graph(name, fn, iters=10):
    """like timed(), but the `iters` launches are replayed from one CUDA graph: no per-launch host time in the
    number (eager ctypes launches cost ~15 us each, more than the small kernels themselves)"""
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e3
    print(f"{name}: {t:.1f} us")
    return t


def levels_cfg4(B):
    g = torch.Generator(device="cuda").manual_seed(1)
    lv = [(8, 80), (16, 40), (32, 20), (64, 10), (128, 5)]
    anchors = OrderedDict((s, S.generate_anchors(s, [1, 2, 0.5], [4.0, 5.04, 6.35])) for s, _ in lv)
    conf = [torch.randn((B, 9 * 80, hw, hw), generator=g, device="cuda") - 4.6 for _, hw in lv]
    loc = [torch.randn((B, 9 * 4, hw, hw), generator=g, device="cuda") * 0.5 for _, hw in lv]
    tg = synth.synthetic_targets(B, seed=4321).cuda()
    return lv, anchors, conf, loc, tg


def run_loss():
    B = 16
    lv, anchors, conf, loc, tg = levels_cfg4(B)
    out = torch.empty(3, device="cuda")
    timed("loss_step (MultiBox + SmoothL1), cfg4 16 img", lambda: P.fused_loss_step(loc, conf, tg, anchors, 80, "MultiBoxLoss",
                                                                                 "SmoothL1Loss", out=out))
    timed("loss_step (MultiBox only), cfg4 16 img", lambda: P.fused_loss_step(loc, conf, tg, anchors, 80, "MultiBoxLoss", None, out=out))
    timed("loss_step (Focal + SmoothL1), cfg4 16 img", lambda: P.fused_loss_step(loc, conf, tg, anchors, 80, "FocalLoss",
                                                                               "SmoothL1Loss", out=out))
    print("scalars", out.tolist())


def run_match():
    B = 16
    lv, anchors, conf, loc, tg = levels_cfg4(B)
    crit, foc, sl1 = S.MultiBoxLoss(3), S.FocalLoss(), S.SmoothL1Loss()
    s, hw = lv[0]
    c5 = conf[0].view(B, 9, 80, hw, hw)
    l5 = loc[0].view(B, 9, 4, hw, hw)
    _, bt, dep = S.extract_targets(tg, anchors, 80, s, (hw, hw), [0.5, 0.4], with_cls_target=False)
    timed("match_kernel cfg4 L0", lambda: S.extract_targets(tg, anchors, 80, s, (hw, hw), [0.5, 0.4], with_cls_target=False))
    timed("mbl sum cfg4 L0", lambda: crit.forward_sum(c5, dep))
    timed("focal sum cfg4 L0", lambda: foc.forward_sum(c5, dep))
    timed("smoothl1 sum cfg4 L0", lambda: sl1.forward_sum(l5, bt, dep))


def score_levels(B, lv, A, C, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    conf = [torch.sigmoid(torch.randn((B, A * C, h, w), generator=g, device="cuda") * 0.3 - 4.6) for _, h, w in lv]
    loc = [torch.randn((B, A * 4, h, w), generator=g, device="cuda") * 0.3 for _, h, w in lv]
    return conf, loc


def run_decode():
    B = 64
    lv = [(8, 64, 64), (16, 32, 32), (32, 16, 16), (64, 8, 8), (128, 4, 4), (256, 2, 2)]
    conf, loc = score_levels(B, lv, 6, 80, 2)
    items = [(s, S.generate_anchors(s, [1, 2, 0.5], [2.0, 2.828])) for s, _, _ in lv]
    timed("decode cfg2 64 img (select + finalize)", lambda: S.decode_levels(conf, loc, items, 0.01, 300, True))
    return S.decode_levels(conf, loc, items, 0.01, 300, True)


def run_decode_large():
    B = 4
    lv = [(8, 160, 160), (16, 80, 80), (32, 40, 40), (64, 20, 20), (128, 10, 10)]
    conf, loc = score_levels(B, lv, 3, 80, 3)
    items = [(s, S.generate_anchors(s, [1, 2, 0.5], [4.0])) for s, _, _ in lv]
    timed("decode cfg5stress 4 img, 20 000/level (3 passes + finalize)",
          lambda: S.decode_levels(conf, loc, items, 0.01, 20000, True))
    return S.decode_levels(conf, loc, items, 0.01, 20000, True)


def run_nms(dec_small=None, dec_large=None):
    if dec_small is None:
        dec_small = run_decode()
    timed("nms N=1800 x 64 img", lambda: S.nms(*dec_small, 0.6, 100, True))
    if dec_large is None:
        dec_large = run_decode_large()
    timed("nms N=100 000 x 4 img", lambda: S.nms(*dec_large, 0.6, 100, True))


def run_dw():
    g = torch.Generator(device="cuda").manual_seed(4)
    for (h, c, stride) in [(150, 32, 1), (150, 96, 2), (75, 160, 1), (75, 160, 2), (38, 192, 1), (38, 192, 2),
                           (19, 384, 1), (19, 576, 1), (19, 576, 2), (10, 960, 1)]:
        x = torch.randn((64, h, h, c), generator=g, device="cuda").to(torch.bfloat16)
        w = K.pack_dw_weight(torch.randn((c, 1, 3, 3)) * 0.3).cuda()
        b = torch.zeros(c, device="cuda")
        ho = (h - 1) // stride + 1
        mb = 64 * (h * h + ho * ho) * c * 2 / 1e6
        timed(f"dwconv3x3 s{stride} {c} @{h} ({mb:.0f} MB)", lambda: K.dwconv3x3(x, w, b, stride, 2))


def run_pw():
    """the 1x1 convs of MobileNetV2-SSD 300x300, B=64 (channels padded to 32 as the plan stores them)"""
    g = torch.Generator(device="cuda").manual_seed(6)
    for (h, cin, cout, relu, res) in [(150, 32, 32, 0, False), (150, 32, 96, 2, False), (75, 96, 32, 0, False),
                                      (75, 32, 160, 2, False), (75, 160, 32, 0, True), (38, 160, 32, 0, False),
                                      (38, 32, 192, 2, False), (38, 192, 32, 0, True), (19, 192, 64, 0, False),
                                      (19, 64, 384, 2, False), (19, 384, 64, 0, True), (19, 384, 96, 0, False),
                                      (19, 96, 576, 2, False), (19, 576, 96, 0, True), (10, 576, 160, 0, False),
                                      (10, 160, 960, 2, False), (10, 960, 160, 0, True), (10, 960, 320, 0, False)]:
        x = torch.randn((64, h, h, cin), generator=g, device="cuda").to(torch.bfloat16)
        w = K.pack_weight(torch.randn((cout, cin, 1, 1)) * (1.0 / np.sqrt(cin))).cuda()
        b = torch.zeros(cout, device="cuda")
        r = torch.randn((64, h, h, cout), generator=g, device="cuda").to(torch.bfloat16) if res else None
        out = torch.empty((64, h, h, cout), dtype=torch.bfloat16, device="cuda")
        mb = 64 * h * h * (cin + cout * (2 if res else 1)) * 2 / 1e6
        timed(f"conv1x1 {cin}->{cout} @{h}{' +res' if res else ''} ({mb:.0f} MB)",
              lambda: K.conv2d(x, w, b, 1, 1, 1, 0, relu, residual=r, out=out))


def run_mb():
    """every inverted residual of MobileNetV2-SSD 300x300, B=64: one fused launch vs expand + depthwise + project"""
    g = torch.Generator(device="cuda").manual_seed(8)
    blocks = [(150, 32, 32, 32, 1, False), (150, 32, 96, 32, 2, False), (75, 32, 160, 32, 1, True),
              (75, 32, 160, 32, 2, False), (38, 32, 192, 32, 1, True), (38, 32, 192, 64, 2, False),
              (19, 64, 384, 64, 1, True), (19, 64, 384, 96, 1, False), (19, 96, 576, 96, 1, True),
              (19, 96, 576, 160, 2, False), (10, 160, 960, 160, 1, True)]
    tot_f = tot_s = 0.0
    for (h, cin, hid, cout, stride, res) in blocks:
        x = torch.randn((64, h, h, cin), generator=g, device="cuda").to(torch.bfloat16)
        has_e = not (hid == cin and stride == 1 and not res)
        we = K.pack_weight(torch.randn((hid, cin, 1, 1)) * (1.0 / np.sqrt(cin))).cuda() if has_e else None
        be = torch.zeros(hid, device="cuda") if has_e else None
        wd = K.pack_dw_weight(torch.randn((hid, 1, 3, 3)) * 0.3).cuda()
        bd = torch.zeros(hid, device="cuda")
        wp = K.pack_weight(torch.randn((cout, hid, 1, 1)) * (1.0 / np.sqrt(hid))).cuda()
        bp = torch.zeros(cout, device="cuda")
        ho = (h - 1) // stride + 1
        hb = torch.empty((64, h, h, hid), dtype=torch.bfloat16, device="cuda")
        db = torch.empty((64, ho, ho, hid), dtype=torch.bfloat16, device="cuda")
        y = torch.empty((64, ho, ho, cout), dtype=torch.bfloat16, device="cuda")

        def separate():
            hh = K.conv2d(x, we, be, 1, 1, 1, 0, 2, out=hb) if has_e else x
            K.dwconv3x3(hh, wd, bd, stride, 2, out=db)
            K.conv2d(db, wp, bp, 1, 1, 1, 0, 0, residual=x if res else None, out=y)

        tag = f"{cin}->{hid}->{cout} s{stride} @{h}{' +x' if res else ''}"
        tf = timed_graph(f"mbconv fused    {tag}", lambda: K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0), out=y))
        print("      ", K.mbconv_last_launch())
        ts = timed_graph(f"three launches  {tag}", separate)
        tot_f += tf
        tot_s += ts
    print(f"sum over the 11 distinct blocks: fused {tot_f:.0f} us, separate {tot_s:.0f} us")


def run_layout():
    g = torch.Generator(device="cuda").manual_seed(5)
    img = torch.randint(0, 256, (64, 512, 512, 3), dtype=torch.uint8, device="cuda")
    timed("pack_image_s2d 64x512x512", lambda: K.pack_image_s2d(img, 0.0, 255.0, padded=True))
    x = torch.randn((64, 256, 256, 64), generator=g, device="cuda").to(torch.bfloat16)
    timed("maxpool3x3s2 64x256x256x64", lambda: K.maxpool3x3s2(x))
    fine = torch.randn((16, 80, 80, 256), generator=g, device="cuda").to(torch.bfloat16)
    coarse = torch.randn((16, 40, 40, 256), generator=g, device="cuda").to(torch.bfloat16)
    timed("upsample2x_add 16x80x80x256", lambda: K.upsample2x_add(coarse, fine))
    timed("upsample2x_concat 16x80x80x(256+256)", lambda: K.upsample2x_concat(fine, coarse))
    a = torch.randn((4, 160, 160, 256), generator=g, device="cuda").to(torch.bfloat16)
    bb = torch.randn((4, 80, 80, 256), generator=g, device="cuda").to(torch.bfloat16)
    timed("bifpn_fuse up 4x160x160x256", lambda: K.bifpn_fuse(a, bb, 0.5, 0.5, mode=0))


if __name__ == "__main__":
    which = sys.argv[1:] or ["loss", "match", "decode", "decode_large", "nms", "dw", "pw", "layout"]
    small = large = None
    for w in which:
        if w == "loss":
            run_loss()
        elif w == "match":
            run_match()
        elif w == "decode":
            small = run_decode()
        elif w == "decode_large":
            large = run_decode_large()
        elif w == "nms":
            run_nms(small, large)
        elif w == "dw":
            run_dw()
        elif w == "pw":
            run_pw()
        elif w == "mb":
            run_mb()
        elif w == "layout":
            run_layout()
