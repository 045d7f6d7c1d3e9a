"""Device-timed throughput of the other BASELINE.json configurations on ONE B200 (per-rank shard of the
8-GPU configs), printed as one JSON line per config.  Not the driver's bench (that is bench.py, cfg 2);
this records that the other configs run end to end on the same kernels and how fast.

    python tools/bench_configs.py [cfg3] [cfg4] [cfg5] [cfg5stress]
"""
import json
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import synth                      # noqa: E402
from ssds_pytorch_b200.ssds import SSDDetector           # noqa: E402
from ssds_pytorch_b200.pipeline import multibox_cls_loss_step, detection_loss_step  # noqa: E402
import ssds_pytorch_b200 as S                            # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def detector(ssds, nets, fl, image, ncls, sizes, ratios, pp=None):
    L = len(fl[0])
    cfg = {"MODEL": {"SSDS": ssds, "NETS": nets, "IMAGE_SIZE": image, "NUM_CLASSES": ncls, "FEATURE_LAYER": fl,
                     "SIZES": [list(sizes)] * L, "ASPECT_RATIOS": [list(ratios)] * L},
           "DATASET": {"PREPROC": {"MEAN": 0, "STD": 255}}}
    if pp:
        cfg["POST_PROCESS"] = pp
    nb = len(sizes) * len(ratios)
    sd = synth.synthetic_state_dict(nets, fl, [nb] * L, ncls, seed=0, style="init", ssds=ssds)
    return SSDDetector(cfg, sd), nb


def cfg3():
    """MobileNetV2-SSD 300x300, batch 64 = one rank's shard of the 512-image / 8-GPU config."""
    fl = [[5, 7, "Conv:S", "Conv:S", "Conv:S", "Conv:S"], [96, 320, 512, 256, 256, 128]]
    det, _ = detector("SSD", "MobileNetV2", fl, [300, 300], 80, [2.0, 2.828], [1, 2, 0.5])
    B = 64
    x = torch.randint(0, 256, (B, 300, 300, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234)).cuda()
    ms = timed(lambda: det.detect_device(x))
    plan = det.model.plan_for(x)
    return {"config": "cfg3 MobileNetV2-SSD 300x300 bf16, 64 images (per-GPU shard of 512/8)", "ms_per_step": ms,
            "images_per_s_per_gpu": B / ms * 1e3, "conv_gflop_per_image": plan["flops"] / B / 1e9,
            "launches": plan["launches"] + 3}


def cfg4(default_losses=False):
    """SSDFPN-ResNet50 640x640 training-step forward: conv stack (train mode: logits) + match +
    MultiBoxLoss hard-negative mining, 16 images = one rank's shard of 128/8.  default_losses: the
    reference's configured defaults instead (FocalLoss + SmoothL1Loss, config.py:151-152), cls + loc."""
    fl = [[3, 4, 5, "Conv:S", "Conv:S"], [512, 1024, 2048, 2048, 256]]
    det, nb = detector("SSDFPN", "ResNet50", fl, [640, 640], 80, [4.0, 5.04, 6.35], [1, 2, 0.5])
    B = 16
    det.model.train()
    g = torch.Generator().manual_seed(4321)
    x = torch.randint(0, 256, (B, 640, 640, 3), dtype=torch.uint8, generator=g).cuda()
    tg = torch.full((B, 32, 5), -1.0)
    for b in range(B):
        n = int(torch.randint(1, 33, (1,), generator=g))
        tg[b, :n, :2] = torch.rand((n, 2), generator=g) * 480
        tg[b, :n, 2:4] = torch.rand((n, 2), generator=g) * 240 + 16
        tg[b, :n, 4] = torch.randint(0, 80, (n,), generator=g).float()
    tg = tg.cuda()
    out = {}

    def step():
        loc, conf = det.model(x, use_graph=True)
        if default_losses:
            out["loss"], out["loc_loss"], _ = detection_loss_step(loc, conf, tg, det.anchors, 80)
        else:
            out["loss"], _ = multibox_cls_loss_step(conf, tg, det.anchors, 80)

    ms = timed(step)
    ms_fwd = timed(lambda: det.model(x, use_graph=True))
    anchors_per_img = sum(c.shape[1] // 80 * c.shape[2] * c.shape[3] for c in det.model.plan_for(x)["conf"])
    what = "match + FocalLoss + SmoothL1Loss" if default_losses else "match + MultiBoxLoss"
    extra = {"loc_loss": float(out["loc_loss"])} if default_losses else {}
    return {**extra, "config": f"cfg4 SSDFPN-ResNet50 640x640: forward (logits) + {what}, 16 images "
                      "(per-GPU shard of 128/8)", "ms_per_step": ms, "ms_forward": ms_fwd,
            "ms_match_plus_loss": ms - ms_fwd, "images_per_s_per_gpu": B / ms * 1e3,
            "anchors_per_image": anchors_per_img, "cls_loss": float(out["loss"])}


def cfg5(stress=False):
    """SSDBiFPN-RegNetX032 1280x1280, 4 images = one rank's shard of 32/8; stress: 20 000 candidates
    per level into NMS (N = 100 000)."""
    fl = [[2, 3, 4, "Conv:S", "Conv:S"], [192, 432, 1008, 1008, 256]]
    pp = {"MAX_DETECTIONS_PER_LEVEL": 20000} if stress else None
    det, _ = detector("SSDBiFPN", "RegNetX032", fl, [1280, 1280], 80, [4.0], [1, 2, 0.5], pp)
    B = 4
    x = torch.randint(0, 256, (B, 1280, 1280, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234)).cuda()
    ms = timed(lambda: det.detect_device(x), iters=5, warm=2)
    ms_fwd = timed(lambda: det.model(x, use_graph=True), iters=5, warm=2)
    plan = det.model.plan_for(x)
    return {"config": "cfg5 SSDBiFPN-RegNetX032 1280x1280 bf16, 4 images (per-GPU shard of 32/8)" +
                      (", 20 000 candidates/level (N=100 000 into NMS)" if stress else ""),
            "ms_per_step": ms, "ms_forward": ms_fwd, "ms_decode_plus_nms": ms - ms_fwd,
            "images_per_s_per_gpu": B / ms * 1e3, "conv_gflop_per_image": plan["flops"] / B / 1e9}


def nms_stress():
    """Kernel-only NMS stress of SURVEY 8d cfg 5: N = 100 000 uniform candidates, D = 100, 80 classes."""
    rng = np.random.default_rng(99)
    B, N = 4, 100_000
    scores = torch.from_numpy(rng.uniform(0.01, 1, (B, N)).astype(np.float32)).cuda()
    xy = rng.uniform(0, 1180, (B, N, 2))
    wh = rng.uniform(8, 400, (B, N, 2))
    boxes = torch.from_numpy(np.clip(np.concatenate([xy, xy + wh], -1), 0, 1279).astype(np.float32)).cuda()
    classes = torch.from_numpy(rng.integers(0, 80, (B, N)).astype(np.float32)).cuda()
    ms = timed(lambda: S.nms(scores, boxes, classes, 0.6, 100, True))
    return {"config": "NMS stress kernel-only: N=100 000, D=100, 80 classes, 4 images", "ms": ms,
            "us_per_image": ms / B * 1e3, "algorithmic_GB_per_s": (24 * N + 24 * 100) * B / ms / 1e6}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5", "cfg5stress", "nms_stress"]
    fns = {"cfg3": cfg3, "cfg4": cfg4, "cfg4focal": lambda: cfg4(True), "cfg5": cfg5, "cfg5stress": lambda: cfg5(True), "nms_stress": nms_stress}
    for w in which:
        t0 = time.time()
        r = fns[w]()
        r["wall_s_incl_setup"] = round(time.time() - t0, 1)
        print(json.dumps(r))
        torch.cuda.empty_cache()
