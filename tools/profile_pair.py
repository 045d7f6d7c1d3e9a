"""conv3 -> next conv1 of the cfg-2 bottleneck stages (B=64): separate launches vs the fused
conv_pair launch, CUDA-event timed in a realistic sequence (c2-sized input written first so that the
pair's input is as warm/cold as in the network).

    python tools/profile_pair.py [l1 l2 l3 l4]      (SSDSB_PAIR_LAG=n to sweep the layer-2 lag)
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K  # noqa: E402

STAGES = {"l1": (128, 64), "l2": (64, 128), "l3": (32, 256), "l4": (16, 512)}


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B = 64
    g = torch.Generator().manual_seed(0)
    for name in (sys.argv[1:] or list(STAGES)):
        hw, C = STAGES[name]
        x = torch.randn((B, hw, hw, C), generator=g).to(torch.bfloat16).cuda()
        res = torch.randn((B, hw, hw, 4 * C), generator=g).to(torch.bfloat16).cuda()
        w1 = K.pack_weight(torch.randn((4 * C, C, 1, 1), generator=g) * 0.05).cuda()
        w2 = K.pack_weight(torch.randn((C, 4 * C, 1, 1), generator=g) * 0.02).cuda()
        b1, b2 = torch.zeros(4 * C).cuda(), torch.zeros(C).cuda()
        y1 = torch.empty((B, hw, hw, 4 * C), dtype=torch.bfloat16, device="cuda")
        y2 = torch.empty((B, hw, hw, C), dtype=torch.bfloat16, device="cuda")
        t3 = timed(lambda: K.conv2d(x, w1, b1, 1, 1, 1, 0, True, res, out=y1))
        t1 = timed(lambda: K.conv2d(y1, w2, b2, 1, 1, 1, 0, True, out=y2))

        def both():
            K.conv2d(x, w1, b1, 1, 1, 1, 0, True, res, out=y1)
            K.conv2d(y1, w2, b2, 1, 1, 1, 0, True, out=y2)
        tb = timed(both)
        tp = timed(lambda: K.conv1x1_pair(x, w1, b1, True, res, w2, b2, True, out1=y1, out2=y2))
        px = B * hw * hw
        floor_sep = px * 2 * (C + 4 * C + 4 * C + 4 * C + C) / 6.57e6        # us at the measured HBM peak
        floor_pair = px * 2 * (C + 4 * C + 4 * C + C) / 6.57e6
        print(f"{name}: conv3 {t3:.1f} + conv1 {t1:.1f} = {t3 + t1:.1f} us (back to back {tb:.1f}); pair {tp:.1f} us; "
              f"HBM floors {floor_sep:.0f} / {floor_pair:.0f} us")


if __name__ == "__main__":
    main()
