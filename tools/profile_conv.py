"""Run ONE conv configuration of the cfg-2 model (SSD-ResNet50 512^2, B=64) a few times so that
`ncu --set full -k regex:conv_igemm -s 2 -c 1` captures a warm launch of exactly that layer.

    python tools/profile_conv.py head_l0 | l1_conv3 | l3_conv2 | l3_conv3 | l2_conv1 | stem
Prints CUDA-event timings (not under ncu) when run without a profiler.
"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K  # noqa: E402

CFG = {
    # name: (H, W, Cin, Cout, k, stride, pad, relu, residual, head)
    "l1_conv1": (128, 128, 256, 64, 1, 1, 0, True, False, False),
    "l1_conv2": (128, 128, 64, 64, 3, 1, 1, True, False, False),
    "l1_conv3": (128, 128, 64, 256, 1, 1, 0, True, True, False),
    "l1_down": (128, 128, 64, 256, 1, 1, 0, False, False, False),
    "l2_conv1": (64, 64, 512, 128, 1, 1, 0, True, False, False),
    "l2_conv2": (64, 64, 128, 128, 3, 1, 1, True, False, False),
    "l2_conv3": (64, 64, 128, 512, 1, 1, 0, True, True, False),
    "l3_conv1": (32, 32, 1024, 256, 1, 1, 0, True, False, False),
    "l3_conv2": (32, 32, 256, 256, 3, 1, 1, True, False, False),
    "l3_conv3": (32, 32, 256, 1024, 1, 1, 0, True, True, False),
    "l4_conv2": (16, 16, 512, 512, 3, 1, 1, True, False, False),
    "head_l0": (64, 64, 512, 504, 3, 1, 1, False, False, True),
    "head_l1": (32, 32, 1024, 504, 3, 1, 1, False, False, True),
}


def main():
    for name in sys.argv[1:]:
        one(name)


def one(name):
    B = 64
    g = torch.Generator().manual_seed(0)
    if name == "stem":
        img = torch.randint(0, 256, (B, 512, 512, 3), generator=g, dtype=torch.uint8).cuda()
        packed = K.pack_image_s2d(img, 0.0, 255.0, padded=True)
        w = K.pack_stem_weight_s2d(torch.randn((64, 3, 7, 7), generator=g) * 0.05).cuda()
        b = torch.zeros(64).cuda()
        run = lambda: K.conv2d(packed, w, b, 4, 4, 1, 2, True, Ho=256, Wo=256, x_kind=1, x_width=256)
        flops = 2 * 147 * 64 * B * 256 * 256
    else:
        H, W, Cin, Cout, k, stride, pad, relu, res, head = CFG[name]
        x = torch.randn((B, H, W, Cin), generator=g).to(torch.bfloat16).cuda()
        w = K.pack_weight(torch.randn((Cout, Cin, k, k), generator=g) * 0.02).cuda()
        b = torch.zeros(Cout).cuda()
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        flops = 2 * Cin * k * k * Cout * B * Ho * Wo
        if head:
            loc = torch.empty((B, 24, Ho, Wo), device="cuda")
            conf = torch.empty((B, Cout - 24, Ho, Wo), device="cuda")
            run = lambda: K.conv2d_head(x, w, b, 24, True, loc=loc, conf=conf)
        else:
            r = torch.randn((B, Ho, Wo, Cout), generator=g).to(torch.bfloat16).cuda() if res else None
            y = torch.empty((B, Ho, Wo, Cout), dtype=torch.bfloat16, device="cuda")
            run = lambda: K.conv2d(x, w, b, k, k, stride, pad, relu, r, out=y)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: {ms * 1e3:.1f} us  {flops / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
