"""A/B of launch-heuristic knobs on one conv shape (CUDA-graph timed): the 3x3 64->64 @128x128, B=64 layer of cfg 2
under SSDSB_WAYS / SSDSB_STAGING / SSDSB_STORE_LAG.   python tools/conv_variants.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K                   # noqa: E402
from tools.profile_misc import timed_graph                # noqa: E402

g = torch.Generator(device="cuda").manual_seed(3)
for (n, h, cin, cout, k, stride) in [(64, 128, 64, 64, 3, 1), (64, 64, 128, 128, 3, 1)]:
    x = torch.randn((n, h, h, cin), generator=g, device="cuda").to(torch.bfloat16)
    w = K.pack_weight(torch.randn((cout, cin, k, k)) * (1.0 / np.sqrt(cin * k * k))).cuda()
    b = torch.zeros(cout, device="cuda")
    y = torch.empty((n, h, h, cout), dtype=torch.bfloat16, device="cuda")
    fl = 2.0 * n * h * h * cin * k * k * cout
    for env in [{}, {"SSDSB_WAYS": "2"}, {"SSDSB_WAYS": "1"}, {"SSDSB_STAGING": "2"}, {"SSDSB_WAYS": "2", "SSDSB_STAGING": "2"},
                {"SSDSB_STORE_LAG": "0"}]:
        for kk in ("SSDSB_WAYS", "SSDSB_STAGING", "SSDSB_STORE_LAG"):
            os.environ.pop(kk, None)
        os.environ.update(env)
        t = timed_graph(f"conv{k}x{k} {cin}->{cout} @{h} {env}", lambda: K.conv2d(x, w, b, k, k, stride, k // 2, 1, out=y))
        print("      ", K.last_launch(), f"{fl / t / 1e6:.0f} TFLOP/s")
