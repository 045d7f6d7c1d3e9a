"""Sweep tile geometry / hidden-chunk width of the fused inverted-residual kernel on the block shapes where three
launches are still as fast or faster (B=64):  python tools/mbconv_sweep.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K                   # noqa: E402
from tools.profile_misc import timed_graph                # noqa: E402

g = torch.Generator(device="cuda").manual_seed(8)
SHAPES = [(75, 32, 160, 32, 1, True, ["16x8", "8x16", "32x4", "25x5", "15x8", "8x15"], ["32"]),
          (10, 160, 960, 160, 1, True, ["10x10", "5x10", "10x5", "10x4", "5x5"], ["64", "32"]),
          (19, 96, 576, 160, 2, False, ["10x5", "10x10", "5x10", "10x4", "5x5"], ["64", "32"]),
          (150, 32, 96, 32, 2, False, ["15x5", "8x8", "16x4", "15x8", "10x8"], ["32"])]
for (h, cin, hid, cout, stride, res, tiles, hcs) in SHAPES:
    x = torch.randn((64, h, h, cin), generator=g, device="cuda").to(torch.bfloat16)
    we = K.pack_weight(torch.randn((hid, cin, 1, 1)) * (1.0 / np.sqrt(cin))).cuda()
    be = torch.zeros(hid, device="cuda")
    wd = K.pack_dw_weight(torch.randn((hid, 1, 3, 3)) * 0.3).cuda()
    bd = torch.zeros(hid, device="cuda")
    wp = K.pack_weight(torch.randn((cout, hid, 1, 1)) * (1.0 / np.sqrt(hid))).cuda()
    bp = torch.zeros(cout, device="cuda")
    ho = (h - 1) // stride + 1
    y = torch.empty((64, ho, ho, cout), dtype=torch.bfloat16, device="cuda")
    tag = f"{cin}->{hid}->{cout} s{stride} @{h}"
    for k in ("SSDSB_MB_TILE", "SSDSB_MB_HC"):
        os.environ.pop(k, None)
    t0 = timed_graph(f"{tag} default", lambda: K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0), out=y))
    print("      ", K.mbconv_last_launch())
    for hc in hcs:
        for tl in tiles:
            os.environ["SSDSB_MB_HC"], os.environ["SSDSB_MB_TILE"] = hc, tl
            try:
                t = timed_graph(f"{tag} hc={hc} tile={tl}", lambda: K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0), out=y))
                info = K.mbconv_last_launch()
                print("      ", {k: info[k] for k in ("hc", "tile_w", "tile_h", "pm", "dw_segments", "dw_rows", "x_buffers", "staging", "grid")})
            except Exception as e:       # noqa: BLE001
                print(f"{tag} hc={hc} tile={tl}: {type(e).__name__} {str(e)[:80]}")
