"""repeat the no-expand block (and one expand block) many times against the three-launch result; count mismatching runs"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K
g = torch.Generator().manual_seed(1)
for case in [(2, 150, 150, 32, 32, 32, 1, False), (5, 150, 150, 32, 32, 32, 1, False), (2, 75, 75, 32, 160, 32, 1, True)]:
    N, H, W, Cin, hid, Cout, stride, res = case
    x = torch.randn((N, H, W, Cin), generator=g).to(torch.bfloat16).cuda()
    has_e = hid != Cin
    we = K.pack_weight(torch.randn((hid, Cin, 1, 1), generator=g) * (2.0 / np.sqrt(Cin))).cuda() if has_e else None
    be = (torch.randn((hid,), generator=g) * 0.5).cuda() if has_e else None
    wd = K.pack_dw_weight(torch.randn((hid, 1, 3, 3), generator=g) * 0.4).cuda()
    bd = (torch.randn((hid,), generator=g) * 0.3).cuda()
    wp = K.pack_weight(torch.randn((Cout, hid, 1, 1), generator=g) * (1.0 / np.sqrt(hid))).cuda()
    bp = (torch.randn((Cout,), generator=g) * 0.2).cuda()
    h = K.conv2d(x, we, be, 1, 1, 1, 0, 2) if has_e else x
    d = K.dwconv3x3(h, wd, bd, stride, 2)
    want = K.conv2d(d, wp, bp, 1, 1, 1, 0, 0, residual=x if res else None)
    bad_runs, worst = 0, 0
    for it in range(60):
        got = K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0))
        nb = int((got != want).sum().item())
        bad_runs += nb > 0
        worst = max(worst, nb)
    print(os.environ.get("SSDSB_MB_DEBUG", "0"), case, K.mbconv_last_launch()["tile_w"], K.mbconv_last_launch()["tile_h"],
          f"bad runs {bad_runs}/60, worst {worst} elements")
