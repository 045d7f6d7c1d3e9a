"""Small launches of the round-2 kernels for compute-sanitizer (memcheck): the fused inverted residual (stride 1 / 2,
with / without expand, residual), the pipelined depthwise kernel, the SPP max-pool, a ragged-Cout conv.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K                   # noqa: E402

g = torch.Generator().manual_seed(0)
for (N, H, W, Cin, hid, Cout, stride, res) in [(1, 21, 13, 32, 64, 32, 1, True), (1, 7, 9, 64, 128, 96, 2, False),
                                               (2, 19, 19, 32, 32, 32, 1, False), (1, 10, 10, 160, 320, 160, 1, True)]:
    x = torch.randn((N, H, W, Cin), generator=g).to(torch.bfloat16).cuda()
    has_e = not (hid == Cin and not res)
    we = K.pack_weight(torch.randn((hid, Cin, 1, 1), generator=g) * 0.3).cuda() if has_e else None
    be = torch.zeros(hid, device="cuda") if has_e else None
    wd = K.pack_dw_weight(torch.randn((hid, 1, 3, 3), generator=g) * 0.3).cuda()
    bd = torch.zeros(hid, device="cuda")
    wp = K.pack_weight(torch.randn((Cout, hid, 1, 1), generator=g) * 0.1).cuda()
    bp = torch.zeros(Cout, device="cuda")
    y = K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0))
    h = K.conv2d(x, we, be, 1, 1, 1, 0, 2) if has_e else x
    d = K.dwconv3x3(h, wd, bd, stride, 2)
    r = K.conv2d(d, wp, bp, 1, 1, 1, 0, 0, residual=x if res else None)
    torch.cuda.synchronize()
    print("mbconv", (N, H, W, Cin, hid, Cout, stride, res), "equal:", bool(torch.equal(y, r)))
x = torch.randn((1, 6, 7, 64), generator=g).to(torch.bfloat16).cuda()
cat = torch.zeros((1, 6, 7, 256), dtype=torch.bfloat16, device="cuda")
cat[..., :64] = x
for k in range(3):
    K.maxpool5x5s1(cat[..., k * 64:(k + 1) * 64], cat[..., (k + 1) * 64:(k + 2) * 64])
torch.cuda.synchronize()
print("maxpool5x5s1 ok", float(cat.float().abs().sum()) > 0)
