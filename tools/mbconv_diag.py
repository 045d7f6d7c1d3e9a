import sys
import numpy as np
import torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K
g = torch.Generator().manual_seed(1)
N, H, W, Cin, hid, Cout, stride, res = (2, 150, 150, 32, 32, 32, 1, False)
x = torch.randn((N, H, W, Cin), generator=g).to(torch.bfloat16).cuda()
wd = K.pack_dw_weight(torch.randn((hid, 1, 3, 3), generator=g) * 0.4).cuda()
bd = (torch.randn((hid,), generator=g) * 0.3).cuda()
wp = K.pack_weight(torch.randn((Cout, hid, 1, 1), generator=g) * (1.0 / np.sqrt(hid))).cuda()
bp = (torch.randn((Cout,), generator=g) * 0.2).cuda()
d = K.dwconv3x3(x, wd, bd, stride, 2)
want = K.conv2d(d, wp, bp, 1, 1, 1, 0, 0)
got = K.mbconv(x, None, None, wd, bd, wp, bp, stride, res, (2, 2, 0))
torch.cuda.synchronize()
bad = (got != want)
pix = bad.any(dim=3)
print("bad elements", int(bad.sum()), "bad pixels", int(pix.sum()), "of", pix.numel(), K.mbconv_last_launch())
dmax = d.float().amax(dim=3)
print("pixels whose dw vector has a clipped (== 6) element:", int((dmax >= 6.0).sum()), " of them bad:", int(((dmax >= 6.0) & pix).sum()))
print("bad pixels without a clipped element:", int((pix & (dmax < 6.0)).sum()))
# identity project: recover the dw output of the fused kernel (Cout == hid == 32)
eye = K.pack_weight(torch.eye(32).reshape(32, 32, 1, 1)).cuda()
z = torch.zeros(32, device="cuda")
dg = K.mbconv(x, None, None, wd, bd, eye, z, stride, res, (2, 2, 0))
torch.cuda.synchronize()
db = dg != d
print("fused dw output differs from dwconv3x3 at", int(db.sum()), "elements")
idx = db.nonzero()[:12]
for i in idx.tolist():
    n, h, w, c = i
    print(i, "fused", float(dg[n, h, w, c]), "ref", float(d[n, h, w, c]))
