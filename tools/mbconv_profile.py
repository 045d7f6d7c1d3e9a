"""Wait profile of the fused inverted-residual kernel: for a few MobileNetV2-SSD block shapes (B=64) print, per warp
role of CTA 0, the cycles spent blocked on each barrier, per processed chunk.   SSDSB_MB_PROF=1 is set here.
    python tools/mbconv_profile.py [BWxBH]"""
import os
import sys

os.environ.setdefault("SSDSB_MB_PROF", "1")      # 1: event stamps only (light), 2: + per-barrier wait cycles
if len(sys.argv) > 1:
    os.environ["SSDSB_MB_TILE"] = sys.argv[1]
import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ssds_pytorch_b200 import conv as K                   # noqa: E402

g = torch.Generator(device="cuda").manual_seed(8)
for (h, cin, hid, cout, stride, res) in [(75, 32, 160, 32, 1, True), (19, 64, 384, 64, 1, True),
                                         (150, 32, 96, 32, 2, False), (150, 32, 32, 32, 1, False)]:
    x = torch.randn((64, h, h, cin), generator=g, device="cuda").to(torch.bfloat16)
    has_e = not (hid == cin and stride == 1 and not res)
    we = K.pack_weight(torch.randn((hid, cin, 1, 1)) * (1.0 / np.sqrt(cin))).cuda() if has_e else None
    be = torch.zeros(hid, device="cuda") if has_e else None
    wd = K.pack_dw_weight(torch.randn((hid, 1, 3, 3)) * 0.3).cuda()
    bd = torch.zeros(hid, device="cuda")
    wp = K.pack_weight(torch.randn((cout, hid, 1, 1)) * (1.0 / np.sqrt(hid))).cuda()
    bp = torch.zeros(cout, device="cuda")
    for _ in range(3):
        y = K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    y = K.mbconv(x, we, be, wd, bd, wp, bp, stride, res, (2, 2, 0))
    e1.record()
    torch.cuda.synchronize()
    pr = K.mbconv_profile()
    n = max(pr["chunks"], 1)
    print(f"\n{cin}->{hid}->{cout} s{stride} @{h}: {e0.elapsed_time(e1) * 1e3:.1f} us, CTA 0: {pr['cycles']} cycles, "
          f"{pr['chunks']} chunks = {pr['cycles'] / n:.0f} cycles/chunk   {K.mbconv_last_launch()}")
    print("   chunk: dw_start dw_done | project: wait issue | expand committed | convert: start done   (cycles)")
    for gi, ev in enumerate(pr["events"]):
        print(f"   g={gi + 8:2d}: {ev[0]:8d} {ev[1]:8d} | {ev[2]:8d} {ev[3]:8d} | {ev[4]:8d} | {ev[5]:8d} {ev[6]:8d}")
    for wn, row in pr["waits"].items():
        if wn in ("cvt1", "cvt2", "cvt3") or (wn.startswith("dw") and wn != "dw0"):
            continue
        tot = sum(v for k, v in row.items() if k != "WORK")
        print(f"   {wn:6s} blocked {tot / n:7.0f}/chunk ({100.0 * tot / pr['cycles']:4.1f} %): " +
              ", ".join(f"{k} {v / n:.0f}" for k, v in sorted(row.items(), key=lambda kv: -kv[1])))
