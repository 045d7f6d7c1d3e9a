"""Debug aid: run the large-top_n decode on one clustered level and compare the selection state the kernels leave
in the workspace (b1, above1, b2, above2, n_sure, n_maybe per (image, level)) with a numpy emulation."""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
import ssds_pytorch_b200 as S                 # noqa: E402
from ssds_pytorch_b200 import _lib            # noqa: E402
from oracle import box_oracle as O            # noqa: E402

rng = np.random.default_rng(99)
A, C, H, W, stride = 3, 80, int(sys.argv[1]) if len(sys.argv) > 1 else 80, 80, 16
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
anc = O.generate_anchors(stride, [1, 2, 0.5], [4.0])
logits = rng.normal(-4.6, 0.25, (2, A * C, H, W)).astype(np.float32)
conf = (1.0 / (1.0 + np.exp(-logits))).astype(np.float32)
loc = rng.normal(0, 0.3, (2, A * 4, H, W)).astype(np.float32)
got = S.decode(torch.from_numpy(conf), torch.from_numpy(loc), stride, 0.01, K, torch.from_numpy(anc), True,
               return_indices=True)
torch.cuda.synchronize()
ws = list(_lib._workspaces.values())[0]
seg = ws[:2 * 64].view(torch.int32).cpu().numpy().reshape(2, 16)
n = A * C * H * W
for b in range(2):
    c = conf[b].reshape(-1)
    u = c.view(np.uint32) | np.uint32(0x80000000)
    ok = c >= np.float32(0.01)
    bin1 = (u >> 21).astype(np.int64)
    h1 = np.bincount(bin1[ok], minlength=2048)
    cum = 0
    for d in range(2047, -1, -1):
        if cum + h1[d] >= K:
            break
        cum += h1[d]
    b1, above1 = d, cum
    bin2 = ((u >> 10) & 2047).astype(np.int64)
    h2 = np.bincount(bin2[ok & (bin1 == b1)], minlength=2048)
    cum = 0
    for d in range(2047, -1, -1):
        if cum + h2[d] >= K - above1:
            break
        cum += h2[d]
    b2, above2 = d, cum
    print(f"image {b}: numpy b1={b1} above1={above1} b2={b2} above2={above2} sure={above1 + above2} maybe={h2[b2]} total={ok.sum()}")
    print(f"          gpu   b1={seg[b, 0]} above1={seg[b, 1]} b2={seg[b, 2]} above2={seg[b, 3]} n_sure={seg[b, 4]} "
          f"n_maybe={seg[b, 5]} tickets={seg[b, 6]},{seg[b, 7]} total={seg[b, 8]}")
    hist1 = ws[256:].view(torch.int32)      # may be misaligned if segs > 256 B; print only when layout is simple
exp = O.decode(conf, loc, stride, 0.01, K, anc, True, return_indices=True)
gi = got[3].cpu().numpy().astype(np.int64)
bad = np.argwhere(gi != exp[3])
print("index mismatches:", len(bad), "first:", bad[:5].tolist(), "got", gi[tuple(bad[0])] if len(bad) else None,
      "exp", exp[3][tuple(bad[0])] if len(bad) else None)
print("valid counts got/exp:", (gi >= 0).sum(1), (exp[3] >= 0).sum(1))
