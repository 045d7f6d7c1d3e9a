"""Parity of the conv kernels AT THE SHAPES bench.py TIMES (BASELINE.json configs[1]: SSD-ResNet50 512x512,
64 images per GPU) — the multi-way (interleaved accumulator) and weight-resident instantiations of
conv_igemm_kernel, partial groups with ghost tiles, multi-group persistence, and conv_pair_kernel at
full size — vs a plain torch fp32 reference of the same op on the bf16-rounded operands
(reference computation: ssds/modeling/ssds/ssd.py:42-74 over nets/resnet.py:41-56).

Every case asserts (through ssdsb_conv_last_launch) WHICH instantiation the launch heuristics picked, so
a change of the heuristics cannot silently move a case back onto the WAYS=1 path.

Tolerance: |err| <= 2e-2 * max(1, |ref|) for bf16 outputs (1 bf16 ulp = 0.4-0.8 %), as in test_gpu_conv.py;
WAYS / residency / pairing must not change a single bit (checked against the WAYS=1 launch)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from ssds_pytorch_b200 import conv
    return conv


def ref_conv(x_nhwc, w, b, stride, pad, relu, residual=None):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    y = F.conv2d(x, w.to(torch.bfloat16).float(), b, stride=stride, padding=pad)
    if residual is not None:
        y = y + residual.float().permute(0, 3, 1, 2)
    if relu:
        y = y.relu()
    return y.permute(0, 2, 3, 1)


# N, H, W, Cin, Cout, k, stride, pad, relu, residual, expect (subset of last_launch() that must hold)
CASES = [
    # layer1 3x3 64->64 @128x128: BLOCK_N=64, 4 ways (the 72 KiB weight slab does not fit next to 4-way stages)
    (64, 128, 128, 64, 64, 3, 1, 1, True, False, dict(block_n=64, ways=4, b_resident=0, ghost=0)),
    # layer1 conv1 1x1 64->64 @128x128: 4 ways, resident
    (64, 128, 128, 64, 64, 1, 1, 0, True, False, dict(block_n=64, ways=4, b_resident=1)),
    # layer1 conv3 1x1 64->256 + identity @128x128: BLOCK_N=256 (1 way), resident weights, residual prefetch
    (64, 128, 128, 64, 256, 1, 1, 0, True, True, dict(block_n=256, ways=1, b_resident=1)),
    # layer2 3x3 128->128 @64x64: BLOCK_N=128, 2 ways, weights streamed (288 KiB)
    (64, 64, 64, 128, 128, 3, 1, 1, True, False, dict(block_n=128, ways=2, b_resident=0, ghost=0)),
    # layer2.0 conv2 3x3/s2 128->128 @128x128 -> 64x64
    (64, 128, 128, 128, 128, 3, 2, 1, True, False, dict(block_n=128, ways=2)),
    # layer2.0 downsample 1x1/s2 256->512 @128x128 -> 64x64 (no ReLU)
    (64, 128, 128, 256, 512, 1, 2, 0, False, False, dict(block_n=256, ways=1)),
    # layer3 3x3 256->256 @32x32 and layer3 conv3 1x1 256->1024 + identity
    (64, 32, 32, 256, 256, 3, 1, 1, True, False, dict(block_n=256, ways=1)),
    (64, 32, 32, 256, 1024, 1, 1, 0, True, True, dict(block_n=256, ways=1)),
    # layer4 1x1 2048->512 @16x16 (long K) and 512->2048 + identity
    (64, 16, 16, 2048, 512, 1, 1, 0, True, False, dict(block_n=256, ways=1)),
    (64, 16, 16, 512, 2048, 1, 1, 0, True, True, dict(block_n=256, ways=1)),
    # partial last group (ghost tiles): 37 x (5 x 9) = 1665 M-tiles, 4 ways -> 417 groups, 1 real + 3 ghosts
    (37, 72, 80, 64, 64, 3, 1, 1, True, False, dict(block_n=64, ways=4, ghost=1)),
    # same with 2 ways: 21 x 45 = 945 M-tiles -> 473 groups, the last one half empty
    (21, 72, 80, 128, 128, 3, 1, 1, True, False, dict(block_n=128, ways=2, ghost=1)),
    # ghost tiles + residual (staged epilogue must skip nothing and clip everything; 6 staging slots -> 2 ways)
    (37, 72, 80, 64, 64, 1, 1, 0, True, True, dict(block_n=64, ways=2, ghost=1)),
    # 4 ways + resident weights + ghost tiles, no residual
    (37, 72, 80, 64, 64, 1, 1, 0, True, False, dict(block_n=64, ways=4, b_resident=1, ghost=1)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(str(v) for v in c[:8]))
def test_conv_at_baseline_shape(K, case, monkeypatch):
    N, H, W, Cin, Cout, k, stride, pad, relu, use_res, expect = case
    g = torch.Generator(device="cuda").manual_seed(sum(case[:8]))
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, Cin, k, k), generator=g, device="cuda") * (1.0 / np.sqrt(Cin * k * k))
    b = torch.randn((Cout,), generator=g, device="cuda")
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = torch.randn((N, Ho, Wo, Cout), generator=g, device="cuda").to(torch.bfloat16) if use_res else None
    wp = K.pack_weight(w.cpu()).cuda()
    y = torch.full((N, Ho, Wo, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    K.conv2d(x, wp, b, k, k, stride, pad, relu, res, out=y)
    got = K.last_launch()
    torch.cuda.synchronize()
    for key, v in expect.items():
        assert got[key] == v, f"launch heuristics changed: {got} (expected {expect})"
    assert got["groups"] > got["grid"], f"not persistent over several groups: {got}"
    # (1) bit-identical to the 1-way, non-interleaved launch of the same kernel family
    monkeypatch.setenv("SSDSB_WAYS", "1")
    y1 = K.conv2d(x, wp, b, k, k, stride, pad, relu, res)
    assert K.last_launch()["ways"] == 1
    monkeypatch.delenv("SSDSB_WAYS")
    torch.cuda.synchronize()
    assert torch.equal(y, y1), "multi-way / resident launch differs from the 1-way launch"
    del y1
    # (2) vs torch fp32 on the same bf16 operands, image by image (bounded memory)
    worst = 0.0
    for n0 in range(0, N, 8):
        ref = ref_conv(x[n0:n0 + 8], w, b, stride, pad, relu, res[n0:n0 + 8] if use_res else None)
        err = (y[n0:n0 + 8].float() - ref).abs() / ref.abs().clamp(min=1.0)
        worst = max(worst, err.max().item())
    assert worst <= 2e-2, f"max rel err {worst}"
    assert torch.isfinite(y.float()).all()


def test_stem_at_baseline_shape(K):
    """resnet.py:42-44 conv1 7x7/s2 at 512x512, B=64 through the windowed s2d stem (BLOCK_N=64, 4 ways)."""
    g = torch.Generator().manual_seed(5)
    N, H, W = 64, 512, 512
    w = torch.randn((64, 3, 7, 7), generator=g) * 0.1
    b = torch.randn((64,), generator=g) * 0.1
    img = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8).cuda()
    packed = K.pack_image_s2d(img, 0.0, 255.0, padded=True)
    y = K.conv2d(packed, K.pack_stem_weight_s2d(w).cuda(), b.cuda(), 4, 4, 1, 2, True,
                 Ho=H // 2, Wo=W // 2, x_kind=1, x_width=W // 2)
    got = K.last_launch()
    torch.cuda.synchronize()
    assert got["block_n"] == 64 and got["ways"] == 4, got
    wr, br = w.to(torch.bfloat16).float().cuda(), b.cuda()
    worst = 0.0
    for n0 in range(0, N, 8):
        xr = (img[n0:n0 + 8].float() / 255.0).to(torch.bfloat16).float().permute(0, 3, 1, 2)
        ref = F.conv2d(xr, wr, br, stride=2, padding=3).relu().permute(0, 2, 3, 1)
        err = (y[n0:n0 + 8].float() - ref).abs() / ref.abs().clamp(min=1.0)
        worst = max(worst, err.max().item())
    assert worst <= 2e-2, worst


@pytest.mark.parametrize("shape", [(64, 64, 64, 512), (64, 32, 32, 1024)])
def test_head_at_baseline_shape(K, shape):
    """multibox head 3x3 Cin->(24 loc + 480 conf) at the cfg-2 level sizes (ssd.py:100-103), fp32 NCHW out."""
    N, H, W, Cin = shape
    g = torch.Generator(device="cuda").manual_seed(H)
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((504, Cin, 3, 3), generator=g, device="cuda") * (1.0 / np.sqrt(Cin * 9))
    b = torch.cat([torch.zeros(24), torch.full((480,), -4.595)]).cuda()
    loc, conf = K.conv2d_head(x, K.pack_weight(w.cpu()).cuda(), b, 24, True)
    got = K.last_launch()
    torch.cuda.synchronize()
    assert got["block_n"] == 256 and got["groups"] > got["grid"], got
    wl, wc = 0.0, 0.0
    for n0 in range(0, N, 8):
        ref = F.conv2d(x[n0:n0 + 8].float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b, padding=1)
        wl = max(wl, (loc[n0:n0 + 8] - ref[:, :24]).abs().max().item())
        wc = max(wc, (conf[n0:n0 + 8] - ref[:, 24:].sigmoid()).abs().max().item())
    assert wl <= 1e-3 and wc <= 1e-4, (wl, wc)


PAIR_CASES = [
    (64, 128, 128, 64, 256, 64),     # layer1: conv3 64->256 (+identity) -> next conv1 256->64
    (64, 128, 128, 64, 256, 128),    # layer1 -> layer2 transition
    (64, 64, 64, 128, 512, 128),     # layer2
    (64, 64, 64, 128, 512, 256),     # layer2 -> layer3 transition
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=lambda c: "x".join(str(v) for v in c))
def test_conv_pair_at_baseline_shape(K, case):
    """conv_pair_kernel at the sizes model.py pairs at B=64 (>= 8 M-tiles per SM): bit-identical to two
    conv2d launches, and y1 vs torch fp32."""
    N, H, W, Cin, Cmid, Cout2 = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w1f = torch.randn((Cmid, Cin, 1, 1), generator=g, device="cuda") / np.sqrt(Cin)
    w2f = torch.randn((Cout2, Cmid, 1, 1), generator=g, device="cuda") / np.sqrt(Cmid)
    w1, w2 = K.pack_weight(w1f.cpu()).cuda(), K.pack_weight(w2f.cpu()).cuda()
    b1 = torch.randn((Cmid,), generator=g, device="cuda")
    b2 = torch.randn((Cout2,), generator=g, device="cuda")
    res = torch.randn((N, H, W, Cmid), generator=g, device="cuda").to(torch.bfloat16)
    r1 = K.conv2d(x, w1, b1, 1, 1, 1, 0, True, res)
    r2 = K.conv2d(r1, w2, b2, 1, 1, 1, 0, True)
    for _ in range(2):
        y1 = torch.full_like(r1, float("nan"))
        y2 = torch.full_like(r2, float("nan"))
        K.conv1x1_pair(x, w1, b1, True, res, w2, b2, True, out1=y1, out2=y2)
        torch.cuda.synchronize()
        assert torch.equal(y1, r1)
        assert torch.equal(y2, r2)
    ref = ref_conv(x[:8], w1f, b1, 1, 0, True, res[:8])
    err = (y1[:8].float() - ref).abs() / ref.abs().clamp(min=1.0)
    assert err.max().item() <= 2e-2
