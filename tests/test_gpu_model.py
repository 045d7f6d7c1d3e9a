"""The tcgen05 conv stack + fused heads (ssds_pytorch_b200.model.SSDResNet) against the model oracle
run with the SAME rounding policy (bf16 weights/activations, fp32 accumulate), and the whole
detector (conv -> decode -> NMS) against the reference-pinned oracle chain.

Tolerances (the conv stack is a floating-point kernel; bf16 storage has 8 mantissa bits, and a
different fp32 summation order flips individual bf16 roundings that then propagate through ~50
layers): loc |err| <= 2e-2 * (1 + max|loc|), conf (sigmoid-ed) |err| <= 5e-4 + 4e-2*conf, against
values of O(1) and O(0.01).  Decode/NMS on identical inputs are held to the box-op bar in test_gpu_box_ops.py."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {
    "bifpn": ("RegNetX032", [[2, 3, 4, "Conv:S", "Conv:S"], [192, 432, 1008, 1008, 256]], 20, 2, [128, 128]),
    "mbv2": ("MobileNetV2", [[5, 7, "Conv:S", "Conv:S", "Conv:S", "Conv:S"], [96, 320, 512, 256, 256, 128]], 20, 2,
             [300, 300]),
    "fpn50": ("ResNet50", [[3, 4, 5, "Conv:S", "Conv:S"], [512, 1024, 2048, 2048, 256]], 20, 2, [256, 256]),
    "r18": ("ResNet18", [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]], 20, 3, [96, 160]),
    "r50": ("ResNet50", [[3, 4, 5, "Conv:S", "Conv:S", "Conv:S"], [512, 1024, 2048, 512, 256, 256]], 80, 1,
            [256, 256]),
    # the model of the reference's shipped experiments/cfgs/tests/test.yml (YOLOV3 + ResNet18, 80 classes)
    "yolo": ("ResNet18", [[3, 4, 5], [128, 256, 512]], 80, 2, [128, 160]),
    "yolo50x": ("ResNet50", [[3, 4, 5, "Conv:S"], [512, 1024, 2048, 512]], 20, 2, [128, 128]),
    # [r2] YOLOv3 with [in, out] depth pairs (yolo.py:120-131): neck widths chosen independently of the backbone's
    "yolo_pairs": ("ResNet18", [[3, 4, 5, "Conv:S"], [[128, 128], [256, 128], [512, 256], 256]], 20, 2, [128, 160]),
    # [r2] YOLOv4 neck (SPP + PAN, yolo.py:161-392)
    "yolo4": ("ResNet18", [[3, 4, 5, "Conv:S"], [128, 256, 512, 256]], 20, 2, [128, 160]),
    "yolo4_50": ("ResNet50", [[3, 4, 5], [512, 1024, 2048]], 80, 1, [192, 192]),
}
NBOX = {"yolo": [6, 6, 9]}
SSDS_OF = {"fpn50": "SSDFPN", "bifpn": "SSDBiFPN", "yolo": "YOLOV3", "yolo50x": "YOLOV3", "yolo_pairs": "YOLOV3",
           "yolo4": "YOLOV4",
           "yolo4_50": "YOLOV4"}


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import ssds_pytorch_b200 as S
    return S


def build(tag, S):
    from ssds_pytorch_b200 import synth
    from ssds_pytorch_b200.model import engine_for
    nets, fl, ncls, B, image = CASES[tag]
    L = len(fl[0])
    ssds = SSDS_OF.get(tag, "SSD")
    nb = NBOX.get(tag, [6] * L)
    sd = synth.synthetic_state_dict(nets, fl, nb, ncls, seed=11, style="test", ssds=ssds)
    x = torch.rand((B, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    model = engine_for(ssds, nets)(sd, fl, ncls, nb, device="cuda").eval()
    return sd, fl, x, model, image, ncls


@pytest.mark.parametrize("tag", ["r18", "r50", "fpn50", "mbv2", "bifpn", "yolo", "yolo50x", "yolo_pairs", "yolo4",
                                 "yolo4_50"])
def test_conv_stack_vs_oracle_bf16_policy(env, tag):
    from oracle import model_oracle as M
    sd, fl, x, model, image, ncls = build(tag, env)
    loc, conf = model(x.cuda())
    torch.cuda.synchronize()
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        fwd = {"fpn50": M.ssdfpn_resnet_forward, "mbv2": M.ssd_mobilenetv2_forward, "bifpn": M.ssdbifpn_forward,
               "yolo": M.yolov3_resnet_forward, "yolo50x": M.yolov3_resnet_forward,
               "yolo_pairs": M.yolov3_resnet_forward,
               "yolo4": M.yolov4_resnet_forward, "yolo4_50": M.yolov4_resnet_forward}.get(tag, M.ssd_resnet_forward)
        rloc, rconf = fwd(sd_gpu, x.cuda(), fl, training=False, policy="bf16")
    worst_l = worst_c = 0.0
    for l, c, rl, rc in zip(loc, conf, rloc, rconf):
        assert l.shape == rl.shape and c.shape == rc.shape
        worst_l = max(worst_l, (l - rl).abs().max().item() / (1.0 + rl.abs().max().item()))
        worst_c = max(worst_c, ((c - rc).abs() / (5e-4 + 4e-2 * rc)).max().item())
    print(f"{tag}: max |loc err|/(1+max|loc|) {worst_l:.3e}, max conf err / tol {worst_c:.3f}")
    limit_l, limit_c = 2e-2, 1.0
    if tag == "yolo4_50":
        # the deep ResNet50 + SPP (4096 -> 1024 3x3) + PAN stack with synthetic weights amplifies bf16 storage
        # rounding: the bf16-policy oracle itself deviates from the fp32 oracle by MORE than the tolerance (conf ~3x),
        # and a different fp32 summation order moves the roundings as much.  Bound the kernel by that measured noise.
        with torch.no_grad():
            floc, fconf = fwd(sd_gpu, x.cuda(), fl, training=False, policy="fp32")
        noise_l = max((a - b).abs().max().item() / (1.0 + b.abs().max().item()) for a, b in zip(rloc, floc))
        noise_c = max(((a - b).abs() / (5e-4 + 4e-2 * b)).max().item() for a, b in zip(rconf, fconf))
        limit_l, limit_c = max(limit_l, 1.5 * noise_l), max(limit_c, 1.5 * noise_c)
        print(f"{tag}: bf16-policy oracle vs fp32 oracle: loc {noise_l:.3e}, conf {noise_c:.2f} x tol")
    assert worst_l <= limit_l
    assert worst_c <= limit_c
    # CUDA-graph replay gives bit-identical outputs
    keep = [t.clone() for t in loc + conf]
    loc2, conf2 = model(x.cuda(), use_graph=True)
    torch.cuda.synchronize()
    for a, b in zip(keep, loc2 + conf2):
        assert torch.equal(a, b)


def test_training_mode_returns_logits(env):
    sd, fl, x, model, image, ncls = build("r18", env)
    _, conf = model(x.cuda())
    conf = [c.clone() for c in conf]
    model.train()
    _, logits = model(x.cuda())
    torch.cuda.synchronize()
    for c, lg in zip(conf, logits):
        np.testing.assert_allclose(torch.sigmoid(lg).cpu().numpy(), c.cpu().numpy(), rtol=1e-4, atol=1e-6)
    model.eval()


def test_detector_end_to_end_vs_reference_golden(env):
    """SSDDetector (uint8-free fp32 path) vs the detections the REFERENCE produced for the same
    weights/input (tests/golden/model_small.npz): the conv stack runs in bf16, so compare
    tie-/noise-aware: most reference detections must be found with the same class and a box
    within 1.5 px, with a score within 5 %."""
    from ssds_pytorch_b200 import synth
    from ssds_pytorch_b200.ssds import SSDDetector
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "model_small.npz"))
    tag = "r18"
    nets, fl, ncls, B, image = CASES[tag]
    L = len(fl[0])
    cfg = {"MODEL": {"SSDS": "SSD", "NETS": nets, "IMAGE_SIZE": image, "NUM_CLASSES": ncls,
                     "FEATURE_LAYER": fl, "SIZES": [[2.0, 2.828]] * L, "ASPECT_RATIOS": [[1, 2, 0.5]] * L},
           "DATASET": {"PREPROC": {"MEAN": 0, "STD": 1}}}
    sd = synth.synthetic_state_dict(nets, fl, [6] * L, ncls, seed=11, style="test")
    det = SSDDetector(cfg, sd)
    np.testing.assert_array_equal(list(det.anchors.keys()), gold[tag + "_strides"])
    for i, a in enumerate(det.anchors.values()):
        np.testing.assert_array_equal(a.cpu().numpy(), gold[f"{tag}_anchors{i}"])
    x = torch.rand((B, 3, image[0], image[1]), generator=torch.Generator().manual_seed(1234))
    s, b, c = det.detect_device(x.cuda())
    torch.cuda.synchronize()
    s, b, c = s.cpu().numpy(), b.cpu().numpy(), c.cpu().numpy()
    rs, rb, rc = gold[tag + "_det_scores"], gold[tag + "_det_boxes"], gold[tag + "_det_classes"]
    found = total = 0
    for i in range(B):
        for j in range(rs.shape[1]):
            if rs[i, j] <= 0:
                continue
            total += 1
            m = (c[i] == rc[i, j]) & (np.abs(b[i] - rb[i, j]).max(axis=1) <= 1.5) & \
                (np.abs(s[i] - rs[i, j]) <= 0.05 * rs[i, j] + 1e-5)
            found += bool(m.any())
    print(f"end-to-end: {found}/{total} reference detections reproduced")
    assert found >= 0.9 * total
    # numpy API of the reference facade (ssds.py:41-68): [N,3,H,W] float input, int boxes/classes
    out = det(x.numpy())
    assert out[0].shape == (B, 100) and out[1].dtype.kind == "i" and out[2].dtype.kind == "i"
    np.testing.assert_allclose(out[0], s, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["r50", "fpn50"])
def test_pair_launch_plan_is_bit_identical_to_separate_launches(env, tag, monkeypatch):
    """The ResNet50 plan with every eligible conv3 -> next conv1 fused into one conv_pair launch (threshold
    lowered so that it applies to this small input: all four stages, stage transitions included) must
    give exactly the tensors of the plan that launches them separately."""
    from ssds_pytorch_b200 import model as MD
    sd, fl, x, _, image, ncls = build(tag, env)
    L = len(fl[0])
    ssds = "SSDFPN" if tag == "fpn50" else "SSD"
    outs = []
    for pair in (False, True):
        monkeypatch.setattr(MD, "PAIR_MIN_TILES_PER_SM", 0 if pair else 10 ** 9)
        m = MD.engine_for(ssds, "ResNet50")(sd, fl, ncls, [6] * L, device="cuda").eval()
        plan = m.plan_for(x.cuda())
        loc, conf = m(x.cuda())
        torch.cuda.synchronize()
        outs.append((plan["launches"], [t.clone() for t in loc], [t.clone() for t in conf]))
    assert outs[1][0] == outs[0][0] - 15          # 15 of ResNet50's 16 bottlenecks have a successor
    for a, b in zip(outs[0][1] + outs[0][2], outs[1][1] + outs[1][2]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["always", "auto"])
def test_mobilenet_plan_with_fused_inverted_residuals_is_bit_identical(env, mode, monkeypatch):
    """[r2] SSD-MobileNetV2 300x300 with every inverted residual the fused kernel supports in ONE launch
    (SSDSB_MBFUSE=1: 16 of 17 blocks, the 320-channel one has no configuration), and with the plan's own per-block
    choice, vs three launches per block (SSDSB_NO_MBFUSE=1): exactly the same loc / conf tensors."""
    from ssds_pytorch_b200 import model as MD
    sd, fl, x, _, image, ncls = build("mbv2", env)
    L = len(fl[0])
    outs = []
    for fused in (False, True):
        monkeypatch.delenv("SSDSB_MBFUSE", raising=False)
        monkeypatch.delenv("SSDSB_NO_MBFUSE", raising=False)
        if not fused:
            monkeypatch.setenv("SSDSB_NO_MBFUSE", "1")
        elif mode == "always":
            monkeypatch.setenv("SSDSB_MBFUSE", "1")
        m = MD.engine_for("SSD", "MobileNetV2")(sd, fl, ncls, [6] * L, device="cuda").eval()
        plan = m.plan_for(x.cuda())
        loc, conf = m(x.cuda())
        torch.cuda.synchronize()
        n_mb = sum(v["kind"].startswith("mbconv") for v in plan["info"].values())
        outs.append((n_mb, plan["launches"], [t.clone() for t in loc], [t.clone() for t in conf]))
    monkeypatch.delenv("SSDSB_MBFUSE", raising=False)
    monkeypatch.delenv("SSDSB_NO_MBFUSE", raising=False)
    assert outs[0][0] == 0
    if mode == "always":
        assert outs[1][0] == 16 and outs[1][1] == outs[0][1] - (15 * 2 + 1)     # block 0 has no expand launch
    for a, b in zip(outs[0][2] + outs[0][3], outs[1][2] + outs[1][3]):
        assert torch.equal(a, b)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[1] at its real shape: SSD-ResNet50 512x512, the plan bench.py times
# ---------------------------------------------------------------------------------------------------------------
CFG2_FL = [[3, 4, 5, "Conv:S", "Conv:S", "Conv:S"], [512, 1024, 2048, 512, 256, 256]]


def _cfg2(B, style="init"):
    from ssds_pytorch_b200 import synth
    from ssds_pytorch_b200.model import engine_for
    sd = synth.synthetic_state_dict("ResNet50", CFG2_FL, [6] * 6, 80, seed=0, style=style)
    x = torch.randint(0, 256, (B, 512, 512, 3), generator=torch.Generator().manual_seed(1234), dtype=torch.uint8)
    model = engine_for("SSD", "ResNet50")(sd, CFG2_FL, 80, [6] * 6, device="cuda", mean=0.0, std=255.0).eval()
    return sd, x, model


def test_cfg2_full_shape_plan_vs_oracle_and_vs_1way_unpaired_plan(env, monkeypatch):
    """The 512x512, B=64 plan (4-way / 2-way / weight-resident conv_igemm instantiations, conv_pair launches,
    CUDA-graph replay — exactly what bench.py times) against
      (a) the SAME model planned with SSDSB_WAYS=1 SSDSB_NO_PAIR=1, un-graphed: bit-identical;
      (b) the model oracle under the bf16 policy (ssd.py:42-74 / resnet.py:41-56 restated with torch fp32
          ops) on sampled images of the batch (images are independent): the test_conv_stack tolerance."""
    from oracle import model_oracle as M
    from ssds_pytorch_b200 import conv as K
    B = 64
    sd, x, model = _cfg2(B, style="test")
    xg = x.cuda()
    loc, conf = model(xg, use_graph=True)
    torch.cuda.synchronize()
    got = [t.clone() for t in loc + conf]
    plan = model.plan_for(xg)
    kinds = [v["kind"] for v in plan["info"].values()]
    assert sum(k.startswith("pair1x1") for k in kinds) >= 6, kinds       # layer1 + layer2 pairs at B=64
    # which instantiations does this plan launch?  replay it un-graphed, reading the launch record
    seen = set()
    for s in plan["steps"]:
        s()
        ll = K.last_launch()
        seen.add((ll["block_n"], ll["block_k"], ll["ways"], ll["b_resident"]))
    torch.cuda.synchronize()
    assert any(w == 4 for _, _, w, _ in seen) and any(w == 2 for _, _, w, _ in seen), seen
    assert any(r == 1 for _, _, _, r in seen), seen
    for a, b in zip(got, loc + conf):
        assert torch.equal(a, b), "graph replay differs from eager replay"
    # (a) 1-way, unpaired plan
    monkeypatch.setenv("SSDSB_WAYS", "1")
    monkeypatch.setenv("SSDSB_NO_PAIR", "1")
    _, _, plain = _cfg2(B, style="test")
    loc1, conf1 = plain(xg, use_graph=False)
    torch.cuda.synchronize()
    assert not any(v["kind"].startswith("pair1x1") for v in plain.plan_for(xg)["info"].values())
    monkeypatch.delenv("SSDSB_WAYS")
    monkeypatch.delenv("SSDSB_NO_PAIR")
    for a, b in zip(got, loc1 + conf1):
        assert torch.equal(a, b), "multi-way / paired plan differs from the 1-way unpaired plan"
    del plain, loc1, conf1
    # (b) oracle on sampled images
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    worst_l = worst_c = 0.0
    for i in (0, 21, 63):
        xi = (xg[i:i + 1].float() / 255.0).permute(0, 3, 1, 2).contiguous()
        with torch.no_grad():
            rloc, rconf = M.ssd_resnet_forward(sd_gpu, xi, CFG2_FL, training=False, policy="bf16")
        for l, c, rl, rc in zip(got[:6], got[6:], rloc, rconf):
            worst_l = max(worst_l, (l[i:i + 1] - rl).abs().max().item() / (1.0 + rl.abs().max().item()))
            worst_c = max(worst_c, ((c[i:i + 1] - rc).abs() / (5e-4 + 4e-2 * rc)).max().item())
    print(f"cfg2 512x512 B={B}: max |loc err|/(1+max|loc|) {worst_l:.3e}, max conf err / tol {worst_c:.3f}")
    assert worst_l <= 2e-2 and worst_c <= 1.0


def test_cfg2_bf16_stack_vs_fp32_reference_detections(env):
    """How far is the bf16 B200 stack from the reference's fp32 arithmetic AT cfg 2 (the number
    north_star's 1e-4 bar cannot apply to: bf16 storage has 8 mantissa bits)?  fp32 oracle (torch fp32
    convs + BN, TF32 off = what the reference computes) -> the same Decoder, vs the bf16 engine -> Decoder,
    on 4 images of 512x512.  Reported (printed) and bounded loosely; profiles/NOTES.md quotes the numbers."""
    from oracle import model_oracle as M
    from ssds_pytorch_b200.ssds import SSDDetector
    import bench
    B = 4
    sd, x, _ = _cfg2(B, style="test")
    det = SSDDetector(bench.cfg_dict(), sd, use_graph=False)
    s, b, c = [t.cpu().numpy() for t in det.detect_device(x.cuda())]
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    xi = (x.cuda().float() / 255.0).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        rloc, rconf = M.ssd_resnet_forward(sd_gpu, xi, CFG2_FL, training=False, policy="fp32")
    rs, rb, rc = [t.cpu().numpy() for t in det.decoder(rloc, rconf, det.anchors)]
    torch.cuda.synchronize()
    found = total = 0
    dscore, dbox = [], []
    for i in range(B):
        for j in range(rs.shape[1]):
            if rs[i, j] <= 0:
                continue
            total += 1
            m = (c[i] == rc[i, j]) & (np.abs(b[i] - rb[i, j]).max(axis=1) <= 2.0)
            if m.any():
                k = int(np.argmax(m))
                found += 1
                dscore.append(abs(s[i, k] - rs[i, j]) / rs[i, j])
                dbox.append(np.abs(b[i, k] - rb[i, j]).max())
    print(f"cfg2 fp32-reference vs bf16 engine: {found}/{total} detections matched (same class, box within 2 px); "
          f"median/max rel score diff {np.median(dscore):.3e}/{np.max(dscore):.3e}, "
          f"median/max box diff {np.median(dbox):.3f}/{np.max(dbox):.3f} px")
    assert total > 0 and found >= 0.8 * total


def test_detect_host_pipelined_unpinned_batches(env):
    """detect_host stages unpinned input through a per-slot pinned buffer; a caller that
    pipelines batches through alternating slots WITHOUT syncing must not overwrite bytes whose H2D copy is still
    queued.  Six different unpinned batches in flight, results vs the synchronous path, bit for bit."""
    from ssds_pytorch_b200 import synth
    from ssds_pytorch_b200.ssds import SSDDetector
    nets, fl, ncls, _, image = CASES["r18"]
    L = len(fl[0])
    cfg = {"MODEL": {"SSDS": "SSD", "NETS": nets, "IMAGE_SIZE": image, "NUM_CLASSES": ncls,
                     "FEATURE_LAYER": fl, "SIZES": [[2.0, 2.828]] * L, "ASPECT_RATIOS": [[1, 2, 0.5]] * L},
           "DATASET": {"PREPROC": {"MEAN": 0, "STD": 255}}}
    sd = synth.synthetic_state_dict(nets, fl, [6] * L, ncls, seed=11, style="test")
    det = SSDDetector(cfg, sd)
    g = torch.Generator().manual_seed(77)
    batches = [torch.randint(0, 256, (8, image[0], image[1], 3), generator=g, dtype=torch.uint8).numpy() for _ in range(6)]
    want = []
    for x in batches:                                   # synchronous reference results
        s, b, c = det.detect_device(torch.from_numpy(x).cuda())
        want.append(torch.cat([s[..., None], b, c[..., None]], -1).cpu())
    torch.cuda.synchronize()
    outs = [torch.empty((8, 100, 6), dtype=torch.float32).pin_memory() for _ in batches]
    for i, x in enumerate(batches):                     # no sync between calls, alternating slots
        det.detect_host(x, out=outs[i], slot=i % 2)
    det.join()
    torch.cuda.synchronize()
    for i in range(len(batches)):
        assert torch.equal(outs[i], want[i]), f"batch {i} was corrupted in the pipelined path"
