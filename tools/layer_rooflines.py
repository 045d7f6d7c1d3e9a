"""Per-launch roofline table of the cfg-2 step: joins an ncu launch list (gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum of `bench.py --steps 1 --warmup 3`) with the algorithmic work
the model plan records per launch (plan["info"]).  Needs a GPU (the plan allocates its buffers there).

    python tools/layer_rooflines.py profiles/r1q_launches_with_dram_bytes.csv > profiles/r1q_layer_rooflines.md

For every conv launch: time, algorithmic TFLOP/s as a fraction of the sustained bf16 peak, algorithmic and
measured DRAM GB/s as a fraction of the measured copy bandwidth, and the larger of the two = fraction of the
RELEVANT roofline.  ncu times are cold-cache and serialised: a few % pessimistic for the small layers.
"""
import csv
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ssds_pytorch_b200 import synth                      # noqa: E402
from ssds_pytorch_b200.ssds import SSDDetector           # noqa: E402


def load_launches(fn):
    lines = [l for l in open(fn) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {h: i for i, h in enumerate(hdr)}
    rows, order = {}, []
    for row in r:
        if len(row) < len(hdr):
            continue
        k = int(row[ix["ID"]])
        if k not in rows:
            rows[k] = {"name": row[ix["Kernel Name"]]}
            order.append(k)
        v = float(row[ix["Metric Value"]].replace(",", ""))
        u = row[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
        rows[k][row[ix["Metric Name"]]] = v * scale
    ids = [k for k in order if "pack_image" in rows[k]["name"]]
    return [rows[k] for k in order if ids[-2] <= k < ids[-1]]


def main():
    peaks = {"hbm_gbs": 6569.3, "bf16_tflops_sustained": 1455.4}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks.update(json.load(open(p)))
    fl = [[3, 4, 5, "Conv:S", "Conv:S", "Conv:S"], [512, 1024, 2048, 512, 256, 256]]
    cfg = {"MODEL": {"SSDS": "SSD", "NETS": "ResNet50", "IMAGE_SIZE": [512, 512], "NUM_CLASSES": 80, "FEATURE_LAYER": fl,
                     "SIZES": [[2.0, 2.828]] * 6, "ASPECT_RATIOS": [[1, 2, 0.5]] * 6}}
    sd = synth.synthetic_state_dict("ResNet50", fl, [6] * 6, 80, seed=0, style="init")
    det = SSDDetector(cfg, sd, use_graph=False)
    x = torch.zeros((64, 512, 512, 3), dtype=torch.uint8, device="cuda")
    plan = det.model.plan_for(x)
    launches = load_launches(sys.argv[1])
    conv = [l for l in launches if "conv_igemm" in l["name"] or "conv_pair" in l["name"]]
    info = [plan["info"][i] for i in sorted(plan["info"]) if plan["info"][i]["flops"] > 0]
    assert len(conv) == len(info), (len(conv), len(info))
    print(f"| # | launch | us | TFLOP/s | of {peaks['bf16_tflops_sustained']:.0f} | algorithmic GB/s | measured DRAM GB/s | "
          f"of {peaks['hbm_gbs']:.0f} | relevant roofline |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot_t = 0.0
    weighted = 0.0
    for i, (l, w) in enumerate(zip(conv, info)):
        t = l["gpu__time_duration.sum"]
        tf = w["flops"] / t / 1e6
        ag = w["bytes"] / t / 1e3
        mg = (l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) / t / 1e3
        ft, fh = tf / peaks["bf16_tflops_sustained"], max(ag, mg) / peaks["hbm_gbs"]
        best = max(ft, fh)
        tot_t += t
        weighted += best * t
        print(f"| {i} | {w['kind']} | {t:.1f} | {tf:.0f} | {ft:.2f} | {ag:.0f} | {mg:.0f} | {fh:.2f} | "
              f"**{best:.2f}** ({'tensor' if ft >= fh else 'hbm'}) |")
    print(f"\nconv launches: {tot_t:.0f} us (cold, serialised); time-weighted fraction of the relevant roofline: "
          f"{weighted / tot_t:.2f}")


if __name__ == "__main__":
    main()
