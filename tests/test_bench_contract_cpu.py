"""The bench JSON contract (task statement, "bench.py keeps the contract below"): checked on the line the
final round-1 code printed on a B200 (profiles/r1q_bench_line.json, copied from gpurun, not under a profiler)
and on the static parts of bench.py that do not need a GPU."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_every_contract_key():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1q_bench_line.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"].split(" at ")[0] in base["metric"] and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None   # nothing published
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - d["config"]["global_batch"] / d["ms_per_step"] * 1e3) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]                 # copies inside the timed region
    assert e["h2d_bytes_per_step"] == 64 * 512 * 512 * 3 and e["d2h_bytes_per_step"] == 64 * 100 * 6 * 4
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    k = d["clocks"]
    assert k["sm_mhz"] <= k["sm_max_mhz"] and isinstance(k["reasons"], list)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(k["reasons"])


def test_bench_static_contract():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in src
    assert "oracle" in src and "/root/reference" not in src          # the reference tree is never read at run time
    traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_conv_traffic.json")))
    assert os.path.exists(os.path.join(ROOT, traffic["source"].split(" ")[0]))   # the ncu list the number comes from
