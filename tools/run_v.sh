#!/bin/bash
mkdir -p gpurun_out
{
echo "== converged elect.sync issue in a non-inlined function (default)"; timeout -s KILL 200 python tools/conv_variants.py 2>&1 | grep -A1 "{}:"
echo "== SSDSB_MMA_LANE0=1"; SSDSB_MMA_LANE0=1 timeout -s KILL 200 python tools/conv_variants.py 2>&1 | grep -A1 "{}:"
} > gpurun_out/r2v_mma_issue_ab.txt 2>&1
cat gpurun_out/r2v_mma_issue_ab.txt
timeout -s KILL 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_baseline_shapes.py -q -m gpu --tb=short 2>&1 | tail -5
for c in cfg5 cfg2; do
  timeout -s KILL 400 python bench.py --config $c --steps 10 --no-cpu > gpurun_out/r2v_bench_$c.json 2> gpurun_out/r2v_bench_$c.err
  SSDSB_MMA_LANE0=1 timeout -s KILL 400 python bench.py --config $c --steps 10 --no-cpu > gpurun_out/r2v_bench_${c}_lane0.json 2> /dev/null
done
python - <<'PY'
import json
for c in ("cfg5","cfg5_lane0","cfg2","cfg2_lane0"):
    try:
        d=json.load(open(f"gpurun_out/r2v_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()})
    except Exception as e:
        print(c, "ERR", e)
PY
