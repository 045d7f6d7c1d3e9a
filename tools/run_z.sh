#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_loss_step.py -q -m gpu --tb=short 2>&1 | tail -6
timeout -s KILL 400 python bench.py --config cfg4 --steps 20 --no-cpu > gpurun_out/r2z_bench_cfg4.json 2> gpurun_out/r2z_bench_cfg4.err
tail -2 gpurun_out/r2z_bench_cfg4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2z_bench_cfg4.json"))
print("cfg4", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), d["e2e"])
PY
