#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu --tb=short 2>&1 | tail -8 > gpurun_out/r2s_tests.log
cat gpurun_out/r2s_tests.log
timeout -s KILL 200 python tools/profile_misc.py dw > gpurun_out/r2s_dw_timings.txt 2>&1
cat gpurun_out/r2s_dw_timings.txt
timeout -s KILL 400 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2s_bench_cfg3.json 2> gpurun_out/r2s_bench_cfg3.err
tail -2 gpurun_out/r2s_bench_cfg3.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2s_bench_cfg3.json"))
print("cfg3", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d.get("gpu_launches"))
PY
