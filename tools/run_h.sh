#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -30 > gpurun_out/r2h_tests_all.log
timeout 300 python tools/profile_misc.py dw layout > gpurun_out/r2h_misc_timings.txt 2>&1
timeout 300 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2h_bench_cfg3.json 2> gpurun_out/r2h_bench_cfg3.err
tail -12 gpurun_out/r2h_tests_all.log
cat gpurun_out/r2h_misc_timings.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_cfg3.json"))
print("cfg3", d["value"], d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()})
PY
