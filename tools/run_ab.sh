#!/bin/bash
# final verification of the round's last commits: full GPU suite, smoke, one short bench per config
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 > gpurun_out/r2ab_tests_gpu_full.log
cat gpurun_out/r2ab_tests_gpu_full.log
timeout -s KILL 300 python __graft_entry__.py --smoke 2>&1 | tail -5 > gpurun_out/r2ab_smoke.log
cat gpurun_out/r2ab_smoke.log
for c in cfg2 cfg3 cfg4; do
  timeout -s KILL 400 python bench.py --config $c --steps 10 --no-cpu > gpurun_out/r2ab_bench_$c.json 2> gpurun_out/r2ab_bench_$c.err
done
python - <<'PY'
import json
for c in ("cfg2","cfg3","cfg4"):
    try:
        d=json.load(open(f"gpurun_out/r2ab_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d["roofline"]["frac"])
    except Exception as e:
        print(c, "ERR", e)
PY
