#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -m gpu --tb=short 2>&1 | tail -8 > gpurun_out/r2p_tests.log
cat gpurun_out/r2p_tests.log
timeout -s KILL 300 python tools/profile_misc.py mb > gpurun_out/r2p_mb_timings.txt 2>&1
grep -v "^ " gpurun_out/r2p_mb_timings.txt
timeout -s KILL 400 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2p_bench_cfg3.json 2> gpurun_out/r2p_bench_cfg3.err
SSDSB_MBFUSE=1 timeout -s KILL 400 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2p_bench_cfg3_allfused.json 2> /dev/null
tail -3 gpurun_out/r2p_bench_cfg3.err
python - <<'PY'
import json
for c in ("cfg3","cfg3_allfused"):
    try:
        d=json.load(open(f"gpurun_out/r2p_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d.get("gpu_launches"), d.get("self_check"))
    except Exception as e:
        print(c, "ERR", e)
PY
