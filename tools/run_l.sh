#!/bin/bash
# r2l: fused inverted-residual kernel — parity first (each step under its own timeout), then timings
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -k "mbconv" 2>&1 | tail -40 > gpurun_out/r2l_tests_mbconv.log
cat gpurun_out/r2l_tests_mbconv.log
if grep -q "failed\|error\|Killed" gpurun_out/r2l_tests_mbconv.log; then echo "PARITY NOT GREEN: skipping model/bench"; fi
timeout -s KILL 300 python tools/profile_misc.py mb > gpurun_out/r2l_mb_timings.txt 2>&1
cat gpurun_out/r2l_mb_timings.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_model.py -q -m gpu --tb=short -x 2>&1 | tail -8 > gpurun_out/r2l_tests_model.log
cat gpurun_out/r2l_tests_model.log
timeout -s KILL 400 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2l_bench_cfg3.json 2> gpurun_out/r2l_bench_cfg3.err
SSDSB_NO_MBFUSE=1 timeout -s KILL 400 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2l_bench_cfg3_unfused.json 2> /dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -s KILL 600 ncu --metrics $M --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r2l_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
tail -3 gpurun_out/r2l_bench_cfg3.err
python - <<'PY'
import json
for c in ("cfg3","cfg3_unfused"):
    try:
        d=json.load(open(f"gpurun_out/r2l_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d.get("self_check"), d.get("gpu_launches"))
    except Exception as e:
        print(c, "ERR", e)
PY
