#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_box_ops.py tests/test_gpu_loss_step.py -q -m gpu --tb=short 2>&1 | tail -60 > gpurun_out/r2e_tests.log
timeout 300 python tools/profile_misc.py loss match decode_large nms > gpurun_out/r2e_misc_timings.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2e_launches_misc.csv python tools/profile_misc.py loss decode_large nms > /dev/null 2>&1
for c in cfg5stress cfg4; do
  timeout 300 python bench.py --config $c --steps 10 --no-cpu > gpurun_out/r2e_bench_$c.json 2> gpurun_out/r2e_bench_$c.err
done
rm -f gpurun_out/*.ncu-rep
tail -25 gpurun_out/r2e_tests.log
cat gpurun_out/r2e_misc_timings.txt
python - <<'PY'
import json
for c in ("cfg5stress","cfg4"):
    try:
        d=json.load(open(f"gpurun_out/r2e_bench_{c}.json"))
        print(c, d["value"], d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()})
    except Exception as e:
        print(c, "ERR", e)
PY
