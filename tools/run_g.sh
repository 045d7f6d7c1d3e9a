#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_box_ops.py tests/test_gpu_loss_step.py tests/test_gpu_dropin_vs_reference.py -q -m gpu --tb=short -s 2>&1 | tail -40 > gpurun_out/r2g_tests.log
timeout 300 python tools/profile_misc.py loss decode_large nms > gpurun_out/r2g_misc_timings.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2g_launches_misc.csv python tools/profile_misc.py loss decode_large nms > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dwconv3x3_rows -s 2 -c 2 -o /tmp/dw python tools/profile_misc.py dw > /dev/null 2>&1
ncu -i /tmp/dw.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_dwconv_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:loss_step -s 1 -c 1 -o /tmp/ls python tools/profile_misc.py loss > /dev/null 2>&1
ncu -i /tmp/ls.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_loss_step_raw.csv 2>/dev/null
for c in cfg4 cfg5stress; do
  timeout 300 python bench.py --config $c --steps 10 --no-cpu > gpurun_out/r2g_bench_$c.json 2> gpurun_out/r2g_bench_$c.err
done
tail -12 gpurun_out/r2g_tests.log
cat gpurun_out/r2g_misc_timings.txt
python - <<'PY'
import json
for c in ("cfg4","cfg5stress"):
    try:
        d=json.load(open(f"gpurun_out/r2g_bench_{c}.json"))
        print(c, d["value"], d["ms_per_step"], {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()})
    except Exception as e:
        print(c, "ERR", e)
PY
