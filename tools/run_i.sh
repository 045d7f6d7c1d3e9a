#!/bin/bash
# round-2 final single-GPU records: every config with the reference CPU baseline, launch lists with DRAM bytes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss_step.py -q -m gpu --tb=short 2>&1 | tail -5 > gpurun_out/r2i_tests_loss.log
timeout 120 python tools/profile_misc.py loss > gpurun_out/r2i_loss_timings.txt 2>&1
SSDSB_PDL=1 timeout 300 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2i_bench_cfg3_pdl.json 2> /dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/r2i_launches_cfg2.csv python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/r2i_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -s 150 -c 330 --csv --log-file gpurun_out/r2i_launches_cfg4.csv python bench.py --config cfg4 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
for c in cfg2 cfg3 cfg4 cfg5 cfg5stress; do
  timeout 600 python bench.py --config $c --steps 20 > gpurun_out/r2i_bench_$c.json 2> gpurun_out/r2i_bench_$c.err
done
timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/r2i_bench_cfg2_reference_arm.json 2> gpurun_out/r2i_ref.err
cat gpurun_out/r2i_tests_loss.log gpurun_out/r2i_loss_timings.txt
python - <<'PY'
import json
for c in ("cfg3_pdl","cfg2","cfg3","cfg4","cfg5","cfg5stress"):
    try:
        d=json.load(open(f"gpurun_out/r2i_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(c, "ERR", e)
PY
tail -c 600 gpurun_out/r2i_bench_cfg2_reference_arm.json
