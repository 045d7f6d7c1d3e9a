#!/bin/bash
# r2w: final single-GPU evidence of the round: full GPU test suite, smoke(), every config with the reference CPU arm,
# launch lists (time + DRAM bytes) of cfg 2 / 3, per-layer rooflines of cfg 2, ncu --set full of the depthwise kernel
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -6 > gpurun_out/r2w_tests_gpu_full.log
cat gpurun_out/r2w_tests_gpu_full.log
timeout -s KILL 300 python __graft_entry__.py --smoke 2>&1 | tail -6 > gpurun_out/r2w_smoke.log
cat gpurun_out/r2w_smoke.log
for c in cfg2 cfg3 cfg4 cfg5 cfg5stress; do
  timeout -s KILL 600 python bench.py --config $c --steps 20 > gpurun_out/r2w_bench_$c.json 2> gpurun_out/r2w_bench_$c.err
done
timeout -s KILL 600 python bench.py > gpurun_out/r2w_bench_default_flags.json 2> gpurun_out/r2w_bench_default_flags.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 > gpurun_out/r2w_bench_cfg2_reference_arm.json 2> gpurun_out/r2w_ref.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -s KILL 900 ncu --metrics $M --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/r2w_launches_cfg2.csv python bench.py --config cfg2 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
timeout -s KILL 900 ncu --metrics $M --clock-control none -s 60 -c 140 --csv --log-file gpurun_out/r2w_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
SSDSB_NO_BRANCH=1 timeout -s KILL 900 ncu --metrics $M --clock-control none -s 100 -c 100 --csv --log-file gpurun_out/r2w_launches_cfg2_one_stream.csv python bench.py --config cfg2 --steps 1 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
timeout -s KILL 300 python tools/layer_rooflines.py gpurun_out/r2w_launches_cfg2_one_stream.csv > gpurun_out/r2w_layer_rooflines_cfg2.md 2> gpurun_out/r2w_layer_rooflines.err
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:dwconv3x3_stream -s 4 -c 2 -o /tmp/dw_full python tools/profile_misc.py dw > /dev/null 2>&1
ncu -i /tmp/dw_full.ncu-rep --page raw --csv > gpurun_out/r2w_ncu_full_dwconv_stream.csv 2>/dev/null
python - <<'PY'
import json
for c in ("cfg2","cfg3","cfg4","cfg5","cfg5stress","default_flags"):
    try:
        d=json.load(open(f"gpurun_out/r2w_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d.get("cpu_baseline",{}).get("value"), d.get("gpu_launches"), d["roofline"]["frac"], d["clocks"])
    except Exception as e:
        print(c, "ERR", e)
PY
tail -c 500 gpurun_out/r2w_bench_cfg2_reference_arm.json
tail -5 gpurun_out/r2w_layer_rooflines_cfg2.md
ls -la gpurun_out/r2w*
