#!/bin/bash
# r2k: software-pipelined depthwise kernel + staged epilogue for Cout % 64 != 0 — parity, A/B timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_baseline_shapes.py tests/test_gpu_model.py -q -m gpu --tb=short -x 2>&1 | tail -15 > gpurun_out/r2k_tests.log
{
echo "== default"; timeout 200 python tools/profile_misc.py dw pw
echo "== SSDSB_DW_SHALLOW=1"; SSDSB_DW_SHALLOW=1 timeout 200 python tools/profile_misc.py dw
echo "== SSDSB_DW_ROWS=8"; SSDSB_DW_ROWS=8 timeout 200 python tools/profile_misc.py dw
echo "== SSDSB_DW_ROWS=16"; SSDSB_DW_ROWS=16 timeout 200 python tools/profile_misc.py dw
echo "== SSDSB_DW_ROWS=1000"; SSDSB_DW_ROWS=1000 timeout 200 python tools/profile_misc.py dw
echo "== SSDSB_DW_SIMPLE=1"; SSDSB_DW_SIMPLE=1 timeout 200 python tools/profile_misc.py dw
echo "== SSDSB_DIRECT_RAGGED=1"; SSDSB_DIRECT_RAGGED=1 timeout 200 python tools/profile_misc.py pw
} > gpurun_out/r2k_dw_pw_timings.txt 2>&1
for c in cfg3 cfg5 cfg2 cfg4; do
  timeout 600 python bench.py --config $c --steps 10 --no-cpu > gpurun_out/r2k_bench_$c.json 2> gpurun_out/r2k_bench_$c.err
done
SSDSB_DIRECT_RAGGED=1 timeout 600 python bench.py --config cfg3 --steps 10 --no-cpu > gpurun_out/r2k_bench_cfg3_direct_ragged.json 2> /dev/null
SSDSB_DIRECT_RAGGED=1 timeout 600 python bench.py --config cfg5 --steps 10 --no-cpu > gpurun_out/r2k_bench_cfg5_direct_ragged.json 2> /dev/null
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -s 100 -c 200 --csv --log-file gpurun_out/r2k_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
cat gpurun_out/r2k_tests.log
cat gpurun_out/r2k_dw_pw_timings.txt
python - <<'PY'
import json
for c in ("cfg3","cfg3_direct_ragged","cfg5","cfg5_direct_ragged","cfg2","cfg4"):
    try:
        d=json.load(open(f"gpurun_out/r2k_bench_{c}.json"))
        print(c, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), {k:round(v["ms_per_step"],3) for k,v in d["rooflines"].items()}, d.get("self_check"))
    except Exception as e:
        print(c, "ERR", e)
PY
