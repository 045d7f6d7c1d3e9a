#!/bin/bash
# round-2 GPU batch C: failing-test details, decode_large debug, timings, launch lists (small outputs only)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_box_ops.py -q -m gpu -k "negpos or decode_large or packed" --tb=short 2>&1 | tail -120 > gpurun_out/r2c_tests_failing.log
timeout 120 python tools/debug_decode_large.py 80 20000 > gpurun_out/r2c_debug_decode_80.txt 2>&1
timeout 120 python tools/debug_decode_large.py 16 3000 > gpurun_out/r2c_debug_decode_16.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_dropin_vs_reference.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r2c_tests_model.log
timeout 300 python tools/profile_misc.py loss decode_large nms > gpurun_out/r2c_misc_timings.txt 2>&1
# launch lists: durations + DRAM bytes per launch (CSV, small)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2c_launches_misc.csv python tools/profile_misc.py loss decode_large nms > /dev/null 2>&1
timeout 900 ncu --metrics $M --clock-control none -s 80 -c 160 --csv --log-file gpurun_out/r2c_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
timeout 300 python bench.py --config cfg4 --steps 10 --no-cpu > gpurun_out/r2c_bench_cfg4.json 2> gpurun_out/r2c_bench_cfg4.err
timeout 300 python bench.py --config cfg2 --steps 10 --no-cpu > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/r2c_bench_cfg2.err
rm -f gpurun_out/*.ncu-rep
tail -30 gpurun_out/r2c_tests_failing.log
cat gpurun_out/r2c_debug_decode_80.txt gpurun_out/r2c_debug_decode_16.txt gpurun_out/r2c_misc_timings.txt
tail -5 gpurun_out/r2c_tests_model.log
