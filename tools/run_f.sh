#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_box_ops.py tests/test_gpu_loss_step.py -q -m gpu --tb=short 2>&1 | tail -40 > gpurun_out/r2f_tests.log
timeout 300 python tools/profile_misc.py loss decode_large nms > gpurun_out/r2f_misc_timings.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2f_launches_misc.csv python tools/profile_misc.py decode_large nms > /dev/null 2>&1
# full-metric captures of single kernels, exported to CSV here (the .ncu-rep files are too big to bring back)
timeout 600 ncu --set full --clock-control none -k regex:loss_step -s 1 -c 1 -o /tmp/ls python tools/profile_misc.py loss > /dev/null 2>&1
ncu -i /tmp/ls.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_loss_step_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:dl_ -s 8 -c 8 -o /tmp/dl python tools/profile_misc.py decode_large > /dev/null 2>&1
ncu -i /tmp/dl.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_decode_large_raw.csv 2>/dev/null
timeout 300 python bench.py --config cfg5stress --steps 10 --no-cpu > gpurun_out/r2f_bench_cfg5stress.json 2> gpurun_out/r2f_bench_cfg5stress.err
tail -8 gpurun_out/r2f_tests.log
cat gpurun_out/r2f_misc_timings.txt
ls -la gpurun_out/ | tail -12
