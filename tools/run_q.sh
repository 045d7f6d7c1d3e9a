#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/conv_variants.py > gpurun_out/r2q_conv_variants.txt 2>&1
cat gpurun_out/r2q_conv_variants.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -s KILL 600 ncu --metrics $M --clock-control none -s 60 -c 140 --csv --log-file gpurun_out/r2q_launches_cfg3.csv python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --import-source on --clock-control none -k regex:mbconv_kernel -s 2 -c 3 -o /tmp/mbconv_full python tools/profile_misc.py mb > /dev/null 2>&1
ncu -i /tmp/mbconv_full.ncu-rep --page raw --csv > gpurun_out/r2q_ncu_full_mbconv.csv 2>/dev/null
ls -la gpurun_out/r2q*
