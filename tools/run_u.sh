#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout -s KILL 900 ncu --metrics $M --clock-control none -s 200 -c 420 --csv --log-file gpurun_out/r2u_launches_cfg5.csv python bench.py --config cfg5 --steps 2 --warmup 3 --no-cpu --no-selfcheck --no-sections > /dev/null 2>&1
ls -la gpurun_out/r2u*
