#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --tb=short -k "mbconv" 2>&1 | tail -30 > gpurun_out/r2n_tests_mbconv.log
cat gpurun_out/r2n_tests_mbconv.log
timeout -s KILL 300 python tools/profile_misc.py mb > gpurun_out/r2n_mb_timings.txt 2>&1
grep -v "^ " gpurun_out/r2n_mb_timings.txt
timeout -s KILL 200 python tools/mbconv_profile.py > gpurun_out/r2n_mbconv_wait_profile.txt 2>&1
cat gpurun_out/r2n_mbconv_wait_profile.txt
