#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 > gpurun_out/r2af_tests_gpu_full.log
cat gpurun_out/r2af_tests_gpu_full.log
