#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2ae_memcheck.txt 2>&1
tail -25 gpurun_out/r2ae_memcheck.txt
