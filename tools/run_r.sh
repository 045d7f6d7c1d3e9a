#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 python tools/mbconv_diag.py > gpurun_out/r2r_diag.txt 2>&1
cat gpurun_out/r2r_diag.txt
