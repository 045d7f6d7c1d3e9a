#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/mbconv_sweep.py > gpurun_out/r2ad_mbconv_sweep.txt 2>&1
grep -v "^ " gpurun_out/r2ad_mbconv_sweep.txt
