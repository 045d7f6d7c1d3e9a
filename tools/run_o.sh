#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 200 python tools/mbconv_profile.py > gpurun_out/r2o_mbconv_events.txt 2>&1
cat gpurun_out/r2o_mbconv_events.txt
