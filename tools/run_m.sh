#!/bin/bash
mkdir -p gpurun_out
{
timeout -s KILL 200 python tools/mbconv_profile.py
echo; echo "=========== SSDSB_MB_TILE=16x8"
timeout -s KILL 200 python tools/mbconv_profile.py 16x8
} > gpurun_out/r2m_mbconv_wait_profile.txt 2>&1
cat gpurun_out/r2m_mbconv_wait_profile.txt
