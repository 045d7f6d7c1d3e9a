#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_dropin_vs_reference.py -q -m gpu --tb=short -k "maxpool5 or yolo4 or yolov4 or oracle_bf16" 2>&1 | tail -25 > gpurun_out/r2aa_tests_yolov4.log
cat gpurun_out/r2aa_tests_yolov4.log
