#!/bin/bash
# round-2 GPU batch B: new kernels' parity tests, timings, ncu captures of every non-conv kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_loss_step.py tests/test_gpu_dropin_vs_reference.py tests/test_gpu_box_ops.py tests/test_gpu_model.py tests/test_gpu_conv_baseline_shapes.py -q -m gpu -s 2>&1 | tail -80 > gpurun_out/r2b_tests.log
timeout 300 python tools/profile_misc.py > gpurun_out/r2b_misc_timings.txt 2>&1
timeout 300 python bench.py --config cfg4 --steps 10 --no-cpu > gpurun_out/r2b_bench_cfg4.json 2> gpurun_out/r2b_bench_cfg4.err
timeout 300 python bench.py --config cfg5stress --steps 5 --no-cpu > gpurun_out/r2b_bench_cfg5stress.json 2> gpurun_out/r2b_bench_cfg5stress.err
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/r2b_misc python tools/profile_misc.py loss match decode decode_large nms > gpurun_out/r2b_ncu_misc.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:dwconv\|pack_image\|maxpool\|upsample\|bifpn -o gpurun_out/r2b_layout python tools/profile_misc.py dw layout > gpurun_out/r2b_ncu_layout.log 2>&1
tail -5 gpurun_out/r2b_tests.log
cat gpurun_out/r2b_misc_timings.txt
