#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_baseline_shapes.py tests/test_gpu_model.py -q -m gpu -s 2>&1 | tail -60 > gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 > gpurun_out/r2a_bench_cfg2.json 2> gpurun_out/r2a_bench_cfg2.err
for c in cfg3 cfg4 cfg5 cfg5stress; do
  timeout 600 python bench.py --config $c --steps 5 > gpurun_out/r2a_bench_$c.json 2> gpurun_out/r2a_bench_$c.err
done
tail -3 gpurun_out/r2a_tests.log
