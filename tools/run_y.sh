#!/bin/bash
# 2-GPU checks: weak and strong scaling lines through torchrun, the way the driver launches them
mkdir -p gpurun_out
run() { # name, extra args
  n=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/r2y_$n.json 2> gpurun_out/r2y_$n.err
  tail -c 300 gpurun_out/r2y_$n.err | tail -3
}
run cfg2_weak --config cfg2 --steps 10 --no-cpu
run cfg2_strong --config cfg2 --steps 10 --no-cpu --scaling strong
run cfg3_strong --config cfg3 --steps 10 --no-cpu --scaling strong
run cfg4_weak --config cfg4 --steps 10 --no-cpu
run cfg5_strong --config cfg5 --steps 5 --no-cpu --scaling strong
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --cpu-images 2 > gpurun_out/r2y_reference_arm.json 2> gpurun_out/r2y_reference_arm.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2y_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d.get("n_gpus"), d.get("scaling"), round(d["value"]), round(d["e2e"]["value"]), d["config"].get("per_gpu_batch"), d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("samples"), d.get("impl"))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json",".err")).read()[-400:])
PY
