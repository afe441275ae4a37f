#!/bin/bash
# A/B of an environment setting on the 1-GPU bench: gpu_ab_val.sh VAR=VALUE
mkdir -p gpurun_out
show() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', d['value'], 'img/s', d['ms_per_step'], 'ms')"; }
for i in 1 2; do
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_err.log | show default; tail -2 gpurun_out/ab_err.log
  env $1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/ab_err.log | show "$1"; tail -2 gpurun_out/ab_err.log
done
