#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dwconv_cuda.py tests/test_layernorm_cuda.py -q --tb=short 2>&1 | tail -3
timeout 600 python tools/graph_step.py 16 2>&1 | tail -4
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01_dw3.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "[ncu launches] exit $?"
