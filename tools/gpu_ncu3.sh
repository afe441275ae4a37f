#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01_dw2.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "[ncu launches] exit $?"
