#!/bin/bash
# launch list (gpu__time_duration) of exactly one training step (batch 16, 512x512)
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_step.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "[ncu launches] exit $?"
