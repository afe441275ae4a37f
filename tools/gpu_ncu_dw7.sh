#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_dwconv7.py > gpurun_out/plain_dw7.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"dwconv7" \
  -c 2 -o gpurun_out/prof_r01_dwconv7 python tools/profile_dwconv7.py > gpurun_out/ncu_dw7.log 2>&1
echo "[ncu] exit $?"; tail -1 gpurun_out/ncu_dw7.log
