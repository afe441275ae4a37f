#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_dwconv3.py > gpurun_out/plain_dw3.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"dwconv3" \
  -c 8 -o gpurun_out/prof_r01_dwconv3 python tools/profile_dwconv3.py > gpurun_out/ncu_dw3.log 2>&1
echo "[ncu] exit $?"; tail -2 gpurun_out/ncu_dw3.log; ls -la gpurun_out/
