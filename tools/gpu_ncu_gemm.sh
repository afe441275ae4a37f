#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_gemm.py > gpurun_out/plain_gemm.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" \
  -c 2 -o gpurun_out/prof_r01_gemm python tools/profile_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "[ncu] exit $?"; tail -2 gpurun_out/ncu_gemm.log
