#!/bin/bash
mkdir -p gpurun_out
for cfg in "256 108" "128 108" "128 72" "128 54" "64 54" "256 72"; do set -- $cfg; echo "=== BN cap $1, smem budget $2 KB"; DAT_B200_GEMM_BN=$1 DAT_B200_GEMM_SMEM_KB=$2 timeout 300 python tools/time_gemm_abi.py 2>&1 | grep -v wgrad | grep -E "s0|s2"; done | tee gpurun_out/time_gemm_sweep.log
