#!/bin/bash
# --set full of the attention forward / backward kernels of one block.   usage: bash tools/gpu_ncu_attn.sh <stage> <fwd|fwdbwd> <kernel regex> <out name>
mkdir -p gpurun_out
STAGE=${1:-2}; MODE=${2:-fwd}; REGEX=${3:-attn_fwd_tc2}; OUT=${4:-prof_attn}
timeout 300 python tools/run_stage_kernels.py $STAGE $MODE > gpurun_out/plain_${OUT}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" --launch-skip 1 -c 1 \
  -o gpurun_out/$OUT python tools/run_stage_kernels.py $STAGE $MODE > gpurun_out/ncu_${OUT}.log 2>&1
echo "[ncu] exit $?"; tail -2 gpurun_out/ncu_${OUT}.log; ls -la gpurun_out/*.ncu-rep
