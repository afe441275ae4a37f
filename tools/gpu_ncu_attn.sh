#!/bin/bash
# --set full of the attention forward / backward kernels of one block at stage 2 and stage 0
mkdir -p gpurun_out
timeout 300 python tools/run_stage_kernels.py 2 fwdbwd > gpurun_out/plain_attn2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)_tc" --launch-skip 4 -c 2 \
  -o gpurun_out/prof_r01_attn_s2 python tools/run_stage_kernels.py 2 fwdbwd > gpurun_out/ncu_attn2.log 2>&1
echo "[ncu] exit $?"; tail -2 gpurun_out/ncu_attn2.log; ls -la gpurun_out/*.ncu-rep
