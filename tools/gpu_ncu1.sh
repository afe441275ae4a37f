#!/bin/bash
mkdir -p gpurun_out
python tools/run_stage_kernels.py 2 fwd > gpurun_out/plain_stage2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc|offset_pos_fwd|sample_fwd|gemm_tc" -s 8 -c 8 \
    -o gpurun_out/prof_r01_stage2_fwd python tools/run_stage_kernels.py 2 fwd > gpurun_out/ncu_stage2.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_stage2.log
