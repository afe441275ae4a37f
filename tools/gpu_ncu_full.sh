#!/bin/bash
# --set full captures of the dat_b200 kernels inside one real training step (batch 16, 512x512)
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 2400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"attn_fwd_tc|attn_bwd_tc|offset_pos_fwd_vec|sample_fwd|sample_bwd_dx|gemm_tc_kernel|gemm_tc_wgrad|layernorm_fwd|layernorm_bwd_kernel|dwconv_cl_kernel|dwconv_wgrad_kernel" \
    -c 60 -o gpurun_out/prof_r01_step_full python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "[ncu full] exit $?"; tail -2 gpurun_out/ncu_full.log
