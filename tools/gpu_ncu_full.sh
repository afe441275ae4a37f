#!/bin/bash
# --set full captures of the dat_b200 kernels inside one real training step (batch 16, 512x512).
# The raw-page CSV is exported on the box; the .ncu-rep is dropped when it would not fit the
# 64 MiB return limit.
mkdir -p gpurun_out
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 2400 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"attn_fwd_tc|attn_bwd_tc|offset_pos_fwd_vec|sample_fwd|sample_bwd_dx|gemm_tc_kernel|gemm_tc_wgrad|layernorm_fwd|layernorm_bwd_kernel|dwconv_cl_kernel|dwconv_wgrad_kernel" \
    -c ${NCU_COUNT:-60} -o gpurun_out/prof_r01_step_full python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "[ncu full] exit $?"; tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_r01_step_full.ncu-rep --page raw --csv > gpurun_out/prof_r01_step_full_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/prof_r01_step_full.ncu-rep); echo "rep bytes $sz"
if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/prof_r01_step_full.ncu-rep; echo "rep dropped (too large)"; fi
du -sh gpurun_out
