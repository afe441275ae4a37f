#!/bin/bash
# --set full of one launch of each kernel that changed in round 2, taken from inside one real training step
# (tools/profile_step.py: eager fwd+bwd of the DAT-T++ backbone, batch 16).  One ncu run per kernel: the regex picks the
# kernel, --launch-skip picks a stage-2 launch where the step has many (stage 2 = 18 of the 28 blocks).
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 || exit 1
run() {   # name regex skip
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 -c 1 \
    -o gpurun_out/r02_end_$1 -f python tools/profile_step.py > gpurun_out/ncu_r02_end_$1.log 2>&1
  echo "[ncu $1] exit $?"
}
run wgrad "gemm_tc_wgrad_kernel" 40
run ln_fwd_sub "layernorm_fwd_sub_kernel" 1
run ln_bwd_sub "layernorm_bwd_sub_kernel" 1
run dw3_fwd "dwconv3_fwd_kernel<__nv_bfloat16, __nv_bfloat16, 2" 10
run dw3_bwd "dwconv3_bwd_kernel<__nv_bfloat16, __nv_bfloat16, 2" 10
run gather_kv "gather_kv_tc_kernel" 1
run k1_fwd "offset_pos_fwd_unrolled_kernel" 5
run bwd_dx "sample_bwd_dx_kernel" 5
run dw7_fwd "dwconv7_fwd_kernel<__nv_bfloat16, __nv_bfloat16" 6
ls -la gpurun_out/r02_end_*.ncu-rep
