#!/bin/bash
# ncu --set full: persistent tcgen05 GEMM (3 MLP shapes), attention backward with dS streamed out, table-gradient GEMMs
mkdir -p gpurun_out
timeout 300 python tools/profile_gemm.py > gpurun_out/plain_gemm.log 2>&1 &&
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_tc_persistent" \
  -c 3 -o gpurun_out/prof_r01b_gemm python tools/profile_gemm.py > gpurun_out/ncu_gemm.log 2>&1
echo "[ncu gemm] exit $?"
for st in 2 0; do
  DAT_B200_SERIAL_WGRAD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd_tc_kernel|rpe_table_grad|attn_fwd_tc" \
    -c 3 -o gpurun_out/prof_r01b_attn_s$st python tools/run_block_bwd.py $st 1 > gpurun_out/ncu_attn_$st.log 2>&1
  echo "[ncu attn s$st] exit $?"
done
