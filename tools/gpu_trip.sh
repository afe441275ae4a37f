#!/bin/bash
# One GPU-box trip: parity tests (each group in its own process so a fault in one does not
# poison the rest), smoke, quick timings.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
for grp in "library or ref_points or pointwise" "offset_net" "sampling or rpe_bias" "attention_core" \
           "block_forward_fp32" "block_backward_fp32" "bf16" "pos_and_ref or cpu_tensor" "full_size"; do
  tag=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_cuda_parity.py -q -rA --tb=short -k "$grp" > "gpurun_out/pytest_${tag}.log" 2>&1
  echo "[$grp] exit $?" | tee -a gpurun_out/summary.txt
  tail -3 "gpurun_out/pytest_${tag}.log" | tee -a gpurun_out/summary.txt
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks.log 2>&1; echo "time exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/time_blocks.log
