#!/bin/bash
# ncu --set full of the two kernels of bench.py's roofline leg, under the leg's own conditions (L2 flushed before every
# launch): the last timed launch of each kernel.  Then the DRAM-traffic summary bench.py reads.
mkdir -p gpurun_out
timeout 300 python bench.py --roofline-only > gpurun_out/roofline_only.json 2> gpurun_out/roofline_only.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_persistent" --launch-skip 12 -c 1 \
  -o gpurun_out/r02_roofline_gemm python bench.py --roofline-only > gpurun_out/ncu_roofline_gemm.log 2>&1
echo "[ncu gemm] exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc" --launch-skip 12 -c 1 \
  -o gpurun_out/r02_roofline_attn python bench.py --roofline-only > gpurun_out/ncu_roofline_attn.log 2>&1
echo "[ncu attn] exit $?"; ls -la gpurun_out/r02_roofline_*.ncu-rep
