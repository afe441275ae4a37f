#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
DAT_B200_ATTN_BWD_GENERIC_TABLE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c80-200
done
bash tools/gpu_ncu_step.sh
