#!/bin/bash
mkdir -p gpurun_out
DAT_B200_NCHW_CONVS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-200
bash tools/gpu_ncu_step.sh
python tools/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_step.md
grep -v "dat::" gpurun_out/launches_step.md | head -30
