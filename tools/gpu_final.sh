#!/bin/bash
# Round-end evidence run: bench (with CPU baseline), reference arm, launch list, kernel rooflines, sweep, family.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "[bench] exit $?"; cut -c1-300 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "[ref] exit $?"; cut -c1-300 gpurun_out/bench_reference.json
bash tools/gpu_ncu_step.sh
python tools/summarize_launches.py gpurun_out/launches_step.csv > gpurun_out/launches_step.md
timeout 600 python tools/kernel_rooflines.py > gpurun_out/kernel_rooflines.md 2> gpurun_out/kernel_rooflines.err; echo "[rooflines] exit $?"
timeout 600 python tools/sweep_dattn.py > gpurun_out/sweep_dattn.md 2> gpurun_out/sweep_dattn.err; echo "[sweep] exit $?"
timeout 600 python tools/bench_family.py > gpurun_out/family.md 2> gpurun_out/family.err; echo "[family] exit $?"; tail -4 gpurun_out/family.md
