#!/bin/bash
# The standard GPU trip: all GPU tests (full log in gpurun_out/pytest.log), smoke(), 1-GPU bench, optional extra command.
#   usage: gpurun -- 'bash tools/gpu_check.sh [extra command ...]'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/pytest.log 2>&1; echo "[pytest] exit $?"; tail -8 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_check.json; tail -3 gpurun_out/bench_check.err
if [ $# -gt 0 ]; then "$@"; fi
