#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary32.txt
run() { tag=$1; shift; timeout 600 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary32.txt; grep -E "passed|failed|Error|timed out|^E  " "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-300 | tee -a gpurun_out/summary32.txt; }
run parity tests/test_cuda_parity.py
run ln tests/test_layernorm_cuda.py
timeout 600 python tools/time_blocks.py > gpurun_out/time_blocks32.log 2>&1; head -4 gpurun_out/time_blocks32.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_o.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_r01_o.json; tail -3 gpurun_out/bench_err.log
