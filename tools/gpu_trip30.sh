#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary30.txt
run() { tag=$1; shift; timeout 600 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary30.txt; grep -E "passed|failed|Error|timed out|^E  " "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-300 | tee -a gpurun_out/summary30.txt; }
run ln tests/test_layernorm_cuda.py tests/test_residual_cuda.py
run backbone tests/test_backbone_host.py -m gpu
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_n.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_r01_n.json; tail -3 gpurun_out/bench_err.log
