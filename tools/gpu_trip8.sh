#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary8.txt
run() { tag=$1; shift; timeout 900 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary8.txt; grep -E "^stage [0-3] \{|^cfg1_stage2 \{|passed|failed|Error|timed out" "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-700 | tee -a gpurun_out/summary8.txt; }
run bf16 tests/test_cuda_parity.py -k "bf16"
run rest tests/test_cuda_parity.py -k "not bf16"
run backbone tests/test_backbone_host.py -m gpu
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks8.log 2>&1; head -4 gpurun_out/time_blocks8.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_d.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_r01_d.json; tail -3 gpurun_out/bench_err.log
