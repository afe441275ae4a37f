#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary7.txt
run() { tag=$1; shift; timeout 900 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary7.txt; grep -E "^stage|^cfg1|passed|failed|Error|timed out" "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -14 | cut -c1-900 | tee -a gpurun_out/summary7.txt; }
run bwd_full tests/test_cuda_parity.py -k "full_size_vs_oracle"
run bf16 tests/test_cuda_parity.py -k "bf16 and not full_size_vs_oracle"
run rest tests/test_cuda_parity.py -k "not bf16"
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks7.log 2>&1; head -4 gpurun_out/time_blocks7.log
