#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary5.txt
run() { tag=$1; shift; timeout 600 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary5.txt; grep -E "passed|failed|Error|error" "gpurun_out/pytest_${tag}.log" | tail -6 | tee -a gpurun_out/summary5.txt; }
run all tests/test_cuda_parity.py
run backbone tests/test_backbone_host.py -m gpu
timeout 300 python tools/gemm_timing.py > gpurun_out/gemm_timing.log 2>&1; cat gpurun_out/gemm_timing.log
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks5.log 2>&1; cat gpurun_out/time_blocks5.log
