#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary25.txt
run() { tag=$1; shift; timeout 900 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary25.txt; grep -E "passed|failed|Error|timed out|^E  " "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-300 | tee -a gpurun_out/summary25.txt; }
run pw tests/test_pointwise_cuda.py
run parity tests/test_cuda_parity.py
run backbone tests/test_backbone_host.py -m gpu
timeout 300 python tools/time_gemm_abi.py 2>&1 | grep wgrad | tee gpurun_out/time_wgrad_db.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_l.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_r01_l.json; tail -3 gpurun_out/bench_err.log
