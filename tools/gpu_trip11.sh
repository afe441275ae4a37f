#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary11.txt
run() { tag=$1; shift; timeout 900 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary11.txt; grep -E "passed|failed|Error|timed out|^E  " "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-300 | tee -a gpurun_out/summary11.txt; }
run pw tests/test_pointwise_cuda.py
run backbone tests/test_backbone_host.py -m gpu
run parity tests/test_cuda_parity.py
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_r01_f.json; tail -3 gpurun_out/bench_err.log
timeout 600 python tools/time_blocks.py > gpurun_out/time_blocks11.log 2>&1; tail -6 gpurun_out/time_blocks11.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/plain11.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches11.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu11.log 2>&1; echo "[ncu] exit $?"
