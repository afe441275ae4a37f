#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary13.txt
run() { tag=$1; shift; timeout 900 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary13.txt; grep -E "passed|failed|Error|timed out|^E  " "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-300 | tee -a gpurun_out/summary13.txt; }
run dw tests/test_dwconv_cuda.py
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/plain13.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches13.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/ncu13.log 2>&1; echo "[ncu] exit $?"
