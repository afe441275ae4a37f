#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary6.txt
run() { tag=$1; shift; timeout 600 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary6.txt; grep -E "passed|failed|Error|error" "gpurun_out/pytest_${tag}.log" | tail -6 | tee -a gpurun_out/summary6.txt; }
run all tests/test_cuda_parity.py
timeout 300 python tools/gemm_timing.py > gpurun_out/gemm_timing.log 2>&1; cat gpurun_out/gemm_timing.log
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks6.log 2>&1; head -4 gpurun_out/time_blocks6.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_c.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-400 gpurun_out/bench_r01_c.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r01_c.json')); print(d['roofline'])"
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01_tcfwd.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "[ncu launches] exit $?"
