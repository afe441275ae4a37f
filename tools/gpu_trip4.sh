#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary4.txt
run() { tag=$1; shift; timeout 600 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary4.txt; grep -E "^stage|passed|failed|Error|error" "gpurun_out/pytest_${tag}.log" | tail -12 | tee -a gpurun_out/summary4.txt; }
run tc_attn tests/test_cuda_parity.py -k "attention_core_fwd_tensor_core"
run bf16 tests/test_cuda_parity.py -k "bf16"
run rest tests/test_cuda_parity.py -k "not bf16 and not attention_core_fwd_tensor_core"
run backbone tests/test_backbone_host.py -m gpu
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks4.log 2>&1; echo "time exit $?" | tee -a gpurun_out/summary4.txt
cat gpurun_out/time_blocks4.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_b.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cat gpurun_out/bench_r01_b.json; tail -3 gpurun_out/bench_err.log
