#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary37.txt
run() { tag=$1; shift; timeout 900 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary37.txt; grep -E "passed|failed|Error|timed out|^E  " "gpurun_out/pytest_${tag}.log" | sort | uniq -c | tail -8 | cut -c1-300 | tee -a gpurun_out/summary37.txt; }
run all tests -m gpu
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01_p.json 2> gpurun_out/bench_err.log; echo "[bench] exit $?"; cut -c1-230 gpurun_out/bench_r01_p.json; tail -3 gpurun_out/bench_err.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r01_p_2gpu.json 2> gpurun_out/bench2_err.log; echo "[bench 2gpu] exit $?"; cut -c1-230 gpurun_out/bench_r01_p_2gpu.json; tail -3 gpurun_out/bench2_err.log
