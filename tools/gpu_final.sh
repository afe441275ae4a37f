#!/bin/bash
# Round-end validation: all GPU tests, smoke(), default bench (with CPU baseline), reference arm, 2-GPU bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final_1gpu.err; echo "[bench] exit $?"; cut -c1-220 gpurun_out/bench_final_1gpu.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>/dev/null; echo "[ref] exit $?"; cut -c1-200 gpurun_out/bench_final_ref.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_final_2gpu.json 2> gpurun_out/bench_final_2gpu.err; echo "[bench 2gpu] exit $?"; grep metric gpurun_out/bench_final_2gpu.json | cut -c1-220
fi
