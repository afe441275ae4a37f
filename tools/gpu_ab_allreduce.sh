#!/bin/bash
# A/B of the overlapped gradient all-reduce at 2 GPUs (debug).  Results: gpurun_out/ab_allreduce.txt
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" timeout 200 $TR --master-port 29520 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$name', d['ms_per_step'], d['value'])" | tee -a gpurun_out/ab_allreduce.txt; }
run overlap_nocap DAT_B200_NCCL_MAX_CTAS=0
run skip_ar DAT_B200_BENCH_SKIP_AR=1
run overlap_cta8 A=1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('single', d['ms_per_step'], d['value'])" | tee -a gpurun_out/ab_allreduce.txt
