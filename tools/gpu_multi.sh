#!/bin/bash
# Multi-GPU evidence on N GPUs of one box.   usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh N <what...>'
#   what: bench (bench.py at N, overlap on and off), ar (all-reduce alone), c3 (UperNet + DAT-S++ DDP step), c4 (config-4 inference)
N=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for what in "$@"; do
  case $what in
    bench)
      timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "[bench $N] exit $?"; cut -c1-200 gpurun_out/bench_${N}gpu.json; tail -2 gpurun_out/bench_${N}gpu.err
      DAT_B200_BENCH_OVERLAP=0 timeout 400 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu_nooverlap.json 2> gpurun_out/bench_${N}gpu_nooverlap.err; echo "[bench $N no-overlap] exit $?"; cut -c1-200 gpurun_out/bench_${N}gpu_nooverlap.json;;
    bench1)
      timeout 400 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "[bench $N] exit $?"; grep '^{' gpurun_out/bench_${N}gpu.json | cut -c1-200; tail -2 gpurun_out/bench_${N}gpu.err;;
    ar)
      timeout 200 $TR --master-port 29513 tools/time_allreduce.py > gpurun_out/allreduce_${N}gpu.json 2> gpurun_out/allreduce_${N}gpu.err; echo "[ar $N] exit $?"; cat gpurun_out/allreduce_${N}gpu.json
      DAT_B200_NCCL_MAX_CTAS=8 timeout 200 $TR --master-port 29514 tools/time_allreduce.py >> gpurun_out/allreduce_${N}gpu.json 2>> gpurun_out/allreduce_${N}gpu.err; tail -1 gpurun_out/allreduce_${N}gpu.json;;
    c3)
      timeout 400 $TR --master-port 29515 tools/bench_upernet.py --model small --batch 2 --steps 8 --warmup 3 > gpurun_out/upernet_small_b2_${N}gpu.json 2> gpurun_out/upernet_${N}gpu.err; echo "[c3 b2 $N] exit $?"; cat gpurun_out/upernet_small_b2_${N}gpu.json; tail -2 gpurun_out/upernet_${N}gpu.err
      timeout 400 $TR --master-port 29516 tools/bench_upernet.py --model small --batch 16 --steps 8 --warmup 3 > gpurun_out/upernet_small_b16_${N}gpu.json 2>> gpurun_out/upernet_${N}gpu.err; echo "[c3 b16 $N] exit $?"; cat gpurun_out/upernet_small_b16_${N}gpu.json;;
    c4)
      timeout 400 $TR --master-port 29517 tools/bench_config4.py > gpurun_out/config4_${N}gpu.json 2> gpurun_out/config4_${N}gpu.err; echo "[c4 $N] exit $?"; cat gpurun_out/config4_${N}gpu.json; tail -2 gpurun_out/config4_${N}gpu.err;;
  esac
done
