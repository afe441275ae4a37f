#!/bin/bash
mkdir -p gpurun_out
for st in 2 0; do
python tools/run_stage_kernels.py $st fwdbwd > gpurun_out/plain_stage${st}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_stage${st}_fwdbwd.csv python tools/run_stage_kernels.py $st fwdbwd > gpurun_out/ncu_stage${st}.log 2>&1
echo "ncu stage $st exit $?"
done
DAT_B200_DISABLE_TC=1 timeout 600 python -m pytest tests/test_cuda_parity.py -q -rA --tb=line -k "full_size_vs_oracle" 2>&1 | grep -E "^stage" | cut -c1-700
