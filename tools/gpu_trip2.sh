#!/bin/bash
mkdir -p gpurun_out
for grp in "bf16" ; do
  tag=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_cuda_parity.py -q -rA --tb=short -k "$grp" > "gpurun_out/pytest_${tag}.log" 2>&1
  echo "[$grp] exit $?" | tee -a gpurun_out/summary2.txt; tail -2 "gpurun_out/pytest_${tag}.log" | tee -a gpurun_out/summary2.txt
done
timeout 900 python -m pytest tests/test_backbone_host.py -q -rA --tb=short -m gpu > gpurun_out/pytest_backbone.log 2>&1
echo "[backbone] exit $?" | tee -a gpurun_out/summary2.txt; tail -3 gpurun_out/pytest_backbone.log | tee -a gpurun_out/summary2.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_err.log
echo "[bench] exit $?" | tee -a gpurun_out/summary2.txt; cat gpurun_out/bench_r01_a.json; tail -5 gpurun_out/bench_err.log
timeout 600 python tools/profile_step.py > gpurun_out/plain_profile_step.log 2>&1 &&
timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r01_simt.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "[ncu launches] exit $?" | tee -a gpurun_out/summary2.txt
