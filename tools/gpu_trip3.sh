#!/bin/bash
# usage: gpu_trip3.sh  (tests by group, each in its own process with a timeout; then timings)
mkdir -p gpurun_out; rm -f gpurun_out/summary3.txt
run() { tag=$1; shift; timeout 600 python -m pytest "$@" -q -rA --tb=short > "gpurun_out/pytest_${tag}.log" 2>&1; echo "[$tag] exit $?" | tee -a gpurun_out/summary3.txt; tail -4 "gpurun_out/pytest_${tag}.log" | tee -a gpurun_out/summary3.txt; }
run tc_gemm tests/test_cuda_parity.py -k "tensor_core"
run fwd tests/test_cuda_parity.py -k "block_forward_fp32 or attention_core or offset_net"
run bwd tests/test_cuda_parity.py -k "block_backward_fp32"
run bf16 tests/test_cuda_parity.py -k "bf16"
run full tests/test_cuda_parity.py -k "full_size"
run backbone tests/test_backbone_host.py -m gpu
timeout 600 python tools/time_blocks.py 16 > gpurun_out/time_blocks3.log 2>&1; echo "time exit $?" | tee -a gpurun_out/summary3.txt
cat gpurun_out/time_blocks3.log
