#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointwise_cuda.py tests/test_cuda_parity.py -q -x --tb=short 2>&1 | tail -3
timeout 300 python tools/time_gemm_abi.py 2>&1 | grep -v wgrad | grep -E "s0|s2" | tee gpurun_out/time_gemm_persistent2.log
