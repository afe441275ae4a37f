#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pointwise_cuda.py tests/test_cuda_parity.py -q -x --tb=short 2>&1 | tail -4
timeout 300 python tools/time_gemm_abi.py 2>&1 | tee gpurun_out/time_gemm_2cta.log
