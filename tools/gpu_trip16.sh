#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dwconv_cuda.py -q --tb=short 2>&1 | tail -5
timeout 300 python tools/time_dwconv3_abi.py 0 2>&1 | tee gpurun_out/time_dwconv3_abi2.log
