#!/bin/bash
mkdir -p gpurun_out
echo "== generic table path"; DAT_B200_ATTN_BWD_GENERIC_TABLE=1 timeout 600 python -m pytest tests/test_cuda_parity.py -q -x --tb=line -k "backward_bf16 or attention" 2>&1 | tail -4
echo "== fast table path"; timeout 600 python -m pytest tests/test_cuda_parity.py -q -x --tb=line -k "backward_bf16" 2>&1 | tail -6
