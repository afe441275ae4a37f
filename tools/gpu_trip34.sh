#!/bin/bash
mkdir -p gpurun_out
echo "== generic table path"; DAT_B200_ATTN_BWD_GENERIC_TABLE=1 timeout 600 python -m pytest tests/test_cuda_parity.py -q -x --tb=short -k "backward_bf16" 2>&1 | grep -E "^E|passed|failed|timed out" | head -12 | cut -c1-300
