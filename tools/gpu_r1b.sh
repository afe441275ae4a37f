#!/bin/bash
# GPU trip: all GPU tests, kernel rooflines (incl. block kernels K1-K3), config-5 sweep.
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15) 2>&1
timeout 600 python tools/kernel_rooflines.py > gpurun_out/kernel_rooflines.md 2> gpurun_out/kernel_rooflines.err; echo "[rooflines] exit $?"; tail -3 gpurun_out/kernel_rooflines.err
grep -E "K1|K2|K3|block" gpurun_out/kernel_rooflines.md
timeout 600 python tools/sweep_dattn.py > gpurun_out/sweep_dattn.md 2> gpurun_out/sweep_dattn.err; echo "[sweep] exit $?"; tail -3 gpurun_out/sweep_dattn.err
tail -8 gpurun_out/sweep_dattn.md
