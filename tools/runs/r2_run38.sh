#!/bin/bash
# round 2, GPU run 38 (1 GPU): score array sized by the candidate count (two 8-warp CTAs per SM at C3) — shapes, then parity
set -x
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --steps 10 0:0:0:0 0:14:0:0 0:16:0:0 0:12:0:0 2>&1 | tail -5
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_sharded.py tests/test_fuzz.py tests/test_full_size.py -m gpu -x -q > gpurun_out/r2_pytest_smax.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_smax.log
