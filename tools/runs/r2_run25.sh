#!/bin/bash
# round 2, GPU run 25 (1 GPU): D = 384 / 512 through the packed kernel by default — parity, A/B, sweep
set -x
mkdir -p gpurun_out
timeout 600 python tools/ab_packed_wide.py 2>&1 | tail -10
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_index.py tests/test_fuzz.py tests/test_early_stop.py -m gpu -x -q > gpurun_out/r2_pytest_packed3.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_packed3.log
timeout 600 python tools/dim_bench.py 384,512 2>/dev/null
FFX_OPT_kernel=2 timeout 600 python tools/dim_bench.py 384,512 2>/dev/null
