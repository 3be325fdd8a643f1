#!/bin/bash
# round 2, GPU run 21 (1 GPU): 32-chain trees with two chains per lane — parity and sweep
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_index.py tests/test_fuzz.py -m gpu -x -q > gpurun_out/r2_pytest_packed2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_packed2.log
timeout 900 python tools/dim_bench.py 260,300,312,400,500,1000 > gpurun_out/r2_dim_bench10_any.json 2> gpurun_out/r2_dim_bench10.err; cat gpurun_out/r2_dim_bench10_any.json
