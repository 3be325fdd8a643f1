#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_fuzz.py tests/test_sharded.py tests/test_early_stop.py -m gpu -x -q > gpurun_out/r2_pytest_run14.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_pytest_run14.log
timeout 900 python tools/dim_bench.py 64,96,128,192,256,100,130,200,300 > gpurun_out/r2_dim_bench6.json 2> gpurun_out/r2_dim_bench6.err; cat gpurun_out/r2_dim_bench6.json; tail -3 gpurun_out/r2_dim_bench6.err
