#!/bin/bash
# round 2, GPU run 18 (1 GPU): packed kernel with paired steps — parity, dimension sweep, wide rows A/B
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_index.py tests/test_early_stop.py tests/test_vs_reference.py tests/test_fuzz.py -m gpu -x -q > gpurun_out/r2_pytest_packed.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_packed.log
timeout 900 python tools/dim_bench.py 64,96,128,192,256,100,130,200 > gpurun_out/r2_dim_bench7.json 2> gpurun_out/r2_dim_bench7.err; cat gpurun_out/r2_dim_bench7.json
timeout 900 python tools/dim_bench.py 260,300,1000,1280,2000,4000 > gpurun_out/r2_dim_bench7_any.json 2> gpurun_out/r2_dim_bench7_any.err; cat gpurun_out/r2_dim_bench7_any.json
FFX_OPT_kernel=3 timeout 900 python tools/dim_bench.py 260,300,1000,1280,2000,4000 > gpurun_out/r2_dim_bench7_packed.json 2> gpurun_out/r2_dim_bench7_packed.err; cat gpurun_out/r2_dim_bench7_packed.json
