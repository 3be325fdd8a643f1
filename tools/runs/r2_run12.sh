#!/bin/bash
# round 2, GPU run 12 (1 GPU): build-side additions (sgemm, OPQ on device, coalescing) + the quick regression set
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_pq_build.py tests/test_util.py tests/test_quantizer.py tests/test_ranking.py tests/test_index.py -m gpu -x -q > gpurun_out/r2_pytest_run12.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/r2_pytest_run12.log
python tools/pq_build_bench.py > gpurun_out/r2_pq_build_bench.log 2>&1; tail -5 gpurun_out/r2_pq_build_bench.log
