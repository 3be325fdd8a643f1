#!/bin/bash
# round 2, GPU run 28 (1 GPU): index load with the id columns adopted beside the row stream; shard tests on packed dims
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_disk.py tests/test_index.py tests/test_sharded.py tests/test_h5.py tests/test_quantizer.py -m gpu -x -q > gpurun_out/r2_pytest_load.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_load.log
timeout 900 python tools/h5_load_bench.py > gpurun_out/r2_h5_load.json 2> gpurun_out/r2_h5_load.err; echo "load rc=$?"; cat gpurun_out/r2_h5_load.json; tail -3 gpurun_out/r2_h5_load.err
