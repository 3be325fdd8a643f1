#!/bin/bash
# round 2, GPU run 32 (8 GPUs): the drop-in API on every GPU of the box from one process (devices=[0..7])
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_multi_device.py tests/test_sharded.py -m gpu -x -q > gpurun_out/r2_pytest_8gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_8gpu.log
