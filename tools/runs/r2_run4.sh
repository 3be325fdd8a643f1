#!/bin/bash
# round 2, GPU run 4 (2 GPUs): multi-device API tests, the 2-GPU peer-memory test, then the 1-GPU bench line with api_e2e
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_device.py tests/test_sharded.py -m gpu -x -q > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_2gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_api.json 2> gpurun_out/r2_bench_c3_api.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_c3_api.err
cat gpurun_out/r2_bench_c3_api.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d.get('api_e2e'),indent=1)); print(d['ms_per_step'], d['e2e'])"
