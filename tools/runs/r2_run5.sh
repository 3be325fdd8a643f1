#!/bin/bash
# round 2, GPU run 5 (1 GPU): any-dimension kernel parity, dimension sweep, api_e2e after the pool / encoder changes
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "any_dimension or bit_exact_all_modes or shards" > gpurun_out/r2_pytest_any.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_any.log
timeout 1200 python tools/dim_bench.py 100,128,130,256,300,768,1000,1280,2000,4000 > gpurun_out/r2_dim_bench.json 2> gpurun_out/r2_dim_bench.err; echo "dim rc=$?"
cat gpurun_out/r2_dim_bench.json; tail -3 gpurun_out/r2_dim_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_api.json 2> gpurun_out/r2_bench_c3_api.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_c3_api.err
cat gpurun_out/r2_bench_c3_api.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d.get('api_e2e'),indent=1)); print(d['ms_per_step'], d['e2e'])"
