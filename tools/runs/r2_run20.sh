#!/bin/bash
# round 2, GPU run 20 (1 GPU): full GPU suite after the packed kernel; dimension sweep of record
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_gpu_final2.log
timeout 900 python tools/dim_bench.py 64,96,128,192,256,100,130,200 > gpurun_out/r2_dim_bench9_short.json 2> gpurun_out/r2_dim_bench9.err; cat gpurun_out/r2_dim_bench9_short.json
timeout 900 python tools/dim_bench.py 260,300,1000,1280,2000,4000 > gpurun_out/r2_dim_bench9_any.json 2>> gpurun_out/r2_dim_bench9.err; cat gpurun_out/r2_dim_bench9_any.json
timeout 900 python tools/dim_bench.py 384,768,1024,2048,3072,4096 > gpurun_out/r2_dim_bench9_uniform.json 2>> gpurun_out/r2_dim_bench9.err; cat gpurun_out/r2_dim_bench9_uniform.json
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_check.json 2> gpurun_out/r2_bench_c3_check.err; echo "c3 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c3_check.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('api_e2e'))"
