#!/bin/bash
# round 2, GPU run 24 (1 GPU): A/B of D = 384 / 512 through the packed kernel; API first call after the has_queries fix
set -x
mkdir -p gpurun_out
timeout 600 python tools/ab_packed_wide.py 2>&1 | tail -10
timeout 600 python tools/dim_bench.py 384,512 2>/dev/null
FFX_OPT_kernel=3 timeout 600 python tools/dim_bench.py 384,512 2>/dev/null
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_api_v4.json 2> gpurun_out/r2_bench_c3_api_v4.err; echo "c3 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c3_api_v4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('api_e2e'))"
