#!/bin/bash
# round 2, GPU run 36 (1 GPU): full GPU suite + smoke + default bench line on the final tree
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1]); print(d['steps'], d['warmup'], d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks'], d['api_e2e']['second_call_s'], d['cpu_baseline']['value'])"
