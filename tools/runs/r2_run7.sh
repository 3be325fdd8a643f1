#!/bin/bash
# round 2, GPU run 7 (1 GPU)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_index.py tests/test_fuzz.py tests/test_early_stop.py -m gpu -x -q -k "adc or any_dimension or fuzz or early or skew or long_list or cached or quantized" > gpurun_out/r2_pytest_run7.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_run7.log
python bench.py --workload c4_opq_avep --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/r2_bench_c4.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'])"
timeout 1200 python tools/dim_bench.py 1000,1280,2000,4000 > gpurun_out/r2_dim_bench3.json 2> gpurun_out/r2_dim_bench3.err; echo "dim rc=$?"
cat gpurun_out/r2_dim_bench3.json
FFX_API_PROFILE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_api.json 2> gpurun_out/r2_bench_c3_api.err; echo "bench rc=$?"
tail -14 gpurun_out/r2_bench_c3_api.err
cat gpurun_out/r2_bench_c3_api.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d.get('api_e2e'),indent=1)); print(d['ms_per_step'], d['e2e'])"
