#!/bin/bash
# round 2, GPU run 22 (1 GPU): radix sort with overlapped match.any — parity at full size, C4 / C3 lines
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_full_size.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_pytest_radix.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_radix.log
python bench.py --workload c4_opq_avep --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_v3.json 2> gpurun_out/r2_bench_c4_v3.err; echo "c4 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4_v3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 10 --warmup 3 --no-api --no-cpu-baseline > gpurun_out/r2_bench_c3_v3.json 2> gpurun_out/r2_bench_c3_v3.err; echo "c3 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c3_v3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_c4_v3.csv python bench.py --workload c4_opq_avep --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_c4_v3.log 2>&1; echo "ncu rc=$?"
grep -c adc gpurun_out/r2_launches_c4_v3.csv
