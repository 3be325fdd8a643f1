#!/bin/bash
# round 2, GPU run 29 (1 GPU): tiled ADC table kernel — parity and the C4 line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_fuzz.py -m gpu -x -q -k "adc or opq or pq or c4 or C4" > gpurun_out/r2_pytest_adc.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_adc.log
python bench.py --workload c4_opq_avep --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_v4.json 2> gpurun_out/r2_bench_c4_v4.err; echo "c4 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4_v4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_c4_v4.csv python bench.py --workload c4_opq_avep --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_c4_v4.log 2>&1; echo "ncu rc=$?"
grep -i "lut\|adc_xor_kernel\|rotate" gpurun_out/r2_launches_c4_v4.csv | tail -6 | cut -c1-260
