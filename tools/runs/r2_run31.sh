#!/bin/bash
# round 2, GPU run 31 (1 GPU): ADC table wait deferred to the first look-up — parity and the C4 line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_full_size.py tests/test_fuzz.py tests/test_quantizer.py tests/test_early_stop.py -m gpu -x -q -k "adc or opq or pq or c4 or C4 or quantiz" > gpurun_out/r2_pytest_adc.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_adc.log
python bench.py --workload c4_opq_avep --steps 20 --warmup 5 > gpurun_out/r2_bench_c4_v5.json 2> gpurun_out/r2_bench_c4_v5.err; echo "c4 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4_v5.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'])"
timeout 300 python tools/adc_determinism.py 2>&1 | tail -3
