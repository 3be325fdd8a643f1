#!/bin/bash
# round 2, GPU run 8 (1 GPU)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_index.py -m gpu -x -q -k "adc or early or golden" > gpurun_out/r2_pytest_run8.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_run8.log
python bench.py --workload c4_opq_avep --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/r2_bench_c4.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['roofline'])"
timeout 1200 python tools/dim_bench.py 1280,2000 > gpurun_out/r2_dim_bench4.json 2> gpurun_out/r2_dim_bench4.err; echo "dim rc=$?"
cat gpurun_out/r2_dim_bench4.json
python bench.py --workload c4_opq_avep --steps 2 --warmup 3 > gpurun_out/r2_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_c4.csv \
    python bench.py --workload c4_opq_avep --steps 2 --warmup 3 > gpurun_out/r2_ncu_c4_launches.log 2>&1
echo "ncu rc=$?"
grep -E "adc_xor|rotate" gpurun_out/r2_launches_c4.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -30
