#!/bin/bash
# round 2, GPU run 6 (1 GPU): ADC kernel changes (parity, C4 bench, ncu), shared-memory look-up peak, dimension sweep, API profile
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_disk.py tests/test_fuzz.py -m gpu -x -q -k "adc or any_dimension or disk or fuzz or load or wide" > gpurun_out/r2_pytest_adc.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest_adc.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/smem_lookup_peak tools/smem_lookup_peak.cu && /tmp/smem_lookup_peak > gpurun_out/r2_smem_lookup_peak.json; echo "smem rc=$?"; cat gpurun_out/r2_smem_lookup_peak.json
python bench.py --workload c4_opq_avep --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/r2_bench_c4.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline'])"
timeout 1200 python tools/dim_bench.py 100,130,300,1000,1280,2000,4000 > gpurun_out/r2_dim_bench2.json 2> gpurun_out/r2_dim_bench2.err; echo "dim rc=$?"
cat gpurun_out/r2_dim_bench2.json
FFX_API_PROFILE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_api.json 2> gpurun_out/r2_bench_c3_api.err; echo "bench rc=$?"
tail -40 gpurun_out/r2_bench_c3_api.err
cat gpurun_out/r2_bench_c3_api.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps(d.get('api_e2e'),indent=1)); print(d['ms_per_step'], d['e2e'])"
python bench.py --workload c4_opq_avep --steps 1 --warmup 3 > gpurun_out/r2_plain_c4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ffx_adc_xor -s 3 -c 1 -o gpurun_out/r2_adc_xor_full \
    python bench.py --workload c4_opq_avep --steps 1 --warmup 3 > gpurun_out/r2_ncu_c4.log 2>&1
echo "ncu rc=$?"
