#!/bin/bash
# round 2, GPU run 9 (1 GPU): the whole -m gpu suite, then the bench lines of record (C3 default with api_e2e and the
# reference CPU baseline, C4, C2, C1)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_pytest_gpu_full.log
python bench.py --workload c4_opq_avep --steps 10 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/r2_bench_c4.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c4.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['roofline']['frac'])"
python bench.py --workload c2_msmarco_passage --steps 10 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "c2 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c2.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_full.json 2> gpurun_out/r2_bench_c3_full.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_c3_full.err
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c3_full.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value']); print(json.dumps(d['api_e2e'])[:600])"
