#!/bin/bash
# round 2, GPU run 11 (1 GPU)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_full.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_pytest_gpu_full.log
python bench.py --workload c5_sharded_maxp --emulate-shards 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_emul8.json 2> gpurun_out/r2_bench_c5_emul8.err; echo "c5emul rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c5_emul8.json').read()); print(d['ms_per_step'], d['value'], d['roofline']['achieved'])"
python bench.py --steps 5 --warmup 3 --no-api --no-cpu-baseline > gpurun_out/r2_bench_c3_quick.json 2> gpurun_out/r2_bench_c3_quick.err; echo "c3 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c3_quick.json').read()); print(d['ms_per_step'], d['value'], d['roofline']['achieved'], d['e2e']['ms_per_step'])"
python bench.py --workload c2_msmarco_passage --steps 10 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "c2 rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c2.json').read()); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['achieved'])"
timeout 600 python tools/dim_bench.py 128,256,100 > gpurun_out/r2_dim_bench5.json 2> gpurun_out/r2_dim_bench5.err; cat gpurun_out/r2_dim_bench5.json
