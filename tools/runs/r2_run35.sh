#!/bin/bash
# round 2, GPU run 35 (4 GPUs): the driver's scaling line at N = 4
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29535 \
    bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
print(d['strong']['value'], d['strong']['ms_per_step'])
s=d['sharded']; print(s['value'], s['ms_per_step'], s['split_ms'], s['verified'], s['per_gpu_hbm_gbs'], s['e2e']['ms_per_step'])
PY
