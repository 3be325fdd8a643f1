#!/bin/bash
# round 2, GPU run 10 (8 GPUs): the driver's scaling line at N = 8 (weak + strong + sharded C5)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"
tail -5 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d.get('strong'), indent=1))
s=d.get('sharded'); print(json.dumps({k:v for k,v in s.items() if k not in ('exchange','kernel')}, indent=1))
PY
