#!/bin/bash
set -x
mkdir -p gpurun_out
for C in 8 3; do
FFX_SHARD_CHUNKS=$C timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --workload c5_sharded_maxp --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_n2_$C.json 2> gpurun_out/r2_bench_c5_n2_$C.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench_c5_n2_$C.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_c5_n2_$C.json').read().strip().splitlines()[-1])
print('chunks $C', d['ms_per_step'], d['e2e']['ms_per_step'], d['sharded']['split_ms'])"
done
