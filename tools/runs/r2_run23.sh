#!/bin/bash
# round 2, GPU run 23 (2 GPUs): multi-device and peer-memory tests after the kernel changes; first-call profile of the API
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multi_device.py tests/test_sharded.py -m gpu -x -q > gpurun_out/r2_pytest_2gpu_v2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_2gpu_v2.log
FFX_API_PROFILE=first python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_firstprof.json 2> gpurun_out/r2_bench_c3_firstprof.err; echo "c3 rc=$?"
grep -A30 "Ordered by" gpurun_out/r2_bench_c3_firstprof.err | head -40
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-api --no-cpu-baseline > gpurun_out/r2_bench_n2_v2.json 2> gpurun_out/r2_bench_n2_v2.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2_v2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
print(d['strong']['ms_per_step'], d['sharded']['ms_per_step'], d['sharded']['split_ms'], d['sharded']['verified'], d['sharded']['e2e']['ms_per_step'])
PY
