#!/bin/bash
# round 2, GPU run 3 (2 GPUs): multi-device API tests, the 2-GPU peer-memory test, bench at N=2 (weak + strong + sharded)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_device.py tests/test_sharded.py -m gpu -x -q > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
print(json.dumps(d.get('strong'), indent=1))
print(json.dumps(d.get('sharded'), indent=1))
PY
