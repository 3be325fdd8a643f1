#!/bin/bash
# round 2, GPU run 15 (1 GPU): the record run — full GPU suite, default bench line (api_e2e + cpu
# baseline + reference arm), C4 / C2 / C1 lines, launch list of the default command
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_final.json 2> gpurun_out/r2_bench_c3_final.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c3_final.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['roofline'], d['e2e'], d['clocks'])
print(d.get('api_e2e')); print(d['cpu_baseline'])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref_final.json
for w in c4_opq_avep c2_msmarco_passage c1_passage_10k; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_${w}_final.json 2> gpurun_out/r2_bench_${w}_final.err; echo "$w rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${w}_final.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline'])"
done
python bench.py --workload c5_sharded_maxp --emulate-shards 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_emul8_final.json 2> gpurun_out/r2_bench_c5_emul8_final.err; echo "c5emul rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_c5_emul8_final.json').read()); print(d['ms_per_step'], d['value'], d['roofline']['achieved'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3_final.csv \
    python bench.py --steps 2 --warmup 3 --no-api --no-cpu-baseline > gpurun_out/r2_ncu_c3_final.log 2>&1; echo "ncu rc=$?"
