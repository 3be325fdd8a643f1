#!/bin/bash
# round 2, GPU run 17 (1 GPU): the packed short-row kernel — parity, dimension sweep, ncu; C4 / C1 record lines
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_index.py tests/test_early_stop.py tests/test_vs_reference.py -m gpu -x -q > gpurun_out/r2_pytest_packed.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_packed.log
timeout 900 python tools/dim_bench.py 64,96,128,192,256,100,130,200 > gpurun_out/r2_dim_bench6.json 2> gpurun_out/r2_dim_bench6.err; cat gpurun_out/r2_dim_bench6.json
for w in c4_opq_avep c1_passage_10k; do
  python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/r2_bench_${w}_final.json 2> gpurun_out/r2_bench_${w}_final.err; echo "$w rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/r2_bench_${w}_final.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e'], d['roofline'])"
done
for D in 128 100; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffx_score -s 3 -c 1 -f -o gpurun_out/r2_dim${D} \
      python tools/dim_bench.py $D > gpurun_out/r2_ncu_dim${D}.log 2>&1; echo "ncu $D rc=$?"
done
ls -la gpurun_out/*.ncu-rep
