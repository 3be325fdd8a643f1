#!/bin/bash
# round 2, GPU run 19 (1 GPU): packed kernel, several chains per lane — parity, dimension sweep, ncu of D = 128
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_index.py tests/test_early_stop.py tests/test_vs_reference.py tests/test_fuzz.py -m gpu -x -q > gpurun_out/r2_pytest_packed.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2_pytest_packed.log
timeout 900 python tools/dim_bench.py 64,96,128,192,256,100,130,200 > gpurun_out/r2_dim_bench8.json 2> gpurun_out/r2_dim_bench8.err; cat gpurun_out/r2_dim_bench8.json
tail -3 gpurun_out/r2_dim_bench8.err
for D in 128 100; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffx_score -s 3 -c 1 -f -o gpurun_out/r2_dim${D}_v2 \
      python tools/dim_bench.py $D > gpurun_out/r2_ncu_dim${D}.log 2>&1; echo "ncu $D rc=$?"
done
