#!/bin/bash
# round 2, GPU run 16 (1 GPU): ncu --set full of the short-row kernels (D = 128 lane-major, D = 100 tree-as-data),
# plus the live reference comparison on the coded route
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vs_reference.py -m gpu -x -q 2>&1 | tail -3
for D in 128 100; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffx_score -s 3 -c 1 -f -o gpurun_out/r2_dim${D} \
      python tools/dim_bench.py $D > gpurun_out/r2_ncu_dim${D}.log 2>&1; echo "ncu $D rc=$?"
done
ls -la gpurun_out/*.ncu-rep
