#!/bin/bash
# round 2, GPU run 26 (1 GPU): ncu of the packed kernel at D = 64 and D = 130
set -x
mkdir -p gpurun_out
for D in 64 130; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffx_score -s 3 -c 1 -f -o gpurun_out/r2_dim${D}_v3 \
      python tools/dim_bench.py $D > gpurun_out/r2_ncu_dim${D}.log 2>&1; echo "ncu $D rc=$?"
done
