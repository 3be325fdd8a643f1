#!/bin/bash
# round 2, GPU run 34 (1 GPU): ncu --set full of the shipped C3 kernel (after the live-candidate masks)
set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:ffx_score_tma_kernel -s 4 -c 1 -f -o gpurun_out/r2_score_tma_final \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r2_ncu_c3_final2.log 2>&1; echo "ncu rc=$?"
