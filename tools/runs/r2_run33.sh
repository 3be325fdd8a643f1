#!/bin/bash
# round 2, GPU run 33 (1 GPU): ncu --set full of the shipped C4 scoring kernel (tables built ahead)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:ffx_adc_xor_kernel -s 4 -c 1 -f -o gpurun_out/r2_adc_xor_final \
    python bench.py --workload c4_opq_avep --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_adc_final.log 2>&1; echo "ncu rc=$?"
