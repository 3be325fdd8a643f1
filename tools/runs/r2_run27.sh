#!/bin/bash
# round 2, GPU run 27 (1 GPU): ncu of the ADC table kernel at C4
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:adc_xor_lut -s 4 -c 1 -f -o gpurun_out/r2_adc_lut \
    python bench.py --workload c4_opq_avep --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_adc_lut.log 2>&1; echo "ncu rc=$?"
