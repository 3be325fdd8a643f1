#!/bin/bash
# round 2, GPU run 39 (1 GPU): one consolidated dimension sweep on the final tree; ncu of the packed kernel at D = 384
set -x
mkdir -p gpurun_out
timeout 900 python tools/dim_bench.py 64,96,128,192,256,384,512,768,1024,2048,4096,100,200,300,1000 > gpurun_out/r2_dim_bench_final.json 2> gpurun_out/r2_dim_bench_final.err; cat gpurun_out/r2_dim_bench_final.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:ffx_score -s 3 -c 1 -f -o gpurun_out/r2_dim384 \
    python tools/dim_bench.py 384 > gpurun_out/r2_ncu_dim384.log 2>&1; echo "ncu rc=$?"
