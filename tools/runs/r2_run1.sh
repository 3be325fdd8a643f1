#!/bin/bash
# round 2, GPU run 1 (1 GPU): smoke, the full-size oracle parity tests, the default bench line, C1 with
# the L2 flush, then the ncu launch list and one full capture of the fused kernel of the same command
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests/test_full_size.py -m gpu -x -q > gpurun_out/r2_full_size.log 2>&1; echo "full_size rc=$?"
tail -5 gpurun_out/r2_full_size.log
python bench.py --steps 10 --warmup 3 --no-api > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench rc=$?"
python bench.py --workload c1_passage_10k --steps 10 --warmup 3 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; echo "c1 rc=$?"
python bench.py --workload c5_sharded_maxp --emulate-shards 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_emul8.json 2> gpurun_out/r2_bench_c5_emul8.err; echo "c5emul rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_c3.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:ffx_score_tma -s 3 -c 1 -o gpurun_out/r2_score_tma_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-api > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"
