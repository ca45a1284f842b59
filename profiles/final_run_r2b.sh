#!/bin/bash
# Round 2, after the paired division: 1-GPU bench lines, launch list and the full-metric capture of k_fused_v4
set -x
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; tail -c 300 gpurun_out/r2b_bench_n1.json
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --mode batch > gpurun_out/r2b_plain_l.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --mode batch > gpurun_out/r2b_ncu_l.log 2>&1; tail -1 gpurun_out/r2b_ncu_l.log
