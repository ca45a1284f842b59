#!/bin/bash
# Round 2: the final 1-GPU test / bench / ncu sequence (run under gpurun from the repository root)
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 400 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -c 300 gpurun_out/r2_bench_ref.json
timeout 300 python bench.py --workload C3 --steps 20 --warmup 5 --mode batch --no-cpu-baseline > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; tail -c 300 gpurun_out/r2_bench_c3.json
# launch list of the device-resident timed region (cold-cache, serialised: shares, not absolutes)
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --mode batch > gpurun_out/r2_plain_l.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --mode batch > gpurun_out/r2_ncu_l.log 2>&1; tail -1 gpurun_out/r2_ncu_l.log
# the dense stencil kernel with bulk-copy staging, full metric set
timeout 100 python profiles/setup_kernels.py C2 > gpurun_out/r2_setup_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_faceflux_tile -s 1 -c 1 -f -o gpurun_out/r2_prof_faceflux_tile python profiles/setup_kernels.py C2 > gpurun_out/r2_ncu_ft.log 2>&1; tail -1 gpurun_out/r2_ncu_ft.log
