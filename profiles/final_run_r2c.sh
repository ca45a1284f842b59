#!/bin/bash
# full-metric capture of k_fused_v4 (its own gpurun call: one profiler per call)
set -x
timeout 200 python profiles/prof_driver.py fused 3 > gpurun_out/r2c_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fused_v4 -s 2 -c 1 -f -o gpurun_out/prof_fused_v5d python profiles/prof_driver.py fused 3 > gpurun_out/r2c_ncu_f.log 2>&1; tail -2 gpurun_out/r2c_ncu_f.log
