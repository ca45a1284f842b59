#!/bin/bash
# full-metric capture of k_makeindices on the 0.25-degree grid (its own gpurun call: one profiler per call)
set -x
timeout 300 python profiles/setup_kernels.py C4 > gpurun_out/r2d_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_makeindices -s 1 -c 1 -f -o gpurun_out/prof_makeindices_c4 python profiles/setup_kernels.py C4 > gpurun_out/r2d_ncu.log 2>&1; tail -2 gpurun_out/r2d_ncu.log
