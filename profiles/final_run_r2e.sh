#!/bin/bash
# launch list of the set-up kernels on the 0.25-degree grid (durations under ncu: cold cache, serialised)
timeout 300 python profiles/setup_kernels.py C4 > gpurun_out/r2e_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2e_launches_setup_c4.csv python profiles/setup_kernels.py C4 > gpurun_out/r2e_ncu.log 2>&1
grep -E "k_makeindices|k_metrics3d|k_faceflux_tile" gpurun_out/r2e_launches_setup_c4.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | head -20
