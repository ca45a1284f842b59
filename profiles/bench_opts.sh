#!/bin/bash
# usage: profiles/bench_opts.sh TAG "0 1 2 3"   -> kernel ms per OTMB_V4_OPT (variant from OTMB_V4_VARIANT)
tag=$1; shift
for m in $1; do
  OTMB_V4_OPT=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${tag}_o$m.json 2>gpurun_out/bench_${tag}_o$m.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_o$m.json')); print('opt $m kernel_ms %.4f step_ms %.4f frac %.3f' % (d['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
done
