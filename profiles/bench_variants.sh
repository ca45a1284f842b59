#!/bin/bash
# usage: profiles/bench_variants.sh TAG "0 1 2"   -> kernel ms per OTMB_V3_VARIANT
tag=$1; shift
for m in $1; do
  OTMB_V4_VARIANT=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${tag}_$m.json 2>gpurun_out/bench_${tag}_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}_$m.json')); print('variant $m kernel_ms %.4f step_ms %.4f frac %.3f' % (d['kernel_ms'], d['ms_per_step'], d['roofline']['frac']))"
done
