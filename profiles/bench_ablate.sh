#!/bin/bash
# usage: profiles/bench_ablate.sh TAG "variants"  -> kernel ms per OTMB_V4_VARIANT, base run in between
tag=$1
run() {
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${tag}.json 2>gpurun_out/bench_${tag}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}.json')); print('$1 kernel_ms %.4f sm_mhz %s' % (d['kernel_ms'], d['clocks']['sm_mhz']))" || tail -3 gpurun_out/bench_${tag}.err
}
run "base"
for v in $2; do OTMB_V4_VARIANT=$v run "variant $v"; done
run "base"
for v in $2; do OTMB_V4_VARIANT=$v run "variant $v"; done
