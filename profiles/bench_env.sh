#!/bin/bash
# usage: profiles/bench_env.sh TAG "ENV=val ..." ... -> kernel ms per environment setting (each run twice, interleaved)
tag=$1; shift
run() {
  env $1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_${tag}.json 2>gpurun_out/bench_${tag}.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}.json')); print('$1 kernel_ms %.4f' % (d['kernel_ms']))" || tail -3 gpurun_out/bench_${tag}.err
}
for rep in 1 2; do for e in "$@"; do run "$e"; done; done
