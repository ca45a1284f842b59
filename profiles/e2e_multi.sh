#!/bin/bash
# usage (gpurun --gpus N): bash profiles/e2e_multi.sh N -> e2e ms per step at N ranks: default fetch, 2 / 4 threads, direct copies
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline; }
for cfg in "default" "OTMB_HOST_THREADS=2" "OTMB_HOST_THREADS=4" "OTMB_FETCH_DIRECT=1"; do
  if [ "$cfg" = "default" ]; then run > gpurun_out/e2e_m.json 2> gpurun_out/e2e_m.err; else env $cfg bash -c "$(declare -f run); N=$N; run" > gpurun_out/e2e_m.json 2> gpurun_out/e2e_m.err; fi
  python -c "
import json; d=json.load(open('gpurun_out/e2e_m.json')); print('$cfg', 'N=$N e2e ms', d['e2e']['ms_per_step'], 'kernel', d['kernel_ms'])" || tail -3 gpurun_out/e2e_m.err
done
