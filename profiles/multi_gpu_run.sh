#!/bin/bash
# usage (gpurun --gpus N): bash profiles/multi_gpu_run.sh N  -> batch (contract line) and sharded C4 numbers at N GPUs
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
timeout 600 bash -c "$(declare -f run); N=$N; run --steps 30 --warmup 5 --no-cpu-baseline" > gpurun_out/bench_batch_v4n_$N.json 2> gpurun_out/bench_batch_v4n_$N.err
tail -c 900 gpurun_out/bench_batch_v4n_$N.json; echo
timeout 900 bash -c "$(declare -f run); N=$N; run --workload C4 --mode sharded --steps 10 --warmup 3 --no-cpu-baseline --no-e2e" > gpurun_out/bench_sh_c4_v4n_$N.json 2> gpurun_out/bench_sh_c4_v4n_$N.err
tail -c 900 gpurun_out/bench_sh_c4_v4n_$N.json; echo
if [ "$N" = "2" ]; then
  OTMB_SHARDED_BACKEND=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/_sharded_worker.py 2>&1 | tail -3
fi
