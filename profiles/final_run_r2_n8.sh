#!/bin/bash
# Round 2, final 8-GPU sequence (gpurun --gpus 8): native sharded worker against the oracle, then the contract line
# (batch value + the sharded C4 record) exactly as the driver launches it.  Every step under its own timeout.
N=${1:-8}
OTMB_SHARDED_BACKEND=native timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tests/_sharded_worker.py 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2b_bench_n$N.json 2> gpurun_out/r2b_bench_n$N.err
tail -c 1500 gpurun_out/r2b_bench_n$N.json; echo; tail -3 gpurun_out/r2b_bench_n$N.err
