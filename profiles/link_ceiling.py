"""Bare host-link ceiling of the box: pinned H2D alone, pinned D2H alone, and both directions at once (two streams),
per rank and aggregated over the ranks of a torchrun launch.  bench.py's e2e moves 260 MB in and 714 MB out per step."""
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_in, n_out = 260 << 20, 714 << 20
h_in = torch.empty(n_in, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n_out, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(do_in, do_out, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        if do_in:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        best = min(best, time.perf_counter() - t)
    return best


for name, a, b in (("H2D 260 MiB alone", True, False), ("D2H 714 MiB alone", False, True), ("both at once", True, True)):
    t = run(a, b)
    nbytes = (n_in if a else 0) + (n_out if b else 0)
    if rank == 0:
        print(f"ranks={world} {name}: {t * 1e3:.2f} ms  per-rank {nbytes / t / 1e9:.1f} GB/s  aggregate {world * nbytes / t / 1e9:.1f} GB/s", flush=True)
if world > 1:
    dist.destroy_process_group()
