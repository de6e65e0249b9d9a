#!/usr/bin/env python3
"""Aggregate host->device copy ceiling of one box: every rank copies a page-locked 1 GiB buffer to its GPU in a loop,
all ranks at once (barrier on both sides), plain cudaMemcpyAsync through torch -- no library code.  bench.py's e2e line
at N GPUs is bound by this number (its timed region copies 8 bytes per sample from host memory).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py
Prints one JSON line on rank 0: per-rank and aggregate GB/s, alone (one rank at a time) and all together."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cpus = sorted(os.sched_getaffinity(0))
    per = len(cpus) // world
    if per >= 1:
        os.sched_setaffinity(0, cpus[local * per:(local + 1) * per])       # first touch of the pinned buffer from this rank's own cores
    n = 1 << 28                                                          # 1 GiB of float32
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    host.fill_(1.0)
    dev = torch.empty(n, dtype=torch.float32, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(reps):
        t0 = time.perf_counter()
        for _ in range(reps):
            dev.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n * 4 / (time.perf_counter() - t0) / 1e9

    run(2)
    alone = torch.zeros(world, dtype=torch.float64, device="cuda")
    for r in range(world):                                               # one rank at a time
        barrier()
        if r == rank:
            alone[r] = run(5)
        barrier()
    barrier()
    t0 = time.perf_counter()
    mine = run(8)                                                        # all ranks together
    barrier()
    wall = time.perf_counter() - t0
    together = torch.zeros(world, dtype=torch.float64, device="cuda")
    together[rank] = mine
    if world > 1:
        dist.all_reduce(alone)
        dist.all_reduce(together)
    if rank == 0:
        print(json.dumps({"ranks": world, "host_cpus": len(cpus), "alone_GBps": [round(v, 2) for v in alone.tolist()],
                          "together_GBps": [round(v, 2) for v in together.tolist()], "aggregate_GBps": round(world * 8 * n * 4 / wall / 1e9, 2),
                          "msamples_per_s_ceiling_fc32": round(world * 8 * n * 4 / wall / 8 / 1e6, 1),
                          "how": "1 GiB page-locked buffer per rank, torch copy_(non_blocking) x 8, barrier + synchronize on both sides"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
