"""Host->device copy ceiling of the box with N concurrent ranks.

    python tools/h2d_ceiling.py                                  # 1 rank
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py [--no-bind]

Every rank owns ONE pinned host buffer (default 160 MB, the size of one e2e step's inputs of bench.py) and copies it to
its GPU with one plain cudaMemcpyAsync per repetition (torch ``copy_(non_blocking=True)`` of a contiguous pinned
tensor), all ranks at the same time between barriers.  Prints one JSON line: per-rank and aggregate GB/s, with
and without binding the rank to its GPU's NUMA-local cores before the pinned allocation, plus D2H.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rag_b200 import dist as D  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=160.0)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--no-bind", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind = {"bound": False} if a.no_bind else D.bind_to_gpu_numa(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n = int(a.mb * 1e6) // 4
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    host.normal_()                                   # first touch on this (possibly bound) thread
    devb = torch.empty(n, dtype=torch.float32, device=dev)
    back = torch.empty(n // 8, dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        barrier()
        return ms

    h2d_ms = timed(lambda: devb.copy_(host, non_blocking=True))
    d2h_ms = timed(lambda: back.copy_(devb[: n // 8], non_blocking=True))
    both = torch.tensor([h2d_ms, d2h_ms], dtype=torch.float64, device=dev)
    alls = [torch.zeros_like(both) for _ in range(world)]
    if world > 1:
        dist.all_gather(alls, both)
    else:
        alls = [both]
    if rank == 0:
        h2d = [n * 4 / (float(t[0]) * 1e-3) / 1e9 for t in alls]
        d2h = [(n // 8) * 4 / (float(t[1]) * 1e-3) / 1e9 for t in alls]
        print(json.dumps({"what": "concurrent pinned H2D ceiling, one cudaMemcpyAsync per rank per rep", "n_ranks": world, "mb": a.mb,
                          "bound_to_gpu_numa": bind, "h2d_GBps_per_rank": [round(v, 2) for v in h2d],
                          "h2d_GBps_aggregate_at_slowest": round(world * min(h2d), 1), "h2d_GBps_sum": round(sum(h2d), 1),
                          "d2h_GBps_per_rank": [round(v, 2) for v in d2h], "host_cpus": os.cpu_count()}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
