"""Multi-GPU plumbing around the hot path.  One process per GPU; stereo pairs
are sharded across ranks; the four kernels need no exchange step, so there is NO data-path
collective.  NCCL is used only for (i) the DDP gradient all-reduce of the trainable units and
(ii) one all-reduce of the 7-element metric vector per evaluation (rag_b200.metrics.MetricAccumulator).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_pairs(n_pairs: int, rank: int, world: int) -> range:
    """Contiguous, balanced shard of ``n_pairs`` stereo pairs for ``rank`` (sizes differ by at most 1)."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def wrap_ddp(model: torch.nn.Module, local_rank: int):
    """(Re)wrap the growable network in DDP.  Must be called again after every ``expand``/``select``
    because the reference adds, deletes and freezes parameters per task (rag_model.py:403-517,725-829);
    ``find_unused_parameters`` because only the units named by ``task_arch`` run in a given step."""
    from torch.nn.parallel import DistributedDataParallel as DDP

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    if any(p.requires_grad for p in model.parameters()):
        dev = [local_rank] if torch.cuda.is_available() else None
        return DDP(model, device_ids=dev, find_unused_parameters=True)
    return model


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin the calling process to the CPU cores that are local to GPU ``local_rank`` (NVML's ideal CPU affinity =
    the cores of the NUMA node / socket the GPU's PCIe root hangs off).  Call it BEFORE allocating pinned host
    buffers: pinned pages are placed on the NUMA node of the thread that first touches them, and a host->device
    copy from the remote socket crosses the inter-socket link (with 8 ranks that link, not PCIe, becomes the
    end-to-end limit).  Returns what was done; never raises (no NVML / no permission = unbound)."""
    info = {"bound": False}
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(local_rank)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, cpus=len(allowed), first=allowed[0], last=allowed[-1])
        try:
            info["numa_node"] = int(nv.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
    except Exception as e:  # noqa: BLE001
        info["error"] = f"{type(e).__name__}: {e}"[:120]
    return info
