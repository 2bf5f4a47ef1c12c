"""Multi-GPU plumbing: env slabs shard across ranks with no data-path collective; the only
exchange is a sum all-reduce of the fixed-length episode-statistics vector (SURVEY.md section
8(e)).  One process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in CPU tests)."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _abi as A


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK/WORLD_SIZE/MASTER_* when WORLD_SIZE > 1."""
    rank, world, local = rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_cpus(device_index: int) -> dict:
    """Pin this process to the CPU cores next to GPU `device_index` (its PCIe root's NUMA node) BEFORE any pinned host
    buffer is allocated, so that the first-touch policy places the e2e result slabs (46 B per env and step over PCIe) in
    that node's memory and the copies do not cross the socket interconnect.  Reads the GPU's PCI address from torch and
    /sys/bus/pci/devices/<addr>/local_cpulist; a box that reports one node for every GPU (or hides sysfs) is left alone.
    Returns what was found / done (for the bench line)."""
    info = {"bound": False}
    try:
        prop = torch.cuda.get_device_properties(device_index)
        addr = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{addr}"
        with open(f"{base}/local_cpulist") as f:
            cpulist = f.read().strip()
        try:
            with open(f"{base}/numa_node") as f:
                info["numa_node"] = int(f.read().strip())
        except OSError:
            pass
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        want = cpus & allowed
        info.update(pci=addr, local_cpus=len(cpus), allowed_cpus=len(allowed))
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            info["bound"] = True
    except (OSError, AttributeError, ValueError, RuntimeError) as exc:
        info["error"] = type(exc).__name__
    return info


def slab(total_envs: int, rank: int, world: int):
    """Contiguous env slab [base, base + count) owned by `rank`; global ids feed the Philox
    counters so trajectories do not depend on the GPU count."""
    per = total_envs // world
    rem = total_envs % world
    count = per + (1 if rank < rem else 0)
    base = rank * per + min(rank, rem)
    return base, count


def allreduce_stats(stats, group=None):
    """Sum-reduce a TVC_NUM_STATS float64 vector over ranks.  Accepts a CUDA tensor (NCCL, stays on
    the device, async on the current stream), a CPU tensor or a numpy array (gloo)."""
    if isinstance(stats, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(stats, np.float64))
    else:
        t = stats
    if t.numel() != A.NUM_STATS:
        raise ValueError(f"statistics vector must have {A.NUM_STATS} entries")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def stats_dict(vec) -> dict:
    v = vec.detach().cpu().numpy() if isinstance(vec, torch.Tensor) else np.asarray(vec)
    return dict(zip(A.STAT_NAMES, v.tolist()))
