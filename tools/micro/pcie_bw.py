"""Diagnostic: pinned-host <-> device copy bandwidth of the box at the e2e transfer sizes (copy engine, CUDA events).

Single process: GPU 0 alone.  Under torchrun (`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1
tools/micro/pcie_bw.py [--bind]`): every rank copies on its own GPU AT THE SAME TIME (barrier before each measurement), which
is the ceiling of the end-to-end step at N ranks of one host -- the ranks share the host's memory and PCIe root bandwidth.
--bind pins each rank to its GPU's local CPUs first (tvc_ai_b200.dist.bind_to_gpu_cpus) so the pinned buffers are first-touched
on that NUMA node."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from tvc_ai_b200 import dist as D

rank, world, local = D.init_from_env("nccl")
torch.cuda.set_device(local)
bind = D.bind_to_gpu_cpus(local) if "--bind" in sys.argv else {"bound": False}
dev = torch.device("cuda", local)
for mb in (2, 12, 24, 128):
    nbytes = mb << 20
    h = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = torch.tensor([nbytes / ms / 1e6], dtype=torch.float64, device=dev)
        if world > 1:
            lo, hi = gbs.clone(), gbs.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN), dist.all_reduce(hi, op=dist.ReduceOp.MAX), dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
            if rank == 0:
                print(f"PCIE {name} {mb:4d} MiB x {world} ranks at once: per rank {lo.item():6.1f} .. {hi.item():6.1f} GB/s, "
                      f"aggregate {gbs.item():7.1f} GB/s  (bind: {bind})", flush=True)
        else:
            print(f"PCIE {name} {mb:4d} MiB: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.1f} GB/s  (bind: {bind})", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
