"""Diagnostic: pinned-host <-> device copy bandwidth of the box at the e2e transfer sizes (copy engine, CUDA events)."""
import torch

dev = torch.device("cuda", 0)
for mb in (2, 12, 24, 128):
    nbytes = mb << 20
    h = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"PCIE {name} {mb:4d} MiB: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.1f} GB/s", flush=True)
