"""Diagnostic: the step path under Contract R (the reference's constants, K = 4) at scale, with the exact and the fast
diversity window -- the exact mode scans a 1,000-entry reward window per env and step (4 KB) and is the HBM-heavy case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pool = [torch.rand((n, 2), device=dev) * 2 - 1 for _ in range(8)]
for name, over in (("R exact window", dict(diversity_mode=A.DIV_EXACT)), ("R fast window", dict(diversity_mode=A.DIV_FAST))):
    eng = BatchedEngine(n, A.default_config(A.CONTRACT_R, autoreset=1, **over), device=0)
    eng.reset()
    for b in range(1100):     # past 1,000 pushes: the window is full
        eng.step(pool[b % 8], want_final=False)
    torch.cuda.synchronize()
    K = 40
    st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for k in range(K):
        flush.zero_()
        st[k].record(); eng.step(pool[k % 8], want_final=False); en[k].record()
    torch.cuda.synchronize()
    ms = sum(s.elapsed_time(e) for s, e in zip(st, en)) / K
    win = 4000 if "exact" in name else 0
    print(f"CR {name}: {ms:.4f} ms/step  {n / (ms * 1e-3):.3e} env-steps/s  algorithmic {(294 + win) * n / ms / 1e6:.0f} GB/s", flush=True)
    eng.close()
