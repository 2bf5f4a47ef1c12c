"""Diagnostic (library built with -DTVC_DBG): time one step restricted to the near-ground groups, then one restricted to the
airborne groups, between normal steps.  Usage: TVC_B200_LIB=variants/dbg.so python tools/ab_split.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset()
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
pool = [torch.rand((n, 2), generator=gen, device=dev) * 2 - 1 for _ in range(16)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for b in range(410):
    eng.step(pool[b % 16], want_final=False)
torch.cuda.synchronize()
res = {"-1": [], "0": [], "1": []}
for rep in range(12):
    for mode in ("-1", "0", "1"):
        os.environ["TVC_DBG_ONLY"] = mode
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.step(pool[rep % 16], want_final=False)
        e1.record()
        torch.cuda.synchronize()
        res[mode].append(e0.elapsed_time(e1))
    os.environ["TVC_DBG_ONLY"] = "-1"
    for b in range(3):
        eng.step(pool[b % 16], want_final=False)
for k, v in res.items():
    v = sorted(v)
    print(f"SPLIT n={n} mode {k}: median {v[len(v) // 2]:.4f} ms  min {v[0]:.4f}", flush=True)
eng.close()
