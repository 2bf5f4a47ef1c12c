#!/usr/bin/env python
"""Driver for ncu captures of the step path in its steady state: `burn` untimed steps (episodes desynchronised: the mix of
free flight, ground contact and resets bench.py times), then `steps` more.  Usage:
  ncu --set full --clock-control none --import-source on -k regex:step_kernel_v2 --launch-skip <burn> --launch-count 1 \
      -o gpurun_out/step python tools/steady_steps.py 262144 400 4"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
burn = int(sys.argv[2]) if len(sys.argv) > 2 else 400
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset()
acts = [torch.rand((n, 2), device="cuda") * 2 - 1 for _ in range(8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for t in range(burn):
    eng.step(acts[t % 8], want_final=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = 0.0
for t in range(steps):
    flush.zero_()
    e0.record()
    eng.step(acts[t % 8], want_final=False)
    e1.record()
    torch.cuda.synchronize()
    ms += e0.elapsed_time(e1)
print(f"steady_steps: {n} envs, {steps} steps after {burn}, {ms / max(steps, 1):.4f} ms/step (cold L2)")
eng.close()
