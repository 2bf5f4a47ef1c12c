"""A/B timing of the step path for one build of the library (TVC_B200_LIB selects it): the bench.py workload
(262,144 envs, Contract X, K=10, autoreset, burn-in), CUDA events per step, L2 flushed between steps.
Usage: TVC_B200_LIB=/path/lib.so python tools/ab_step.py [tag] [envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

tag = sys.argv[1] if len(sys.argv) > 1 else ""
n = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
K = int(sys.argv[3]) if len(sys.argv) > 3 else 60
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
st = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
en = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
for k in range(K):
    flush.zero_()
    st[k].record()
    eng.step(pool[k % 16], want_final=False)
    en[k].record()
torch.cuda.synchronize()
per = sorted(s.elapsed_time(e) for s, e in zip(st, en))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    eng.step(pool[k % 16], want_final=False)
e1.record()
torch.cuda.synchronize()
stats = eng.stats(reset_after=False)
print(f"AB {tag}: cold-L2 mean {sum(per) / K:.4f} ms  median {per[K // 2]:.4f}  min {per[0]:.4f}  warm {e0.elapsed_time(e1) / K:.4f} ms  "
      f"episodes {stats[0]:.0f} sum_return {stats[1]:.6e}", flush=True)
eng.close()
