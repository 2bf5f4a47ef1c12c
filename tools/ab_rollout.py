"""A/B timing of the fused rollout (config 4: 65,536 envs, T = 64, 2x256 actor) for one build of the library
(TVC_B200_LIB selects it).  Usage: TVC_B200_LIB=/path/lib.so python tools/ab_rollout.py [tag] [envs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

tag = sys.argv[1] if len(sys.argv) > 1 else ""
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = 64
dev = torch.device("cuda", 0)
torch.manual_seed(0)
nn = torch.nn
net = nn.Sequential(nn.Linear(10, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 4)).to(dev)
w = dict(w1=net[0].weight.detach(), b1=net[0].bias.detach(), w2=net[2].weight.detach(), b2=net[2].bias.detach(),
         w3=net[4].weight.detach(), b3=net[4].bias.detach())
ro = BatchedEngine(nr, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
ro.reset()
for _ in range(4):
    ro.rollout(w, T)
torch.cuda.synchronize()
reps = 24
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for a, b in ev:
    a.record()
    out = ro.rollout(w, T)
    b.record()
torch.cuda.synchronize()
per = sorted(a.elapsed_time(b) for a, b in ev)
ms = per[len(per) // 2]
st = ro.stats()
print(f"ABR {tag}: median {ms:.3f} (min {per[0]:.3f}, max {per[-1]:.3f}) ms per {T}-step launch of {nr} envs -> {nr * T / (ms * 1e-3):.3e} env-steps/s  episodes {st[0]:.0f} sum_return {st[1]:.6e}", flush=True)
ro.close()
