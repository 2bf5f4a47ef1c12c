"""Row S14 for the batch: tvc_curiosity (tcgen05 forward model + fused epilogue, csrc/tvc_curiosity.cu) against the torch path it
replaces (fp32 nn.Sequential forward + the elementwise kernels around it), CUDA events, one step of `n` envs.
Usage: python tools/ab_curiosity.py [envs] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine
from tvc_ai_b200.env import CuriosityModule

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset()
torch.manual_seed(0)
fm = CuriosityModule(obs_dim=8, action_dim=2, device="cuda").forward_model
acts = torch.rand((n, 2), device="cuda") * 2 - 1
for _ in range(50):
    eng.step(acts, want_final=True)
prev = eng.obs[:, :8].clone()
has = torch.ones(n, dtype=torch.uint8, device="cuda")
out, intr = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
eng.curiosity(acts, prev, has, out, intrinsic=intr, forward_model=fm)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ms = []
    for _ in range(K):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return sum(ms) / len(ms), ms[len(ms) // 2]


def kernel():
    eng.curiosity(acts, prev, has, out, intrinsic=intr, forward_model=None)


hb = has.bool()


def torch_path():
    with torch.no_grad():
        done = (eng.terminated | eng.truncated).bool()
        a = acts.clamp(-1.0, 1.0)
        nxt = torch.where(done[:, None], eng.final_obs[:, :8], eng.obs[:, :8])
        pred = fm(torch.cat([prev, a], dim=1))
        i = 0.01 * ((pred - nxt) ** 2).mean(dim=1)
        r = eng.reward + torch.where(hb, i, torch.zeros_like(i))
        prev.copy_(eng.obs[:, :8]); hb.copy_(~done)
    return r


km, kmed = timed(kernel)
tm, tmed = timed(torch_path)
flops = 2.0 * n * (16 * 256 + 256 * 256 + 256 * 16)
print(f"CURIOSITY {n} envs: tvc_curiosity {km:.4f} ms (median {kmed:.4f}; {flops / km / 1e9:.0f} TFLOP/s of bf16 MMA work incl. padding)  "
      f"torch fp32 path {tm:.4f} ms (median {tmed:.4f})  speed-up {tm / km:.1f}x", flush=True)
eng.close()
