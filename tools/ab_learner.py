"""Where the batched SAC learner's time goes (config 5: 4,096 envs, batch 4,096, 8 updates per graph replay): variants of the
optimiser / matmul precision, ms per update.  Usage: python tools/ab_learner.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine
from tvc_ai_b200.replay import DeviceReplay
from tvc_ai_b200 import sac as S

n, T = 4096, 8
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset()


def run(tag, tf32=False, fused=False, batch=4096):
    cfg = S.SACConfig(batch_size=batch, learning_starts=1, tf32=tf32, fused_adam=fused)
    rp = DeviceReplay(n, T, capacity=1 << 19, device=0, seed=0, reward_scale=cfg.reward_scale)
    L = S.SACLearner(rp, cfg, updates_per_replay=8)
    for _ in range(4):
        rp.collect(eng, L.weights())
    for _ in range(3):
        L.update()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.update()
    e1.record()
    torch.cuda.synchronize()
    print(f"LEARNER {tag:28s} {e0.elapsed_time(e1) / 80:.4f} ms per update (batch {batch})", flush=True)


run("baseline fp32 / foreach Adam")
run("tf32 matmul", tf32=True)
run("fused Adam", fused=True)
run("tf32 + fused Adam", tf32=True, fused=True)
run("tf32 + fused, batch 256", tf32=True, fused=True, batch=256)
eng.close()
