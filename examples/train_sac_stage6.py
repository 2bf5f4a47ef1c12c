#!/usr/bin/env python
"""BASELINE config 5: curriculum "stage 6" (max noise, actuator delay, thrust-curve variation) with an end-to-end SAC
loop on the batched CUDA env -- everything stays on the GPU (no host round trip per transition).

What it mirrors: the reference's train loop (scripts/train.py:406-533: act -> env.step -> agent.update, periodic
evaluation, curriculum update), with the batch-of-1 host loop replaced by the fused rollout kernel (the actor is evaluated on tensor
cores inside the env loop and the kernel writes the transitions itself), an on-device replay buffer and a plain PyTorch SAC
learner (SURVEY.md section 8(f) rank 1 in its simplest form).  The learner is
PyTorch on purpose: only the env path is this repo's product.

    python examples/train_sac_stage6.py --envs 4096 --iters 200

Reports env-steps/s end to end and the learner's share of the wall time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.nn as nn
import torch.nn.functional as F

from tvc_ai_b200 import RocketTVCVectorEnv
from tvc_ai_b200.curriculum import stage6_conditions
from tvc_ai_b200.evaluate import evaluate


def mlp(i, o):
    return nn.Sequential(nn.Linear(i, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, o))


class Actor(nn.Module):
    """Legacy SAC actor shape (10-256-256-4 -> mean, log_std), the same network tvc_rollout evaluates in-kernel."""

    def __init__(self):
        super().__init__()
        self.net = mlp(10, 4)

    def forward(self, obs, deterministic=False):
        out = self.net(obs)
        mean, log_std = out[:, :2], out[:, 2:].clamp(-20, 2)
        if deterministic:
            return torch.tanh(mean), None
        std = log_std.exp()
        u = mean + std * torch.randn_like(mean)
        a = torch.tanh(u)
        logp = (-0.5 * ((u - mean) / std) ** 2 - log_std - 0.9189385).sum(-1) - torch.log(1 - a * a + 1e-6).sum(-1)
        return a, logp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--buffer", type=int, default=1 << 20)
    ap.add_argument("--rollout-steps", type=int, default=8, help="env steps per fused-rollout launch")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--eager-learner", action="store_true", help="run the SAC update eagerly instead of as a CUDA graph")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(args.seed)

    # stage 6 = stage_5 conditions + sensor noise + actuator delay (3 control steps) + thrust-curve variation
    env = RocketTVCVectorEnv(args.envs, config={"globals": {"seed": args.seed}}, contract="X", device=0, final_info=False,
                             delay_steps=3, thrust_curve=1, propellant_fraction=0.2, cg_burn_shift=0.05)
    env.set_curriculum(stage6_conditions())
    env.reset(seed=args.seed, options={"return_torch": True})

    actor, q1, q2 = Actor().to(dev), mlp(12, 1).to(dev), mlp(12, 1).to(dev)
    q1t, q2t = mlp(12, 1).to(dev), mlp(12, 1).to(dev)
    q1t.load_state_dict(q1.state_dict()), q2t.load_state_dict(q2.state_dict())
    opt_a = torch.optim.Adam(actor.parameters(), lr=3e-4, capturable=True)
    opt_q = torch.optim.Adam(list(q1.parameters()) + list(q2.parameters()), lr=3e-4, capturable=True)
    alpha, gamma, tau, rscale = 0.2, 0.99, 0.005, 0.01

    # one SAC update on a static batch; captured into a CUDA graph below (an eager update is ~300 small launches and
    # took 4.4 ms, i.e. 95 % of the device time of this loop)
    bs = args.batch
    sb = dict(s=torch.zeros((bs, 10), device=dev), a=torch.zeros((bs, 2), device=dev), r=torch.zeros(bs, device=dev),
              s2=torch.zeros((bs, 10), device=dev), d=torch.zeros(bs, device=dev))

    def sac_update():
        s, a, r, sn, d = sb["s"], sb["a"], sb["r"], sb["s2"], sb["d"]
        with torch.no_grad():
            an, lpn = actor(sn)
            qn = torch.min(q1t(torch.cat([sn, an], 1)), q2t(torch.cat([sn, an], 1))).squeeze(-1) - alpha * lpn
            y = r + gamma * (1 - d) * qn
        sa = torch.cat([s, a], 1)
        lq = F.mse_loss(q1(sa).squeeze(-1), y) + F.mse_loss(q2(sa).squeeze(-1), y)
        opt_q.zero_grad(set_to_none=False), lq.backward(), opt_q.step()
        ap_, lp = actor(s)
        sap = torch.cat([s, ap_], 1)
        la = (alpha * lp - torch.min(q1(sap), q2(sap)).squeeze(-1)).mean()
        opt_a.zero_grad(set_to_none=False), la.backward(), opt_a.step()
        with torch.no_grad():
            for p, pt in zip(list(q1.parameters()) + list(q2.parameters()), list(q1t.parameters()) + list(q2t.parameters())):
                pt.mul_(1 - tau).add_(p, alpha=tau)

    update_graph = None
    if not args.eager_learner:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    sac_update()          # warm-up on zeros (allocates the Adam state; negligible for the run)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            update_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(update_graph):
                sac_update()
        except Exception as exc:  # noqa: BLE001 -- fall back to the eager learner
            print(f"CUDA-graph capture of the SAC update failed ({exc}); running it eagerly", file=sys.stderr)
            update_graph = None

    B = args.buffer
    T = args.rollout_steps
    buf = dict(s=torch.zeros((B, 10), device=dev), a=torch.zeros((B, 2), device=dev), r=torch.zeros(B, device=dev),
               s2=torch.zeros((B, 10), device=dev), d=torch.zeros(B, device=dev))
    # the fused rollout kernel acts with the actor's current weights for T steps per launch and writes the transitions
    # itself (tvc_rollout_io.obs_all / next_obs_all / ...): no host round trip and no per-step torch forward for acting
    n = args.envs
    tr = dict(obs=torch.zeros((T, n, 10), device=dev), actions=torch.zeros((T, n, 2), device=dev),
              reward=torch.zeros((T, n), device=dev), next_obs=torch.zeros((T, n, 10), device=dev),
              terminated=torch.zeros((T, n), dtype=torch.uint8, device=dev), truncated=torch.zeros((T, n), dtype=torch.uint8, device=dev))
    eng = env.engine
    head, filled = 0, 0
    t_env = t_learn = 0.0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for it in range(args.iters):
        ev[0].record()
        lin = [m for m in actor.net if isinstance(m, nn.Linear)]
        w = dict(w1=lin[0].weight.detach(), b1=lin[0].bias.detach(), w2=lin[1].weight.detach(), b2=lin[1].bias.detach(),
                 w3=lin[2].weight.detach(), b3=lin[2].bias.detach())
        eng.rollout(w, T, transitions=tr)
        m = T * n
        idx = (head + torch.arange(m, device=dev)) % B
        buf["s"][idx], buf["a"][idx] = tr["obs"].reshape(m, 10), tr["actions"].reshape(m, 2)
        buf["r"][idx], buf["s2"][idx] = tr["reward"].reshape(m) * rscale, tr["next_obs"].reshape(m, 10)
        buf["d"][idx] = tr["terminated"].reshape(m).float()
        head, filled = (head + m) % B, min(filled + m, B)
        ev[1].record()
        # ---- SAC updates (one per rollout step) ----
        for _ in range(T):
            j = torch.randint(0, filled, (args.batch,), device=dev)
            for k_ in ("s", "a", "r", "s2", "d"):
                torch.index_select(buf[k_], 0, j, out=sb[k_])
            if update_graph is not None:
                update_graph.replay()
            else:
                sac_update()
        ev[2].record()
        torch.cuda.synchronize()
        t_env += ev[0].elapsed_time(ev[1])
        t_learn += ev[1].elapsed_time(ev[2])
    wall = time.perf_counter() - wall0
    stats = env.episode_stats()
    env.close()
    pol = lambda o: actor(o, deterministic=True)[0]   # noqa: E731
    evalm = evaluate(pol, episodes=64, contract="X", conditions=stage6_conditions(), delay_steps=3, thrust_curve=1)
    print(json.dumps({
        "config": "stage 6: wind 3 N, mass +-30 %, initial tilt 0.7 rad, sensor noise 0.02, actuator delay 3 steps, thrust curve",
        "envs": args.envs, "iters": args.iters, "rollout_steps": T, "env_steps": args.envs * args.iters * T,
        "env_steps_per_sec_end_to_end": args.envs * args.iters * T / wall, "acting": "tvc_rollout (tcgen05 actor in-kernel)",
        "env_ms_per_iter": t_env / args.iters, "learner_ms_per_iter": t_learn / args.iters,
        "learner_share_of_device_time": t_learn / (t_env + t_learn),
        "learner": "PyTorch SAC update, " + ("CUDA graph replay" if update_graph is not None else "eager"),
        "episodes": stats["episodes"], "train_success_rate": stats["successes"] / max(stats["episodes"], 1),
        "eval": evalm}))


if __name__ == "__main__":
    main()
