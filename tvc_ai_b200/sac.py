"""Batched SAC learner fed by the on-device replay ring (BASELINE config 5).

The reference updates on a batch of ONE transition per env step (scripts/train.py:574-591 -> agent/multi_algorithm_agent.py:
950-1016 `_update_sac`).  With the env at 1e9 steps/s that loop is the bottleneck by seven orders of magnitude; this module is
the same update rule (twin critics with min-target, `gamma (1 - done)` bootstrapping, entropy weight 0.2, Polyak tau) on
batches gathered by `tvc_replay_sample`, captured ONCE into a CUDA graph (gather -> critic step -> actor step -> target
update, `updates_per_replay` times) so that an update costs a graph replay instead of ~300 eager launches.

* The actor is the legacy 2x256 SAC shape (10-256-256-4 -> mean, log_std clamp [-20, 2], tanh squashing) that `tvc_rollout`
  evaluates on tensor cores inside the env loop: `weights()` hands its parameters to the rollout kernel without a copy.
* Hyper-parameters come from the reference YAML's `algorithms.sac` block (config/config.yaml:35-54): `lr_actor`, `lr_critic`
  (fallback `learning_rate`), `batch_size`, `tau`, `gamma`, `buffer_size`, `learning_starts`, `grad_clip_norm`, `ent_coef`
  ("auto" = learned temperature with target entropy -|A|, initial value 0.2 = the constant `_update_sac` hard-codes).
* The networks and optimizers are plain PyTorch (cuBLAS GEMMs): the learner is adjacent to the hot path, not part of it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class SACConfig:
    lr_actor: float = 5e-5
    lr_critic: float = 1.5e-4
    lr_alpha: float = 1.5e-4
    batch_size: int = 256
    tau: float = 0.005
    gamma: float = 0.99
    buffer_size: int = 1_000_000
    learning_starts: int = 1000
    grad_clip_norm: float = 5.0
    ent_coef: str | float = "auto"
    init_alpha: float = 0.2          # agent/multi_algorithm_agent.py:996 (the constant entropy weight of `_update_sac`)
    reward_scale: float = 0.01       # rewards span [-1000, 200] (ref env :121): brought to O(1) for the critics
    tf32: bool = True                # the learner's GEMMs on the tensor cores in TF32 (fp32 accumulate); scoped to the update
    fused_adam: bool = True          # one fused multi-tensor Adam kernel per optimiser step (capturable)
    rule: str = "sac"                # "sac": squashed-Gaussian SAC (entropy term in the target, learned temperature);
                                     # "reference": `_update_sac` as the reference wrote it (see reference_update)

    @classmethod
    def reference_rule(cls, **over) -> "SACConfig":
        """The constants agent/multi_algorithm_agent.py hard-codes: Adam(3e-4) for policy and critics (:623-625), gamma 0.99
        (:970), entropy weight 0.2 (:996), tau 0.005 (:1002), no gradient clipping, rewards as the env returns them."""
        c = cls(lr_actor=3e-4, lr_critic=3e-4, lr_alpha=3e-4, tau=0.005, gamma=0.99, grad_clip_norm=0.0, ent_coef=0.2,
                reward_scale=1.0, rule="reference")
        for k, v in over.items():
            if not hasattr(c, k):
                raise AttributeError(f"SACConfig has no field {k!r}")
            setattr(c, k, v)
        return c

    @classmethod
    def from_yaml(cls, config: dict | None) -> "SACConfig":
        """`config` = the whole YAML dict (or its `algorithms` / `algorithms.sac` sub-dict)."""
        blk = config or {}
        if "algorithms" in blk:
            blk = blk["algorithms"] or {}
        if "sac" in blk:
            blk = blk["sac"] or {}
        c = cls()
        lr = blk.get("learning_rate")
        if lr is not None:
            c.lr_actor = c.lr_critic = c.lr_alpha = float(lr)
        for k in ("lr_actor", "lr_critic", "tau", "gamma", "grad_clip_norm", "reward_scale"):
            if blk.get(k) is not None:
                setattr(c, k, float(blk[k]))
        for k in ("batch_size", "buffer_size", "learning_starts"):
            if blk.get(k) is not None:
                setattr(c, k, int(blk[k]))
        if blk.get("ent_coef") is not None:
            c.ent_coef = blk["ent_coef"] if isinstance(blk["ent_coef"], str) else float(blk["ent_coef"])
        if blk.get("lr_critic") is not None:
            c.lr_alpha = float(blk["lr_critic"])
        return c


def _mlp(i: int, o: int, hidden: int = 256) -> nn.Sequential:
    return nn.Sequential(nn.Linear(i, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, o))


class Actor(nn.Module):
    """10-256-256-4 -> (mean, log_std): the network `tvc_rollout` evaluates in-kernel (include/tvc_b200.h tvc_actor_weights)."""

    def __init__(self, obs_dim: int = 10, action_dim: int = 2, hidden: int = 256):
        super().__init__()
        self.action_dim = action_dim
        self.net = _mlp(obs_dim, 2 * action_dim, hidden)   # (tvc_rollout is built for hidden = 256)

    def mean_log_std(self, obs):
        """(mean, log_std clamp [-20, 2]): what the reference's policy networks return (multi_algorithm_agent.py:223-227)."""
        out = self.net(obs)
        return out[:, :self.action_dim], out[:, self.action_dim:].clamp(-20.0, 2.0)

    def forward(self, obs, deterministic: bool = False, eps: torch.Tensor | None = None):
        out = self.net(obs)
        mean, log_std = out[:, :self.action_dim], out[:, self.action_dim:].clamp(-20.0, 2.0)
        if deterministic:
            return torch.tanh(mean), None
        std = log_std.exp()
        if eps is None:
            eps = torch.randn_like(mean)
        u = mean + std * eps
        a = torch.tanh(u)
        logp = (-0.5 * eps * eps - log_std - 0.5 * math.log(2.0 * math.pi)).sum(-1) - torch.log(1.0 - a * a + 1e-6).sum(-1)
        return a, logp


def reference_update(actor: Actor, q1, q2, q1t, q2t, opt_actor, opt_critic, batch: dict, gamma: float = 0.99,
                     alpha: float = 0.2, tau: float = 0.005, losses: dict | None = None):
    """One update by the reference's own rule, agent/multi_algorithm_agent.py:950-1016 `_update_sac`, on a batch:

    * target (:962-970): next action = a SAMPLE of the unsquashed Gaussian N(mean, exp(log_std)) of the policy, target
      `r + gamma (1 - done) min(Q1t, Q2t)(s', a')` -- no entropy term, no tanh;
    * critics (:972-984): MSE of each critic against that target (the reference steps two Adam optimisers; one Adam over both
      parameter sets is the same arithmetic, Adam being per-parameter);
    * policy (:988-1000): reparameterised sample a = mean + std eps, loss `-(min(Q1, Q2)(s, a) - alpha log N(a)).mean()` with the
      constant alpha = 0.2;
    * targets (:1002-1008): Polyak tau = 0.005.
    The random draws are made in the reference's order with the same torch calls (Normal.sample, then Normal.rsample), so that
    from one generator state both produce the same numbers (tests/golden/make_sac_update_golden.py runs the reference's
    function itself on these networks; tests/test_host.py compares)."""
    from torch.distributions import Normal
    s, a, r, sn, d = batch["obs"], batch["actions"], batch["reward"], batch["next_obs"], batch["done"]
    with torch.no_grad():
        mn, lsn = actor.mean_log_std(sn)
        # Normal(mean, std).sample() is torch.normal(mean, std) = eps * std + mean with eps from normal_(0, 1); written out,
        # because torch.normal(tensor, tensor) checks std >= 0 on the host (a sync: illegal inside a CUDA-graph capture)
        an = torch.randn_like(mn) * lsn.exp() + mn
        san = torch.cat([sn, an], -1)
        y = r + gamma * (1.0 - d) * torch.min(q1t(san), q2t(san)).squeeze(-1)
    sa = torch.cat([s, a], -1)
    lq1, lq2 = F.mse_loss(q1(sa).squeeze(-1), y), F.mse_loss(q2(sa).squeeze(-1), y)
    opt_critic.zero_grad(set_to_none=False)
    (lq1 + lq2).backward()
    opt_critic.step()
    m, ls = actor.mean_log_std(s)
    dist = Normal(m, ls.exp(), validate_args=False)
    ap = dist.rsample()
    sap = torch.cat([s, ap], -1)
    la = -(torch.min(q1(sap), q2(sap)) - alpha * dist.log_prob(ap).sum(-1, keepdim=True)).mean()
    opt_actor.zero_grad(set_to_none=False)
    la.backward()
    opt_actor.step()
    with torch.no_grad():
        src = list(q1.parameters()) + list(q2.parameters())
        dst = list(q1t.parameters()) + list(q2t.parameters())
        torch._foreach_mul_(dst, 1.0 - tau)
        torch._foreach_add_(dst, src, alpha=tau)
    if losses is not None:
        losses["q"].copy_((lq1 + lq2).detach())
        losses["actor"].copy_(la.detach())
    return lq1.detach(), lq2.detach(), la.detach()


class SACLearner:
    """`update()` = `updates_per_replay` SAC updates on fresh uniform samples of `replay` (one CUDA-graph replay)."""

    def __init__(self, replay, config: SACConfig | dict | None = None, updates_per_replay: int = 1, use_cuda_graph: bool = True,
                 seed: int = 0):
        self.cfg = config if isinstance(config, SACConfig) else SACConfig.from_yaml(config)
        self.replay, self.dev, self.U = replay, replay.device, int(updates_per_replay)
        replay.reward_scale = self.cfg.reward_scale
        torch.manual_seed(seed)
        dev = self.dev
        self.actor = Actor().to(dev)
        self.q1, self.q2, self.q1t, self.q2t = _mlp(12, 1).to(dev), _mlp(12, 1).to(dev), _mlp(12, 1).to(dev), _mlp(12, 1).to(dev)
        self.q1t.load_state_dict(self.q1.state_dict()), self.q2t.load_state_dict(self.q2.state_dict())
        for p in list(self.q1t.parameters()) + list(self.q2t.parameters()):
            p.requires_grad_(False)
        if self.cfg.rule not in ("sac", "reference"):
            raise ValueError(f"SACConfig.rule must be 'sac' or 'reference', got {self.cfg.rule!r}")
        self.auto_alpha = isinstance(self.cfg.ent_coef, str) and self.cfg.rule == "sac"
        a0 = self.cfg.init_alpha if isinstance(self.cfg.ent_coef, str) else float(self.cfg.ent_coef)
        self._alpha_const = float(a0)     # the reference rule's constant entropy weight (a Python float: graph-safe)
        self.log_alpha = torch.full((), math.log(a0), device=dev, requires_grad=self.auto_alpha)
        self.target_entropy = -float(self.actor.action_dim)
        cap = dict(capturable=True, fused=True) if self.cfg.fused_adam else dict(capturable=True)
        self.opt_actor = torch.optim.Adam(self.actor.parameters(), lr=self.cfg.lr_actor, **cap)
        self.opt_critic = torch.optim.Adam(list(self.q1.parameters()) + list(self.q2.parameters()), lr=self.cfg.lr_critic, **cap)
        self.opt_alpha = torch.optim.Adam([self.log_alpha], lr=self.cfg.lr_alpha, **cap) if self.auto_alpha else None
        self.batch = replay.new_batch(self.cfg.batch_size)
        self.losses = dict(q=torch.zeros((), device=dev), actor=torch.zeros((), device=dev), alpha=torch.zeros((), device=dev))
        self.updates = 0
        self._graph = None
        self._want_graph = use_cuda_graph

    # ------------------------------------------------------------------ acting side
    def weights(self) -> dict:
        """The actor's parameters in the layout of tvc_actor_weights (views, no copy): for `BatchedEngine.rollout`."""
        lin = [m for m in self.actor.net if isinstance(m, nn.Linear)]
        return dict(w1=lin[0].weight.detach(), b1=lin[0].bias.detach(), w2=lin[1].weight.detach(), b2=lin[1].bias.detach(),
                    w3=lin[2].weight.detach(), b3=lin[2].bias.detach())

    def policy(self, obs: torch.Tensor) -> torch.Tensor:
        """Deterministic policy (evaluation): tanh(mean)."""
        with torch.no_grad():
            return self.actor(obs, deterministic=True)[0]

    # ------------------------------------------------------------------ one update on self.batch
    def _clip(self, params):
        if self.cfg.grad_clip_norm and self.cfg.grad_clip_norm > 0:
            torch.nn.utils.clip_grad_norm_(params, self.cfg.grad_clip_norm, foreach=True)

    def _one_update(self):
        if self.cfg.rule == "reference":
            reference_update(self.actor, self.q1, self.q2, self.q1t, self.q2t, self.opt_actor, self.opt_critic, self.batch,
                             gamma=self.cfg.gamma, alpha=self._alpha_const, tau=self.cfg.tau, losses=self.losses)
            return
        b, g, tau = self.batch, self.cfg.gamma, self.cfg.tau
        s, a, r, sn, d = b["obs"], b["actions"], b["reward"], b["next_obs"], b["done"]
        alpha = self.log_alpha.exp().detach()
        with torch.no_grad():
            an, lpn = self.actor(sn)
            san = torch.cat([sn, an], 1)
            qn = torch.min(self.q1t(san), self.q2t(san)).squeeze(-1) - alpha * lpn
            y = r + g * (1.0 - d) * qn                                  # ref :960-970
        sa = torch.cat([s, a], 1)
        lq = F.mse_loss(self.q1(sa).squeeze(-1), y) + F.mse_loss(self.q2(sa).squeeze(-1), y)   # ref :972-976
        self.opt_critic.zero_grad(set_to_none=False)
        lq.backward()
        self._clip(list(self.q1.parameters()) + list(self.q2.parameters()))
        self.opt_critic.step()
        ap, lp = self.actor(s)
        sap = torch.cat([s, ap], 1)
        la = (alpha * lp - torch.min(self.q1(sap), self.q2(sap)).squeeze(-1)).mean()           # ref :988-996
        self.opt_actor.zero_grad(set_to_none=False)
        la.backward()
        self._clip(list(self.actor.parameters()))
        self.opt_actor.step()
        if self.auto_alpha:
            lal = -(self.log_alpha * (lp.detach() + self.target_entropy).mean())
            self.opt_alpha.zero_grad(set_to_none=False)
            lal.backward()
            self.opt_alpha.step()
            self.losses["alpha"].copy_(lal.detach())
        with torch.no_grad():                                                                   # ref :1002-1008
            src = list(self.q1.parameters()) + list(self.q2.parameters())
            dst = list(self.q1t.parameters()) + list(self.q2t.parameters())
            torch._foreach_mul_(dst, 1.0 - tau)
            torch._foreach_add_(dst, src, alpha=tau)
        self.losses["q"].copy_(lq.detach())
        self.losses["actor"].copy_(la.detach())

    def _updates(self):
        # TF32 for the update's matmuls only (the flag is read when a GEMM is dispatched, i.e. at capture time for the graph);
        # measured at batch 4,096: 1.60 ms per update in fp32 with foreach Adam, 0.95 ms with TF32 + fused Adam
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.cfg.tf32)
        try:
            for u in range(self.U):
                self.replay.sample_into(self.batch, draw=u, device_ctl=True)
                self._one_update()
            self.replay.tick(self.U)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

    def _capture(self):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(3):          # warm-up: allocates the Adam state and the autograd buffers outside the capture
                self._updates()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.updates += 3 * self.U
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._updates()
        self._graph = g

    def update(self) -> int:
        """`updates_per_replay` SAC updates; returns the number of updates done so far.  No-op before `learning_starts`."""
        if self.replay.filled < max(self.cfg.learning_starts, 1):
            return self.updates
        with torch.cuda.device(self.dev):
            if self._want_graph and self._graph is None:
                self._capture()
            if self._graph is not None:
                self._graph.replay()
            else:
                self._updates()
        self.updates += self.U
        return self.updates


def train_sac(engine, iters: int, rollout_steps: int = 8, config: SACConfig | dict | None = None, buffer: int | None = None,
              seed: int = 0, use_cuda_graph: bool = True, updates_per_iter: int | None = None):
    """The end-to-end loop of BASELINE config 5 on one GPU: fused rollout (acting on tensor cores in the env loop, transitions
    stored into the ring by the kernel) -> graph-replayed SAC updates, everything resident in HBM (replaces scripts/train.py:
    546-603).  Returns (learner, replay, timing dict with device-timed env / learner milliseconds per iteration)."""
    import time

    from .replay import DeviceReplay
    cfg = config if isinstance(config, SACConfig) else SACConfig.from_yaml(config)
    dev = engine.device
    rp = DeviceReplay(engine.n, rollout_steps, capacity=buffer or cfg.buffer_size, device=dev, seed=seed, reward_scale=cfg.reward_scale)
    U = rollout_steps if updates_per_iter is None else int(updates_per_iter)
    learner = SACLearner(rp, cfg, updates_per_replay=U, use_cuda_graph=use_cuda_graph, seed=seed)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_env = t_learn = 0.0
    timed = 0
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    for it in range(iters):
        ev[0].record()
        rp.collect(engine, learner.weights())
        ev[1].record()
        learner.update()
        ev[2].record()
        torch.cuda.synchronize(dev)
        if it >= 2:       # the first iterations capture the graph / warm the allocator
            t_env += ev[0].elapsed_time(ev[1])
            t_learn += ev[1].elapsed_time(ev[2])
            timed += 1
    wall = time.perf_counter() - wall0
    timed = max(timed, 1)
    return learner, rp, dict(env_ms_per_iter=t_env / timed, learner_ms_per_iter=t_learn / timed,
                             learner_share=t_learn / max(t_env + t_learn, 1e-9), wall_s=wall,
                             env_steps=engine.n * rollout_steps * iters,
                             env_steps_per_sec_e2e=engine.n * rollout_steps * timed / max((t_env + t_learn) * 1e-3, 1e-9),
                             updates=learner.updates, updates_per_iter=U, batch_size=cfg.batch_size)
