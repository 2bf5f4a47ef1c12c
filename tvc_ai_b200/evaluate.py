"""Batched evaluation and telemetry on the CUDA engine.

`evaluate()` turns the reference's `StateOfTheArtTrainer.evaluate` (scripts/train.py:645-700: `episodes`
sequential episodes on the eval env, deterministic policy, safety-violation counting by
`_check_safety_violation`, :620-641) into one batch: every episode is one env of a `BatchedEngine`, all
stepped together until each has ended once.  Same metric names as train.py:691-699.

`evaluate_scenarios()` walks `config['evaluation']['scenarios']` (config/config.yaml:360-379).

`record_trajectories()` writes the per-step columns the legacy tool produced (scripts/evaluate.py:268-281:
step, position, orientation_euler, linear_velocity, angular_velocity, action, reward, tilt_deg, altitude,
fuel_remaining) as columnar arrays [T, N, ...] plus the per-episode summary of scripts/evaluate.py:284-306.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch

from . import _abi as A
from .engine import BatchedEngine
from .env import engine_config_from_yaml

Policy = Callable[[torch.Tensor], torch.Tensor]   # obs [N,10] cuda float32 -> actions [N,2]


def _safety_violation(info, safety: Optional[dict]) -> torch.Tensor:
    """scripts/train.py:620-641 on the terminal-info tensors of step_ex."""
    s = safety or {}
    max_tilt_deg = float(np.degrees(s.get("max_tilt", 0.52)))
    return ((info["tilt_deg"] > max_tilt_deg) | (info["omega_mag"] > float(s.get("max_angular_velocity", 5.0)))
            | (info["altitude"] < float(s.get("min_altitude", 0.1))) | (info["altitude"] > float(s.get("max_altitude", 20.0))))


def evaluate(policy: Policy, episodes: int = 20, config: Optional[dict] = None, contract: str | int = "R",
             max_episode_steps: int = 1000, device: Optional[int] = None, conditions: Optional[dict] = None,
             seed: Optional[int] = None, **engine_over) -> Dict[str, float]:
    if isinstance(contract, str):
        contract = {"R": A.CONTRACT_R, "X": A.CONTRACT_X}[contract.upper()]
    cfg = engine_config_from_yaml(config, contract, max_episode_steps, autoreset=0, **engine_over)
    if seed is not None:
        cfg.seed = int(seed)
    eng = BatchedEngine(int(episodes), cfg, device=device)
    try:
        if conditions:
            eng.set_curriculum(conditions)
        obs = eng.reset()
        n, dev = eng.n, eng.device
        ret = torch.zeros(n, dtype=torch.float64, device=dev)
        length = torch.zeros(n, dtype=torch.int64, device=dev)
        viol = torch.zeros(n, dtype=torch.int64, device=dev)
        success = torch.zeros(n, dtype=torch.bool, device=dev)
        active = torch.ones(n, dtype=torch.bool, device=dev)
        safety = (config or {}).get("safety", {}).get("constrained_rl", {}).get("constraints") or (config or {}).get("safety")
        for _ in range(max_episode_steps + 1):
            with torch.no_grad():
                act = policy(obs).to(dtype=torch.float32).contiguous()
            obs, rew, term, trunc, info = eng.step_ex(act)
            ret += torch.where(active, rew.double(), torch.zeros_like(ret))
            length += active
            viol += active & _safety_violation(info, safety)
            done = (term.bool() | trunc.bool()) & active
            success |= done & info["success"].bool()
            active &= ~done
            if not bool(active.any()):
                break
        r, ln, v = ret.cpu().numpy(), length.cpu().numpy(), viol.cpu().numpy()
        return {"reward_mean": float(np.mean(r)), "reward_std": float(np.std(r)), "length_mean": float(np.mean(ln)),
                "length_std": float(np.std(ln)), "success_rate": float(success.float().mean().item()),
                "safety_violation_rate": float(np.mean(v > 0)), "avg_safety_violations": float(np.mean(v))}
    finally:
        eng.close()


def evaluate_scenarios(policy: Policy, config: dict, device: Optional[int] = None) -> Dict[str, Dict[str, float]]:
    """config/config.yaml:360-379.  `robustness` carries wind_force / mass_variation, which only Contract X can apply;
    `landing` / `hovering` name target altitudes the shipped env never reads, so they run the nominal env."""
    out = {}
    for name, sc in (config.get("evaluation", {}).get("scenarios", {}) or {}).items():
        episodes = int(sc.get("episodes", 20))
        if "wind_force" in sc or "mass_variation" in sc:
            cond = {"domain_randomization": True, "mass_variation": sc.get("mass_variation", 0.0),
                    "wind_enabled": sc.get("wind_force", 0.0) > 0, "wind_force": sc.get("wind_force", 0.0)}
            out[name] = evaluate(policy, episodes, config, contract="X", device=device, conditions=cond)
        else:
            out[name] = evaluate(policy, episodes, config, contract="R", device=device)
    return out


def _euler_from_quat(q: np.ndarray) -> np.ndarray:
    """pybullet getEulerFromQuaternion on [...,4] arrays (display only)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    sarg = np.clip(-2 * (x * z - w * y), -1, 1)
    roll = np.arctan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)
    yaw = np.arctan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)
    return np.stack([roll, np.arcsin(sarg), yaw], axis=-1)


def record_trajectories(policy: Policy, num_envs: int = 16, steps: int = 200, config: Optional[dict] = None,
                        contract: str | int = "R", device: Optional[int] = None, path: Optional[str] = None,
                        **engine_over) -> Dict[str, np.ndarray]:
    if isinstance(contract, str):
        contract = {"R": A.CONTRACT_R, "X": A.CONTRACT_X}[contract.upper()]
    cfg = engine_config_from_yaml(config, contract, 1000, autoreset=0, **engine_over)
    eng = BatchedEngine(int(num_envs), cfg, device=device)
    try:
        obs = eng.reset()
        cols = {k: [] for k in ("position", "orientation_euler", "linear_velocity", "angular_velocity", "action", "reward",
                                "tilt_deg", "altitude", "fuel_remaining", "terminated", "truncated")}
        for _ in range(steps):
            with torch.no_grad():
                act = policy(obs).to(dtype=torch.float32).contiguous()
            obs, rew, term, trunc, info = eng.step_ex(act)
            st = eng.get_state()
            cols["position"].append(st["pos"].copy())
            cols["orientation_euler"].append(_euler_from_quat(st["quat"].astype(np.float64)).astype(np.float32))
            cols["linear_velocity"].append(st["vel"].copy())
            cols["angular_velocity"].append(st["omega"].copy())
            cols["action"].append(act.cpu().numpy())
            cols["reward"].append(rew.cpu().numpy().copy())
            cols["tilt_deg"].append(info["tilt_deg"].cpu().numpy().copy())
            cols["altitude"].append(info["altitude"].cpu().numpy().copy())
            cols["fuel_remaining"].append(info["fuel"].cpu().numpy().copy())
            cols["terminated"].append(term.cpu().numpy().astype(bool))
            cols["truncated"].append(trunc.cpu().numpy().astype(bool))
        out = {k: np.stack(v) for k, v in cols.items()}
        out["step"] = np.arange(1, steps + 1, dtype=np.int32)
        # per-episode summary (scripts/evaluate.py:284-306) over the first episode of each env
        done = out["terminated"] | out["truncated"]
        first = np.where(done.any(0), done.argmax(0), steps - 1)
        idx = np.arange(out["reward"].shape[1])
        mask = np.arange(steps)[:, None] <= first[None, :]
        length = first + 1
        out["episode_length"] = length.astype(np.int32)
        out["episode_reward"] = (out["reward"] * mask).sum(0)
        out["final_tilt_deg"] = out["tilt_deg"][first, idx]
        out["final_altitude"] = out["altitude"][first, idx]
        out["max_tilt_deg"] = np.where(mask, out["tilt_deg"], -np.inf).max(0)
        out["control_effort"] = (np.linalg.norm(out["action"], axis=-1) * mask).sum(0) / length
        out["fuel_used"] = 1.0 - out["fuel_remaining"][first, idx]
        if path:
            np.savez_compressed(path, **out)
        return out
    finally:
        eng.close()
