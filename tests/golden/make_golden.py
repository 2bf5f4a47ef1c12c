"""Freeze golden trajectories of the reference env -- run in the BUILD CONTAINER only.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

What runs: the reference's own, unmodified `EnhancedRocketTVCEnv`
(/root/reference/env/enhanced_rocket_tvc_env.py) imported from /root/reference, on top of
oracle/fake_pybullet.py (PyBullet itself is not installable here).  So every
number below is produced by the reference's Python for reward / phase / success / termination /
info, and by the oracle's physics layer for the Bullet calls.  /root/reference does not exist on
the GPU box, hence the committed fixtures.

Scenarios ("what the builder must create"):
  zero_120        zero action, 120 steps, raw gym.Env semantics (keeps stepping after success@100)
  random_raw      1000 steps of PCG64(42) uniform actions, no reset (continue past termination)
  random_autoreset same actions, env.reset() whenever terminated or truncated (VectorEnv semantics)
  two_episodes    zero action to success, reset, zero action again (Q10: success on step 1)
  burnout_1100    zero action for 1100 steps, no reset (S4 fuel thresholds 200 / 900 / 1000)
  crash_leak      hard-over action until the crash, then 40 more steps (Q13 variance penalty, Q12)
  curiosity_60    the TRAINING env's settings (scripts/train.py:313-321: enable_curiosity=True) under torch.manual_seed(42):
                  60 steps of the PCG64(42) actions with reset-on-done; records the intrinsic reward of every step
                  (row S14, quirks Q14 / Q19) and the never-trained forward model's weights, so that the facade's
                  CuriosityModule can be pinned to the reference's construction order and arithmetic
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fake_pybullet as fp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PHASES = ("boost", "coast", "landing", "touchdown", "hover", "complete", "failed")
COMP = ("mission_completion", "safety_compliance", "fuel_efficiency", "stability_bonus",
        "control_smoothness", "altitude_maintenance", "crash_penalty", "excessive_tilt", "control_saturation")


def actions_random(n=1000):
    return np.random.Generator(np.random.PCG64(42)).uniform(-1, 1, (n, 2)).astype(np.float32)


def run(mod, actions, autoreset=False, reset_at=(), curiosity=False):
    if curiosity:
        import torch
        torch.manual_seed(42)
    env = mod.EnhancedRocketTVCEnv(config={}, max_episode_steps=1000, enable_hierarchical=False,
                                   enable_curiosity=curiosity, enable_physics_informed=False)
    obs0, info0 = env.reset(seed=42)
    T = len(actions)
    rec = dict(obs=np.zeros((T, 10), np.float32), reward=np.zeros(T), terminated=np.zeros(T, bool),
               truncated=np.zeros(T, bool), comp=np.zeros((T, len(COMP))), state=np.zeros((T, 13)),
               altitude=np.zeros(T), tilt_deg=np.zeros(T), omega_mag=np.zeros(T), fuel=np.zeros(T),
               phase=np.zeros(T, np.int32), success=np.zeros(T, bool), step=np.zeros(T, np.int32),
               criteria_met=np.zeros(T, bool), was_reset=np.zeros(T, bool), next_obs=np.zeros((T, 10), np.float32),
               curiosity=np.zeros(T), has_curiosity=np.zeros(T, bool))
    W = fp.world()
    for t in range(T):
        obs, r, term, trunc, info = env.step(actions[t].copy())
        body = [b for b in W.bodies.values() if not isinstance(b, str)][0]
        rec["obs"][t] = obs
        rec["reward"][t] = float(r)
        rec["terminated"][t] = term
        rec["truncated"][t] = trunc
        rc = info["reward_components"]
        rec["comp"][t] = [float(rc.get(k, 0.0)) for k in COMP]
        rec["has_curiosity"][t] = "curiosity" in rc
        rec["curiosity"][t] = float(rc.get("curiosity", 0.0))
        rec["state"][t] = list(body.pos) + list(body.quat) + list(body.vel) + list(body.omega)
        rec["altitude"][t] = info["altitude"]
        rec["tilt_deg"][t] = info["tilt_angle_deg"]
        rec["omega_mag"][t] = info["angular_velocity_mag"]
        rec["fuel"][t] = info["fuel_remaining"]
        rec["phase"][t] = PHASES.index(info["mission_phase"])
        rec["success"][t] = info["mission_successful"]
        rec["step"][t] = info["step"]
        rec["criteria_met"][t] = bool(info["success_criteria_met"])
        rec["next_obs"][t] = obs
        if (autoreset and (term or trunc)) or t in reset_at:
            o2, _ = env.reset()
            rec["was_reset"][t] = True
            rec["next_obs"][t] = o2
    if curiosity:   # the forward model as the reference constructed it (torch default init, inverse model first)
        fm = env.curiosity_module.forward_model
        for k, i in (("1", 0), ("2", 2), ("3", 4)):
            rec["fm_w" + k] = fm[i].weight.detach().numpy().copy()
            rec["fm_b" + k] = fm[i].bias.detach().numpy().copy()
    env.close()
    rec["actions"] = np.asarray(actions, np.float32)
    rec["obs0"] = obs0
    return rec


def main():
    mod = fp.load_reference_env()
    rnd = actions_random(1000)
    np.save(os.path.join(HERE, "actions_pcg64_42.npy"), rnd)
    zero = lambda n: np.zeros((n, 2), np.float32)  # noqa: E731
    hard = np.tile(np.array([[1.0, -1.0]], np.float32), (120, 1))
    scen = {
        "zero_120": run(mod, zero(120)),
        "random_raw": run(mod, rnd),
        "random_autoreset": run(mod, rnd, autoreset=True),
        "two_episodes": run(mod, zero(140), reset_at=(99,)),
        "burnout_1100": run(mod, zero(1100)),
        "crash_leak": run(mod, hard),
        "curiosity_60": run(mod, rnd[:60], autoreset=True, curiosity=True),
    }
    for name, rec in scen.items():
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **rec)
        print(f"{name:18s} T={len(rec['reward']):5d} terminated@{np.flatnonzero(rec['terminated'])[:3]} "
              f"sum_reward={rec['reward'].sum():.6f}")


if __name__ == "__main__":
    main()
