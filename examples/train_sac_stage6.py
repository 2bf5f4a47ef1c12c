#!/usr/bin/env python
"""BASELINE config 5: curriculum "stage 6" (max noise, actuator delay, thrust-curve variation) with an end-to-end SAC
loop on the batched CUDA env -- everything stays on the GPU (no host round trip per transition).

What it mirrors: the reference's train loop (scripts/train.py:406-533: act -> env.step -> agent.update, periodic
evaluation), with the batch-of-1 host loop replaced by the package's components:
`tvc_rollout` evaluates the actor on tensor cores inside the env loop and stores the transitions straight into the on-device
replay ring (`tvc_ai_b200.replay.DeviceReplay`); `tvc_ai_b200.sac.SACLearner` runs the batched updates as one CUDA-graph
replay (`--rule sac`: squashed-Gaussian SAC with the YAML's hyper-parameters; `--rule reference`: the arithmetic of the
reference's own `_update_sac`, agent/multi_algorithm_agent.py:950-1016, pinned by tests/test_host.py).

    python examples/train_sac_stage6.py --envs 4096 --iters 200

Reports env-steps/s end to end and the learner's share of the device time.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tvc_ai_b200 import RocketTVCVectorEnv
from tvc_ai_b200.curriculum import stage6_conditions
from tvc_ai_b200.evaluate import evaluate
from tvc_ai_b200.sac import SACConfig, train_sac


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--buffer", type=int, default=1 << 20)
    ap.add_argument("--rollout-steps", type=int, default=8, help="env steps per fused-rollout launch (and SAC updates per iteration)")
    ap.add_argument("--rule", choices=("sac", "reference"), default="sac")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--eager-learner", action="store_true", help="run the SAC update eagerly instead of as a CUDA graph")
    args = ap.parse_args()

    # stage 6 = stage_5 conditions + sensor noise + actuator delay (3 control steps) + thrust-curve variation
    env = RocketTVCVectorEnv(args.envs, config={"globals": {"seed": args.seed}}, contract="X", device=0, final_info=False,
                             delay_steps=3, thrust_curve=1, propellant_fraction=0.2, cg_burn_shift=0.05)
    env.set_curriculum(stage6_conditions())
    env.reset(seed=args.seed, options={"return_torch": True})
    T = args.rollout_steps
    if args.rule == "reference":
        cfg = SACConfig.reference_rule(batch_size=args.batch, learning_starts=args.envs * T, buffer_size=args.buffer)
    else:
        cfg = SACConfig(batch_size=args.batch, learning_starts=args.envs * T, lr_actor=3e-4, lr_critic=3e-4, ent_coef=0.2,
                        buffer_size=args.buffer)
    learner, replay, timing = train_sac(env.engine, args.iters, rollout_steps=T, config=cfg, seed=args.seed,
                                        use_cuda_graph=not args.eager_learner)
    stats = env.episode_stats()
    env.close()
    evalm = evaluate(learner.policy, episodes=64, contract="X", conditions=stage6_conditions(), delay_steps=3, thrust_curve=1)
    print(json.dumps({
        "config": "stage 6: wind 3 N, mass +-30 %, initial tilt 0.7 rad, sensor noise 0.02, actuator delay 3 steps, thrust curve",
        "envs": args.envs, "iters": args.iters, "rollout_steps": T, "env_steps": timing["env_steps"],
        "env_steps_per_sec_end_to_end": timing["env_steps_per_sec_e2e"], "acting": "tvc_rollout (tcgen05 actor in-kernel)",
        "env_ms_per_iter": timing["env_ms_per_iter"], "learner_ms_per_iter": timing["learner_ms_per_iter"],
        "learner_share_of_device_time": timing["learner_share"],
        "learner": f"tvc_ai_b200.sac.SACLearner (rule={args.rule}), " + ("eager" if args.eager_learner else "CUDA graph replay"),
        "updates": timing["updates"], "replay_filled": replay.filled,
        "episodes": stats["episodes"], "train_success_rate": stats["successes"] / max(stats["episodes"], 1),
        "eval": evalm}))


if __name__ == "__main__":
    main()
