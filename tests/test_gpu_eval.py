"""GPU tests of the batched evaluation / telemetry helpers and of the
curriculum -> engine coupling (rank 2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_evaluate_zero_policy_matches_golden_episode(lib_built, golden_dir):
    """Deterministic reset + zero action: every episode is the zero_120 golden run (success at step 100)."""
    import os
    from tvc_ai_b200.evaluate import evaluate
    g = np.load(os.path.join(golden_dir, "zero_120.npz"))
    m = evaluate(lambda o: torch.zeros((o.shape[0], 2), device=o.device), episodes=8, config={}, contract="R")
    assert m["success_rate"] == 1.0 and m["length_mean"] == 100.0 and m["length_std"] == 0.0
    assert abs(m["reward_mean"] - g["reward"][:100].sum()) < 0.05 and m["reward_std"] < 1e-3
    assert m["safety_violation_rate"] == 0.0
    assert set(m) == {"reward_mean", "reward_std", "length_mean", "length_std", "success_rate", "safety_violation_rate",
                      "avg_safety_violations"}                       # scripts/train.py:691-699


def test_evaluate_scenarios_and_random_policy(lib_built):
    from tvc_ai_b200.evaluate import evaluate_scenarios
    cfg = {"evaluation": {"scenarios": {"nominal": {"episodes": 16}, "robustness": {"episodes": 16, "wind_force": 5.0,
                                                                                    "mass_variation": 0.5}}}}
    gen = torch.Generator(device="cuda").manual_seed(0)
    pol = lambda o: torch.rand((o.shape[0], 2), device=o.device, generator=gen) * 2 - 1  # noqa: E731
    out = evaluate_scenarios(pol, cfg)
    assert set(out) == {"nominal", "robustness"}
    for m in out.values():
        assert 1 <= m["length_mean"] <= 1000 and 0 <= m["success_rate"] <= 1 and np.isfinite(m["reward_mean"])
    assert out["nominal"]["length_mean"] < 200          # random gimbal commands tip the rocket over quickly


def test_record_trajectories_columns(lib_built, tmp_path):
    from tvc_ai_b200.evaluate import record_trajectories
    p = str(tmp_path / "traj.npz")
    out = record_trajectories(lambda o: torch.zeros((o.shape[0], 2), device=o.device), num_envs=4, steps=110, path=p)
    for k in ("step", "position", "orientation_euler", "linear_velocity", "angular_velocity", "action", "reward", "tilt_deg",
              "altitude", "fuel_remaining"):                       # scripts/evaluate.py:268-281
        assert k in out
    assert out["position"].shape == (110, 4, 3) and out["reward"].shape == (110, 4)
    assert np.all(out["episode_length"] == 100) and np.all(out["final_altitude"] > 0.49)
    assert np.allclose(out["fuel_remaining"][99], 0.9, atol=1e-6) and np.all(out["control_effort"] == 0)
    assert set(np.load(p).files) >= {"position", "episode_reward"}


def test_curriculum_conditions_reach_the_kernel(lib_built):
    """tvc_set_curriculum: stage conditions change the per-episode draws (initial tilt, wind, mass variation)."""
    from tvc_ai_b200 import RocketTVCVectorEnv
    from tvc_ai_b200.curriculum import CurriculumManager, stage6_conditions
    v = RocketTVCVectorEnv(2048, config={}, contract="X")
    cm = CurriculumManager({"enabled": True, "stages": {"stage_1": {"name": "hover_training", "episodes": 200, "environment": {
        "wind_force": 0.0, "mass_variation": 0.05, "initial_tilt_max": 0.05, "success_threshold": 0.7}}}})
    cond = cm.apply(v)
    assert cond["max_initial_tilt"] == 0.05 and not cond["wind_enabled"]
    v.reset(seed=1)
    st = v.engine.get_state()
    tilt = 2 * np.arcsin(np.clip(np.linalg.norm(st["quat"][:, :2], axis=1), 0, 1))
    assert tilt.max() <= 0.05 * np.sqrt(2) + 1e-6 and tilt.max() > 0.03
    assert np.all(st["wind"] == 0) and np.abs(st["mass_scale"] - 1).max() <= 0.05 + 1e-6
    v.set_curriculum(stage6_conditions())
    v.reset(seed=1)
    st = v.engine.get_state()
    tilt = 2 * np.arcsin(np.clip(np.linalg.norm(st["quat"][:, :2], axis=1), 0, 1))
    assert tilt.max() > 0.5 and st["wind"].std() > 2.0 and np.abs(st["mass_scale"] - 1).max() > 0.25
    with pytest.raises(RuntimeError):
        RocketTVCVectorEnv(4, config={}, contract="R").set_curriculum(cond)   # the reference env has no such coupling
    v.close()


@pytest.mark.gpu
def test_batched_curiosity_matches_single_env_facade(lib_built):
    """Row S14 / Q14 / Q19 at the VectorEnv boundary: with the same (never-trained) forward model the batched intrinsic
    reward equals the single-env facade's, including "skipped on the first step of an episode" across an autoreset."""
    import numpy as np
    import torch
    from tvc_ai_b200.env import EnhancedRocketTVCEnv
    from tvc_ai_b200.vector_env import RocketTVCVectorEnv
    torch.manual_seed(3)
    single = EnhancedRocketTVCEnv(config={}, enable_curiosity=True)
    vec = RocketTVCVectorEnv(4, config={}, contract="R", enable_curiosity=True, curiosity_module=single.curiosity_module)
    single.reset(seed=0)
    vec.reset(seed=0)
    rng = np.random.default_rng(9)
    episodes = 0
    for t in range(160):
        a = rng.uniform(-1, 1, 2).astype(np.float32)
        o1, r1, te1, tr1, info = single.step(a)
        o4, r4, te4, tr4, _ = vec.step(torch.from_numpy(np.tile(a, (4, 1))).cuda())
        assert abs(float(r4[0]) - float(r1)) <= 1e-5 * max(1.0, abs(float(r1))), (t, float(r4[0]), float(r1))
        assert torch.allclose(r4, r4[0].expand(4))
        assert bool(te4[0]) == bool(te1) and bool(tr4[0]) == bool(tr1)
        if te1 or tr1:
            episodes += 1
            single.reset(seed=0)          # the vector env reset itself in the same step
    assert episodes >= 1
    single.close(); vec.close()
