"""Row S14 (quirks Q14, Q19): the never-trained curiosity forward model, pinned to the REFERENCE's own numbers.
tests/golden/curiosity_60.npz was produced by the unmodified reference class with enable_curiosity=True under
torch.manual_seed(42) (tests/golden/make_golden.py): per-step intrinsic reward and the forward model's weights
(/root/reference/env/enhanced_rocket_tvc_env.py:226-269, :494-502)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "curiosity_60.npz"))


def _load_weights(mod, g):
    fm = mod.forward_model
    with torch.no_grad():
        for k, i in (("1", 0), ("2", 2), ("3", 4)):
            fm[i].weight.copy_(torch.from_numpy(g["fm_w" + k]))
            fm[i].bias.copy_(torch.from_numpy(g["fm_b" + k]))


def test_construction_order_gives_the_reference_weights(golden_dir):
    """Same torch seed -> same random forward model: the facade builds the inverse model first, like ref:233-249."""
    from tvc_ai_b200.env import CuriosityModule
    g = _golden(golden_dir)
    torch.manual_seed(42)
    m = CuriosityModule(obs_dim=8, action_dim=2, device="cpu")
    for k, i in (("1", 0), ("2", 2), ("3", 4)):
        assert np.array_equal(m.forward_model[i].weight.detach().numpy(), g["fm_w" + k]), k
        assert np.array_equal(m.forward_model[i].bias.detach().numpy(), g["fm_b" + k]), k


def test_intrinsic_reward_reproduces_the_reference_per_step(golden_dir):
    """0.01 * MSE(f([s8, a]), s8') on the golden (previous obs, clipped action, obs) triplets; skipped on the first step of
    every episode (state_history is cleared by reset, ref:399-401, :496)."""
    from tvc_ai_b200.env import CuriosityModule
    g = _golden(golden_dir)
    m = CuriosityModule(obs_dim=8, action_dim=2, device="cpu")
    _load_weights(m, g)
    T = len(g["reward"])
    prev = None
    n = 0
    for t in range(T):
        assert bool(g["has_curiosity"][t]) == (prev is not None), t
        if prev is not None:
            r = m.compute_intrinsic_reward(prev, np.clip(g["actions"][t], -1, 1), g["obs"][t][:8])
            assert abs(r - g["curiosity"][t]) <= 1e-9 + 1e-6 * abs(g["curiosity"][t]), (t, r, g["curiosity"][t])
            # Q14: added after the clip, on top of the extrinsic total
            assert abs(g["reward"][t] - (np.clip(g["comp"][t].sum() + 0.0, -1000, 200) + g["curiosity"][t])) < 0.06
            n += 1
        prev = None if g["was_reset"][t] else g["obs"][t][:8].copy()
    assert n >= 55 and g["curiosity"].max() > 0


@pytest.mark.gpu
def test_facade_and_vector_env_match_the_reference_curiosity(lib_built, golden_dir, parity_record):
    """The single-env facade and the batched VectorEnv with the reference's forward-model weights reproduce the reference's
    per-step reward_components['curiosity'] and total reward on the golden run (device obs differ from fp64 by ~1e-6)."""
    from tvc_ai_b200.env import EnhancedRocketTVCEnv, CuriosityModule
    from tvc_ai_b200.vector_env import RocketTVCVectorEnv
    g = _golden(golden_dir)
    T = len(g["reward"])
    env = EnhancedRocketTVCEnv(config={}, enable_hierarchical=False, enable_curiosity=True, enable_physics_informed=False)
    _load_weights(env.curiosity_module, g)
    n = 3
    cm = CuriosityModule(obs_dim=8, action_dim=2, device="cuda")
    _load_weights(cm, g)
    venv = RocketTVCVectorEnv(n, config={}, contract="R", enable_curiosity=True, curiosity_module=cm, final_info=False,
                              curiosity_impl="torch")
    tenv = RocketTVCVectorEnv(n, config={}, contract="R", enable_curiosity=True, curiosity_module=cm, final_info=False,
                              curiosity_impl="tcgen05")      # the batch default: forward model on the tensor cores (bf16)
    env.reset(seed=42)
    venv.reset(seed=42)
    tenv.reset(seed=42)
    worst_c, worst_r, worst_v, worst_tc, worst_tr = 0.0, 0.0, 0.0, 0.0, 0.0
    for t in range(T):
        a = g["actions"][t]
        obs, r, term, trunc, info = env.step(a)
        assert ("curiosity" in info["reward_components"]) == bool(g["has_curiosity"][t]), t
        c = info["reward_components"].get("curiosity", 0.0)
        worst_c = max(worst_c, abs(c - g["curiosity"][t]) / max(1e-3, abs(g["curiosity"][t])))
        worst_r = max(worst_r, abs(float(r) - g["reward"][t]) / max(1.0, abs(g["reward"][t])))
        assert term == bool(g["terminated"][t]) and trunc == bool(g["truncated"][t])
        _, rv, tv, _, _ = venv.step(torch.from_numpy(np.tile(a, (n, 1))).cuda())
        worst_v = max(worst_v, float((rv - float(r)).abs().max()) / max(1.0, abs(float(r))))
        assert bool(tv[0]) == term
        _, rt, _, _, _ = tenv.step(torch.from_numpy(np.tile(a, (n, 1))).cuda())
        worst_tc = max(worst_tc, abs(float(tenv.intrinsic[0]) - g["curiosity"][t]) / max(1e-3, abs(g["curiosity"][t])))
        worst_tr = max(worst_tr, abs(float(rt[0]) - g["reward"][t]) / max(1.0, abs(g["reward"][t])))
        if g["was_reset"][t]:
            env.reset()          # the VectorEnv reset itself in the same step
    parity_record["curiosity_vs_reference"] = dict(steps=T, intrinsic_rel_max=worst_c, reward_rel_max=worst_r,
                                                   batched_vs_facade_rel_max=worst_v, tcgen05_intrinsic_rel_max=worst_tc,
                                                   tcgen05_reward_rel_max=worst_tr)
    assert worst_c <= 2e-3 and worst_r <= 2e-4 and worst_v <= 1e-5, parity_record["curiosity_vs_reference"]
    assert worst_tc <= 3e-2 and worst_tr <= 2e-4, parity_record["curiosity_vs_reference"]     # bf16 bar on the term, the same reward bar
    env.close(); venv.close(); tenv.close()
