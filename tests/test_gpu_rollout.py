"""GPU tests of the fused actor rollout (tvc_rollout, BASELINE config 4).

The actor GEMMs run on tcgen05 tensor cores with bf16 operands and fp32 accumulation, so the parity
reference is a PyTorch fp32 `nn.Sequential` (tolerance 3e-2 on tanh-squashed actions, the bf16 bar)
AND a bit-faithful emulation of the kernel's rounding points (bf16 inputs / weights / first hidden
layer, fp32 accumulate; tolerance 2e-4), which pins the UMMA descriptors and TMEM read-out.
Physics inside the rollout is the same device code as tvc_step: rewards must agree to 1e-4.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _actor(seed=0):
    torch.manual_seed(seed)
    nn = torch.nn
    net = nn.Sequential(nn.Linear(10, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 4)).cuda()
    w = dict(w1=net[0].weight.detach(), b1=net[0].bias.detach(), w2=net[2].weight.detach(), b2=net[2].bias.detach(),
             w3=net[4].weight.detach(), b3=net[4].bias.detach())
    return net, w


def _emulate_bf16(w, obs):
    bf = lambda x: x.to(torch.bfloat16).to(torch.float32)  # noqa: E731
    h1 = torch.relu(bf(obs) @ bf(w["w1"]).T + w["b1"])
    h2 = torch.relu(bf(h1) @ bf(w["w2"]).T + w["b2"])
    return h2 @ w["w3"].T + w["b3"]


def _engines(n, **over):
    from tvc_ai_b200 import _abi as A
    from tvc_ai_b200.engine import BatchedEngine
    cfg = dict(autoreset=1, init_tilt_max=0.1)
    cfg.update(over)
    a = BatchedEngine(n, A.default_config(A.CONTRACT_X, **cfg), device=0)
    b = BatchedEngine(n, A.default_config(A.CONTRACT_X, **cfg), device=0)
    a.reset(), b.reset()
    return a, b


def test_rollout_deterministic_matches_torch_and_step_kernel(lib_built):
    net, w = _actor()
    n, T = 1000, 12          # 1000: not a multiple of the CTA size
    a, b = _engines(n)
    out = a.rollout(w, T, deterministic=True, record=True)
    torch.cuda.synchronize()
    worst_fp32, worst_emu, worst_rew = 0.0, 0.0, 0.0
    with torch.no_grad():
        for t in range(T):
            obs = b.obs.clone()
            ref = torch.tanh(net(obs)[:, :2])
            emu = torch.tanh(_emulate_bf16(w, obs)[:, :2])
            act = out["actions_all"][t]
            worst_fp32 = max(worst_fp32, float((act - ref).abs().max()))
            worst_emu = max(worst_emu, float((act - emu).abs().max()))
            _, rew, _, _ = b.step(act.contiguous())
            worst_rew = max(worst_rew, float(((rew - out["reward_all"][t]).abs() / rew.abs().clamp(min=1.0)).max()))
    print(f"\n[rollout] {n} envs x {T} steps: max |action - torch fp32| = {worst_fp32:.2e}, "
          f"max |action - bf16 emulation| = {worst_emu:.2e}, max reward rel diff vs tvc_step = {worst_rew:.2e}")
    assert worst_emu < 2e-4
    assert worst_fp32 < 3e-2
    assert worst_rew < 1e-4
    assert torch.allclose(a.obs, b.obs, atol=1e-5)
    np.testing.assert_allclose(out["reward_sum"].cpu().numpy(), out["reward_all"].sum(0).cpu().numpy(), rtol=1e-5, atol=1e-3)
    assert a.lifetime_steps == b.lifetime_steps == T
    sa, sb = a.stats(), b.stats()
    assert sa[14] == sb[14] == n * T
    np.testing.assert_allclose(sa[[0, 3, 4, 5, 6, 9]], sb[[0, 3, 4, 5, 6, 9]], atol=2)
    a.close(), b.close()


def test_rollout_noise_is_philox_stream_6(lib_built, oracle_mod):
    """mean = 0, log_std = 0 (zero head) -> action = tanh(eps): eps must be the Box-Muller normals of Philox
    stream 6 at counter (global env id, lifetime step), identical to what the oracle library draws."""
    O = oracle_mod
    _, w = _actor()
    w = dict(w)
    w["w3"] = torch.zeros_like(w["w3"])
    w["b3"] = torch.zeros_like(w["b3"])
    n, T = 256, 3
    a, b = _engines(n, env_id_base=5000)
    seed = a.config.seed
    out = a.rollout(w, T, deterministic=False, record=True)
    act = out["actions_all"].cpu().numpy()
    eps = np.arctanh(np.clip(act, -0.999999, 0.999999))
    for t in range(T):
        for i in (0, 1, 77, 255):
            gid = 5000 + i
            r = O.philox4x32_10((gid & 0xFFFFFFFF, ((gid >> 32) & 0xFFFF) | (6 << 16), t, 0), (seed & 0xFFFFFFFF, seed >> 32))
            u0 = ((r[0] >> 9) + 0.5) / 8388608.0
            u1 = ((r[1] >> 9) + 0.5) / 8388608.0
            rad = math.sqrt(-2.0 * math.log(u0))
            e0, e1 = rad * math.cos(2 * math.pi * u1), rad * math.sin(2 * math.pi * u1)
            if abs(e0) < 3 and abs(e1) < 3:
                assert abs(eps[t, i, 0] - e0) < 2e-3 and abs(eps[t, i, 1] - e1) < 2e-3, (t, i, eps[t, i], e0, e1)
    assert abs(eps.mean()) < 0.1 and 0.8 < eps.std() < 1.2
    out2 = b.rollout(w, T, deterministic=False, record=True)
    assert torch.equal(out2["actions_all"], out["actions_all"])     # same seed, same ids -> same noise
    a.close(), b.close()


def test_rollout_transition_record_feeds_a_replay_buffer(lib_built):
    """tvc_rollout writes (obs, action, reward, next_obs, terminated, truncated) for every step straight into caller
    tensors (on-device replay feed): they must equal what stepping the same actions through tvc_step returns."""
    net, w = _actor(1)
    n, T = 700, 10
    a, b = _engines(n)
    dev = a.device
    tr = dict(obs=torch.zeros((T, n, 10), device=dev), actions=torch.zeros((T, n, 2), device=dev),
              reward=torch.zeros((T, n), device=dev), next_obs=torch.zeros((T, n, 10), device=dev),
              terminated=torch.zeros((T, n), dtype=torch.uint8, device=dev), truncated=torch.zeros((T, n), dtype=torch.uint8, device=dev))
    a.rollout(w, T, deterministic=False, transitions=tr)
    torch.cuda.synchronize()
    n_done = 0
    for t in range(T):
        assert torch.allclose(tr["obs"][t], b.obs, atol=1e-5)
        obs, rew, term, trunc = b.step(tr["actions"][t].contiguous())
        done = (term | trunc).bool()
        nxt = torch.where(done[:, None], b.final_obs, obs)
        assert torch.equal(tr["terminated"][t], term) and torch.equal(tr["truncated"][t], trunc)
        assert torch.allclose(tr["next_obs"][t], nxt, atol=1e-5)
        assert torch.allclose(tr["reward"][t], rew, rtol=1e-5, atol=1e-4)
        n_done += int(done.sum())
    assert bool((tr["actions"].abs() <= 1).all())
    a.close(), b.close()
