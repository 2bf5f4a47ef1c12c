"""Host-side logic on CPU: spaces, YAML mapping, curriculum index logic vs the reference's
CurriculumManager (fixture tests/golden/curriculum.json), env slabs, gloo all-reduce (world 2)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spaces_match_reference_declaration():
    from tvc_ai_b200 import spaces
    o, a = spaces.observation_space(), spaces.action_space()
    assert o.shape == (10,) and a.shape == (2,) and o.dtype == np.float32
    assert list(o.low[:7]) == [-1, -1, -1, -1, -10, -10, -10] and list(o.high[7:]) == [1, 1, 1]
    a.seed(0)
    s = a.sample()
    assert s.shape == (2,) and s.dtype == np.float32 and a.contains(s)
    assert spaces.batch_space(o, 8).shape == (8, 10)


def test_yaml_mapping_honours_the_keys_the_reference_reads(lib_built):
    from tvc_ai_b200 import _abi
    from tvc_ai_b200.env import engine_config_from_yaml
    # the reference looks for gradient_penalty at reward_function top level (ref:83-84); the shipped
    # YAML nests it under anti_hacking, so the defaults apply
    yaml_like = {"reward_function": {"anti_hacking": {"gradient_penalty": 0.7}}}
    c = engine_config_from_yaml(yaml_like, _abi.CONTRACT_R, 1000)
    assert abs(c.gradient_penalty - 0.1) < 1e-7
    c = engine_config_from_yaml({"reward_function": {"gradient_penalty": 0.3}}, _abi.CONTRACT_R, 500)
    assert abs(c.gradient_penalty - 0.3) < 1e-7 and c.max_episode_steps == 500
    x = engine_config_from_yaml({"env": {"domain_randomization": {"enabled": True, "parameters": {
        "mass": {"variation": 0.2}, "wind": {"max_force": 1.5}, "sensor_noise": {"std": 0.01}}}},
        "globals": {"seed": 7}}, _abi.CONTRACT_X, 1000)
    assert abs(x.mass_variation - 0.2) < 1e-7 and x.wind_std == 1.5 and x.seed == 7


def test_curriculum_matches_reference_manager():
    from tvc_ai_b200.curriculum import CurriculumManager
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "curriculum.json")))
    cm = CurriculumManager(g["curriculum_config"])
    assert [(s.name, s.duration_steps) for s in cm.stages] == [(s["name"], s["duration_steps"]) for s in g["stages"]]
    for s, ref in zip(cm.stages, g["stages"]):
        # the reference's update() writes `_curriculum_info` into the stage's own conditions dict
        # (curriculum_manager.py:279-289 mutates what get_environment_config() returned); ours copies
        ref_cond = {k: v for k, v in ref["conditions"].items() if k != "_curriculum_info"}
        assert s.conditions == ref_cond and s.success_criteria == ref["success_criteria"]
    for call, ref in zip(g["schedule"], g["trace"]):
        out = cm.update(call["step"], call["metrics"])
        assert cm.current_stage_idx == ref["stage_index"], call
        assert out.get("_curriculum_info", {}) == ref["info"], call
        assert {k: v for k, v in out.items() if k != "_curriculum_info"} == ref["conditions"]
    assert cm.is_curriculum_complete()


def test_slabs_partition_the_env_range():
    from tvc_ai_b200.dist import slab
    for total, world in ((2097152, 8), (4096, 3), (10, 4)):
        spans = [slab(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (b0, c0), (b1, _) in zip(spans, spans[1:]):
            assert b0 + c0 == b1


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch
from tvc_ai_b200 import dist as D
from tvc_ai_b200.curriculum import CurriculumManager
rank, world, _ = D.init_from_env("gloo")
base, count = D.slab(1000, rank, world)
v = np.zeros(16); v[0] = count; v[4] = count * 0.8; v[1] = 120.0 * count; v[14] = base
out = D.allreduce_stats(v)
d = D.stats_dict(out)
assert d["episodes"] == 1000 and abs(d["successes"] - 800) < 1e-9, d
cm = CurriculumManager({"enabled": True, "stages": {"s1": {"episodes": 1, "environment": {"success_threshold": 0.7}},
                                                    "s2": {"episodes": 1, "environment": {"mass_variation": 0.1}}}})
cm.update_from_stats(900, d)
print("RANK", rank, "stage", cm.current_stage_idx, flush=True)
assert cm.current_stage_idx == 1
torch.distributed.destroy_process_group()
'''


def test_stats_allreduce_gloo_world2(tmp_path):
    """N>1 path on CPU: two ranks, gloo; every rank ends with the same reduced vector and so the
    same curriculum index."""
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "stage 1" in o


def test_reference_sac_rule_matches_the_reference_function():
    """sac.reference_update against the reference's own `_update_sac` (agent/multi_algorithm_agent.py:950-1016), which
    tests/golden/make_sac_update_golden.py executed unmodified on networks of this package's shapes: from the same initial
    parameters, batch and generator state, three consecutive updates give the same losses and the same parameters."""
    import torch

    from tvc_ai_b200.sac import Actor, SACConfig, _mlp, reference_update
    z = np.load(os.path.join(ROOT, "tests", "golden", "sac_update.npz"))
    hidden = z["init/actor/net.0.weight"].shape[0]
    actor = Actor(hidden=hidden)
    q1, q2, q1t, q2t = (_mlp(12, 1, hidden) for _ in range(4))
    nets = dict(actor=actor, q1=q1, q2=q2, q1t=q1t, q2t=q2t)
    for n, net in nets.items():
        net.load_state_dict({k: torch.from_numpy(z[f"init/{n}/{k}"]) for k in net.state_dict()})
    cfg = SACConfig.reference_rule()
    assert (cfg.lr_actor, cfg.lr_critic, cfg.gamma, cfg.tau, cfg.ent_coef, cfg.grad_clip_norm, cfg.reward_scale) == \
        (3e-4, 3e-4, 0.99, 0.005, 0.2, 0.0, 1.0)
    opt_a = torch.optim.Adam(actor.parameters(), lr=cfg.lr_actor)
    opt_c = torch.optim.Adam(list(q1.parameters()) + list(q2.parameters()), lr=cfg.lr_critic)
    batch = dict(obs=torch.from_numpy(z["batch/states"]), actions=torch.from_numpy(z["batch/actions"]),
                 reward=torch.from_numpy(z["batch/rewards"]), next_obs=torch.from_numpy(z["batch/next_states"]),
                 done=torch.from_numpy(z["batch/dones"]))
    worst = 0.0
    for u in range(3):
        torch.manual_seed(1000 + u)
        l1, l2, la = reference_update(actor, q1, q2, q1t, q2t, opt_a, opt_c, batch, gamma=cfg.gamma, alpha=float(cfg.ent_coef),
                                      tau=cfg.tau)
        np.testing.assert_allclose([float(l1), float(l2), float(la)], z[f"loss/{u}"], rtol=2e-6, atol=1e-7)
        for n, net in nets.items():
            for k, v in net.state_dict().items():
                want = z[f"after{u}/{n}/{k}"]
                worst = max(worst, float(np.abs(v.numpy() - want).max()))
                np.testing.assert_allclose(v.numpy(), want, rtol=1e-5, atol=2e-7, err_msg=f"update {u} {n}/{k}")
    # the parameters moved (three Adam steps of 3e-4) and the targets lag behind the critics
    assert float(np.abs(q1.state_dict()["0.weight"].numpy() - z["init/q1/0.weight"]).max()) > 5e-4
    assert float(np.abs(q1t.state_dict()["0.weight"].numpy() - z["init/q1t/0.weight"]).max()) < 2e-5
    print(f"reference SAC rule: worst parameter difference after 3 updates {worst:.2e}")


def test_reference_arm_line_and_evidence_hash():
    """`bench.py --impl reference` (the oracle port on the host cores) prints the contract's JSON line without a GPU, and the
    committed ncu figures of the step path (profiles/step_kernel_traffic.json) were captured from the kernel sources in the tree
    (otherwise bench.py drops `roofline.traffic` / `issue_roofline`: re-capture with tools/capture_step.sh)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["warmup"] == 3 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and "workload" in line["config"]
    import bench
    rec = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))
    assert rec["csrc_sha16"] == bench.csrc_hash(), "profiles/step_kernel_traffic.json is stale against tvc_ai_b200/csrc"
    assert bench._traffic() == rec["dram_bytes_per_launch"] and bench._traffic("warp_instructions_per_step") > 1e7
