"""Freezes the reference's own SAC update, agent/multi_algorithm_agent.py:950-1016 `MultiAlgorithmAgent._update_sac`, on small
networks of this package's shapes (run in the build container, where /root/reference is readable; CPU, fp32):

    python tests/golden/make_sac_update_golden.py      ->  tests/golden/sac_update.npz

The reference FUNCTION is executed unmodified (imported from /root/reference) on a stand-in `self` whose `algorithms['sac']`
dict holds tvc_ai_b200.sac networks (hidden width 32, so the fixture stays small) and the optimisers the reference creates
(:623-625: three Adam(3e-4)).  The fixture holds the initial parameters, the batch, and -- for three consecutive updates, each
after `torch.manual_seed(1000 + k)` -- the three losses the function returns and every parameter afterwards.
tests/test_host.py::test_reference_sac_rule_matches_the_reference_function replays them through sac.reference_update."""
import os
import sys
import types

import numpy as np
import torch
import torch.optim as optim

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from agent.multi_algorithm_agent import MultiAlgorithmAgent  # noqa: E402  (the reference)

from tvc_ai_b200.sac import Actor, _mlp  # noqa: E402

HIDDEN, BATCH, UPDATES = 32, 64, 3


class PolicyAdapter(torch.nn.Module):
    """(mean, log_std, value) as the reference's policy networks return them (:223-227)."""

    def __init__(self, actor):
        super().__init__()
        self.actor = actor

    def forward(self, x):
        m, ls = self.actor.mean_log_std(x)
        return m, ls, None


def main():
    torch.manual_seed(7)
    actor = Actor(hidden=HIDDEN)
    q1, q2, q1t, q2t = (_mlp(12, 1, HIDDEN) for _ in range(4))
    q1t.load_state_dict(q1.state_dict()), q2t.load_state_dict(q2.state_dict())
    nets = dict(actor=actor, q1=q1, q2=q2, q1t=q1t, q2t=q2t)
    out = {}
    for n, net in nets.items():
        for k, v in net.state_dict().items():
            out[f"init/{n}/{k}"] = v.numpy().copy()
    g = torch.Generator().manual_seed(11)
    batch = dict(states=torch.randn(BATCH, 10, generator=g), actions=torch.rand(BATCH, 2, generator=g) * 2 - 1,
                 rewards=torch.randn(BATCH, generator=g) * 30.0 + 20.0, next_states=torch.randn(BATCH, 10, generator=g),
                 dones=(torch.rand(BATCH, generator=g) < 0.1).float())
    for k, v in batch.items():
        out[f"batch/{k}"] = v.numpy().copy()
    fake = types.SimpleNamespace(algorithms={"sac": {
        "policy": PolicyAdapter(actor), "q1": q1, "q2": q2, "target_q1": q1t, "target_q2": q2t,
        "optimizer_policy": optim.Adam(actor.parameters(), lr=3e-4),
        "optimizer_q1": optim.Adam(q1.parameters(), lr=3e-4), "optimizer_q2": optim.Adam(q2.parameters(), lr=3e-4)}})
    for u in range(UPDATES):
        torch.manual_seed(1000 + u)
        res = MultiAlgorithmAgent._update_sac(fake, batch)
        out[f"loss/{u}"] = np.array([res["q1_loss"], res["q2_loss"], res["policy_loss"]], np.float64)
        for n, net in nets.items():
            for k, v in net.state_dict().items():
                out[f"after{u}/{n}/{k}"] = v.numpy().copy()
    path = os.path.join(ROOT, "tests", "golden", "sac_update.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "losses", [out[f"loss/{u}"].tolist() for u in range(UPDATES)])


if __name__ == "__main__":
    main()
