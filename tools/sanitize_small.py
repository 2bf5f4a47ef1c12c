"""Small end-to-end run for compute-sanitizer (memcheck): both step kernels, reset, state exchange, stats, rollout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

for contract, over in ((A.CONTRACT_X, dict(autoreset=1, delay_steps=2, thrust_curve=1, init_tilt_max=0.3)),
                       (A.CONTRACT_R, dict(autoreset=1))):
    n = 1000
    eng = BatchedEngine(n, A.default_config(contract, **over), device=0)
    eng.reset()
    for t in range(60):
        eng.step(None)
    eng.step_ex(torch.zeros((n, 2), device="cuda"))
    st = eng.get_state()
    eng.set_state(st)
    eng.read_info()
    print(contract, eng.stats()[:5])
    if contract == A.CONTRACT_X:
        torch.manual_seed(0)
        w = dict(w1=torch.randn(256, 10, device="cuda") * 0.3, b1=torch.zeros(256, device="cuda"),
                 w2=torch.randn(256, 256, device="cuda") * 0.06, b2=torch.zeros(256, device="cuda"),
                 w3=torch.randn(4, 256, device="cuda") * 0.06, b3=torch.zeros(4, device="cuda"))
        out = eng.rollout(w, 8, record=True)
        torch.cuda.synchronize()
        print("rollout reward_sum mean", float(out["reward_sum"].mean()))
    eng.close()
print("sanitize run ok")
