"""Diagnostic: where the end-to-end (host buffers) step time goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine
from tvc_ai_b200.vector_env import RocketTVCVectorEnv

n = 262144
venv = RocketTVCVectorEnv(n, config={}, contract="X", device=0, final_info=False, copy_outputs=False)
venv.reset(seed=42)
eng = venv.engine
dev = torch.device("cuda", 0)
pool = [torch.rand((n, 2), device=dev) * 2 - 1 for _ in range(8)]
for b in range(400):
    eng.step(pool[b % 8], want_final=False)
torch.cuda.synchronize()
acts = [np.random.default_rng(i).uniform(-1, 1, (n, 2)).astype(np.float32) for i in range(4)]
pacts = [torch.from_numpy(a).pin_memory().numpy() for a in acts]

def timeit(name, fn, reps=30):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps): fn(k)
    torch.cuda.synchronize()
    print(f"E2E {name:46s} {1e3 * (time.perf_counter() - t0) / reps:.3f} ms", flush=True)

timeit("venv.step(pageable numpy)", lambda k: venv.step(acts[k % 4]))
timeit("venv.step(pinned numpy)", lambda k: venv.step(pacts[k % 4]))
timeit("engine.step_host(actions, want_final=True)", lambda k: eng.step_host(acts[k % 4], want_final=True))
timeit("engine.step_host(actions, want_final=False)", lambda k: eng.step_host(acts[k % 4], want_final=False))
timeit("engine.step_host(None, want_final=False)", lambda k: eng.step_host(None, want_final=False))
def dev_step(k):
    eng.step(pool[k % 8], want_final=False); torch.cuda.synchronize()
timeit("engine.step(device actions) + sync", dev_step)

# slab-pipelined host env
from tvc_ai_b200.vector_env import RocketTVCHostPipelineEnv
venv.close()
for slabs in ((1, 2) if os.environ.get('E2E_PROBE_SHORT') else (1, 2, 3, 4)):
    env = RocketTVCHostPipelineEnv(n, config={}, contract="X", device=0, slabs=slabs)
    env.reset(seed=42)
    for e_, (lo, hi) in zip(env.engines, env._ranges):
        for b in range(400):
            e_.step(pool[b % 8][lo:hi].contiguous(), want_final=False)
    torch.cuda.synchronize()
    buf = env.pinned_actions()
    buf[...] = acts[0]
    timeit(f"RocketTVCHostPipelineEnv(slabs={slabs}).step(pinned)", lambda k: env.step(buf))
    env.close()
