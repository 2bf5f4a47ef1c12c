"""Diagnostic: end-to-end step with the kernels storing the results straight into pinned host memory (no D2H copy)
against the staged path (tvc_step_host: one 46 B/env copy).  Usage: python tools/e2e_direct_probe.py [envs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda", 0)
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset(seed=42)
pool = [torch.rand((n, 2), device=dev) * 2 - 1 for _ in range(8)]
for b in range(400):
    eng.step(pool[b % 8], want_final=False)
torch.cuda.synchronize()
acts = [torch.from_numpy(np.random.default_rng(i).uniform(-1, 1, (n, 2)).astype(np.float32)).pin_memory() for i in range(4)]
nacts = [a.numpy() for a in acts]
adev = torch.empty((n, 2), dtype=torch.float32, device=dev)


def timeit(name, fn, reps=40):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps): fn(k)
    torch.cuda.synchronize()
    print(f"E2ED {name:58s} {1e3 * (time.perf_counter() - t0) / reps:.3f} ms", flush=True)


def dev_events(name, reps=40):
    st = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    en = [torch.cuda.Event(enable_timing=True) for _ in range(reps)]
    for k in range(reps):
        st[k].record(); eng.step(pool[k % 8], want_final=False); en[k].record()
    torch.cuda.synchronize()
    per = sorted(s.elapsed_time(e) for s, e in zip(st, en))
    print(f"E2ED {name:58s} kernels {sum(per) / reps:.4f} ms (median {per[reps // 2]:.4f})", flush=True)


timeit("staged: tvc_step_host(pinned actions)", lambda k: eng.step_host(nacts[k % 4], want_final=False))
dev_events("device outputs")
dobs, drew, dterm, dtrunc = eng.obs, eng.reward, eng.terminated, eng.truncated
slab = torch.zeros(46 * n, dtype=torch.uint8).pin_memory()
hobs = slab[:40 * n].view(torch.float32).view(n, 10)
hrew = slab[40 * n:44 * n].view(torch.float32)
hterm, htrunc = slab[44 * n:45 * n], slab[45 * n:46 * n]


def direct(k):
    adev.copy_(acts[k % 4], non_blocking=True)
    eng.step(adev, want_final=False)
    torch.cuda.synchronize()


for what, (o, r, t, u) in (("obs+reward+flags direct", (hobs, hrew, hterm, htrunc)),
                          ("obs direct, reward+flags on device (not copied)", (hobs, drew, dterm, dtrunc)),
                          ("reward+flags direct, obs on device (not copied)", (dobs, hrew, hterm, htrunc))):
    eng.obs, eng.reward, eng.terminated, eng.truncated = o, r, t, u
    timeit(what, direct)
    dev_events(what)
# correctness: direct results equal staged results from the same state
eng.obs, eng.reward, eng.terminated, eng.truncated = dobs, drew, dterm, dtrunc
state = eng.get_state()
eng.step(pool[0], want_final=False); torch.cuda.synchronize()
ref = (dobs.cpu().clone(), drew.cpu().clone(), dterm.cpu().clone(), dtrunc.cpu().clone())
eng.set_state(state)
eng.obs, eng.reward, eng.terminated, eng.truncated = hobs, hrew, hterm, htrunc
eng.step(pool[0], want_final=False); torch.cuda.synchronize()
same = all(torch.equal(a, b) for a, b in zip(ref, (hobs, hrew, hterm, htrunc)))
print("E2ED direct == staged:", same, flush=True)
eng.obs, eng.reward, eng.terminated, eng.truncated = dobs, drew, dterm, dtrunc
eng.close()
