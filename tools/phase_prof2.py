"""Diagnostic: clock64 phase profile of step_kernel_v2 (library built with -DTVC_PHASE_PROF2, e.g.
`tools/mkvariant.sh prof2 -DTVC_PHASE_PROF2`).  Usage: TVC_B200_LIB=$PWD/variants/prof2.so python tools/phase_prof2.py [envs]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset()
acts = [torch.rand((n, 2), device="cuda") * 2 - 1 for _ in range(8)]
for t in range(410):
    eng.step(acts[t % 8], want_final=False)
L = C.CDLL(os.environ["TVC_B200_LIB"])
out = (C.c_ulonglong * 16)()
L.tvc_debug_phase2(out, 1)
S = 20
for t in range(S):
    eng.step(acts[t % 8], want_final=False)
L.tvc_debug_phase2(out, 0)
v = list(out)
names = ["pull (barrier+queue)", "index+loads+env_pre", "substeps w/o solver", "solver entry", "sweeps", "env_post+stores",
         "groups", "substep barrier wait"]
for cls, cn in ((0, "near-ground groups"), (1, "airborne groups")):
    r = v[cls * 8:cls * 8 + 8]
    g = max(r[6], 1)
    tot = sum(r[k] for k in (0, 1, 2, 3, 4, 5, 7))
    print(f"PH2 n={n} {cn}: {r[6] / S:.0f} groups/step, {tot / g:.0f} cycles per group")
    for k in (0, 1, 2, 3, 4, 5, 7):
        print(f"PH2   {names[k]:24s} {r[k] / g:10.0f} cycles/group  {100 * r[k] / max(tot, 1):5.1f} %")
eng.close()
