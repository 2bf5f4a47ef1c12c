"""Diagnostic: how densely the contact solve runs in the groups of each work class (library built with -DTVC_SOLVE_STATS, e.g.
`tools/mkvariant.sh ss -DTVC_SOLVE_STATS`).  Per class (by position in the sequence) and substep: groups, groups in which at least
one lane ran the solve, lanes that ran it.  Usage: TVC_B200_LIB=$PWD/variants/ss.so python tools/solve_stats.py [envs]"""
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
out = (C.c_ulonglong * (3 * 16 * 3))()
L.tvc_debug_solve_stats(out, 1)
S = 20
for t in range(S):
    eng.step(acts[t % 8], want_final=False)
L.tvc_debug_solve_stats(out, 0)
v = list(out)
for cls, name in enumerate(("class 0 (touching)", "class 1 (may touch)", "class 2 (airborne)")):
    rows = [v[(cls * 16 + k) * 3:(cls * 16 + k) * 3 + 3] for k in range(10)]
    groups = rows[0][0] / S
    ws, ls = sum(r[1] for r in rows) / S, sum(r[2] for r in rows) / S
    print(f"SS {name}: {groups:.0f} groups/step, warp-solves/step {ws:.0f} ({ws / max(groups, 1):.2f} per group), "
          f"lane-solves/step {ls:.0f} ({ls / max(ws, 1):.1f} lanes per warp-solve)")
    print("SS    per substep: warps solving " + " ".join(f"{r[1] / max(r[0], 1):.2f}" for r in rows))
    print("SS    per substep: lanes/warp-solve " + " ".join(f"{r[2] / max(r[1], 1):.1f}" for r in rows))
eng.close()
