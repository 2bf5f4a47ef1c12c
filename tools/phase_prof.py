import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['TVC_B200_LIB'] = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tvc_ai_b200', 'libtvc_b200_prof.so')
import torch

from tvc_ai_b200 import build as _B
_B.build_phase_profiler()
os.environ['TVC_STEP_IMPL'] = '1'   # the counters live in the legacy CTA-exchange kernel
from tvc_ai_b200 import _abi as A
from tvc_ai_b200.engine import BatchedEngine
n = 262144
eng = BatchedEngine(n, A.default_config(A.CONTRACT_X, autoreset=1), device=0)
eng.reset()
acts = [torch.rand((n, 2), device='cuda') * 2 - 1 for _ in range(8)]
for t in range(400): eng.step(acts[t % 8], want_final=False)
L = A.load()
out = (C.c_ulonglong * 8)()
L.tvc_debug_phase(out, 1)
for t in range(20): eng.step(acts[t % 8], want_final=False)
L.tvc_debug_phase(out, 0)
v = list(out)
names = ['free-flight', 'count barrier', 'post+barrier', 'solve (solver warps)', 'wait-for-solver', 'readback+pose']
nw, ns = v[6], v[7]
print('warp-substeps', nw, 'solver warp-substeps', ns, 'fraction', ns / nw)
for i, nm in enumerate(names):
    den = ns if i == 3 else nw
    print(f'{nm:24s} avg cycles per warp-substep {v[i] / max(den,1):10.1f}   (total share {v[i] / sum(v[:6]) * 100:5.1f}%)')
