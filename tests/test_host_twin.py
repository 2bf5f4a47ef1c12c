"""CPU tests of the DEVICE source: tvc_ai_b200/csrc/tvc_device.cuh compiled by g++ for the host (tests/host_twin) and
compared with the fp64 oracle on the golden trajectories and on a Contract-X batch, teacher-forced from identical float32
inputs -- the same comparisons as tests/test_gpu_parity.py, minus what only the GPU build has (nvcc's FMA contraction, the
MUFU approximations, the warp-level plumbing of the kernels).  A logic difference between the kernel model and the oracle
shows up here, on the CPU box, before any GPU time is spent.  The twin is test infrastructure: the product never links it.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
TWIN_DIR = os.path.join(HERE, "host_twin")
TWIN_LIB = os.path.join(TWIN_DIR, "libtvc_twin.so")


def _cuda_include():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include", "cuda_runtime.h")):
            return os.path.join(d, "include")
    return None


def build_twin(force=False):
    srcs = [os.path.join(TWIN_DIR, "twin.cpp"), os.path.join(TWIN_DIR, "cuda_host_shim.h"),
            os.path.join(ROOT, "tvc_ai_b200", "csrc", "tvc_device.cuh"), os.path.join(ROOT, "tvc_ai_b200", "csrc", "tvc_internal.h"),
            os.path.join(ROOT, "include", "tvc_b200.h")]
    if not force and os.path.exists(TWIN_LIB) and os.path.getmtime(TWIN_LIB) >= max(os.path.getmtime(s) for s in srcs):
        return TWIN_LIB
    inc = _cuda_include()
    if inc is None:
        pytest.skip("CUDA headers not found (the twin needs cuda_runtime.h for float4 & co.)")
    # -ffp-contract=off: no FMA contraction on the host, so the twin differs from the oracle by float32 rounding only
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-attributes", "-I", inc,
                    "-o", TWIN_LIB, srcs[0]], check=True, capture_output=True)
    return TWIN_LIB


class Twin:
    def __init__(self, cfg, n):
        from tvc_ai_b200 import _abi as A
        from tvc_ai_b200.engine import STATE_DTYPE
        self.A, self.dtype, self.n = A, STATE_DTYPE, n
        T = C.CDLL(build_twin())
        T.twin_create.restype = C.c_void_p
        T.twin_create.argtypes = [C.POINTER(A.TvcConfig), C.c_longlong]
        for f, args in (("twin_reset", 2), ("twin_set_state", 2), ("twin_get_state", 2), ("twin_step", 8), ("twin_destroy", 1)):
            getattr(T, f).restype = None
            getattr(T, f).argtypes = [C.c_void_p] * args
        T.twin_set_step_counter.argtypes = [C.c_void_p, C.c_ulonglong]
        self.T, self.h = T, T.twin_create(C.byref(cfg), n)
        self.obs = np.zeros((n, 10), np.float32)
        self.final = np.zeros((n, 10), np.float32)
        self.rew = np.zeros(n, np.float32)
        self.term, self.trunc = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        self.comp = np.zeros((n, 12), np.float32)
        self.t = 0

    def reset(self):
        self.T.twin_reset(self.h, self.obs.ctypes.data)
        return self.obs.copy()

    def get_state(self):
        st = np.zeros(self.n, self.dtype)
        self.T.twin_get_state(self.h, st.ctypes.data)
        return st

    def set_state(self, st):
        st = np.ascontiguousarray(st, self.dtype)
        self.T.twin_set_state(self.h, st.ctypes.data)

    def step(self, actions):
        a = None if actions is None else np.ascontiguousarray(actions, np.float32)
        self.T.twin_set_step_counter(self.h, self.t)
        self.T.twin_step(self.h, None if a is None else a.ctypes.data, self.obs.ctypes.data, self.rew.ctypes.data,
                         self.term.ctypes.data, self.trunc.ctypes.data, self.final.ctypes.data, self.comp.ctypes.data)
        self.t += 1
        return self.obs.copy(), self.rew.copy(), self.term.astype(bool), self.trunc.astype(bool)

    def close(self):
        self.T.twin_destroy(self.h)


def _helpers():
    # the GPU parity module's helpers (pure numpy / ctypes); importing it does not need a GPU
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gpu_parity_helpers", os.path.join(HERE, "test_gpu_parity.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["zero_120", "random_raw", "random_autoreset", "two_episodes", "burnout_1100", "crash_leak"])
def test_twin_golden_trajectories_teacher_forced(oracle_mod, lib_built, golden_dir, name):
    """Contract R, N = 1, every step of every golden file: state after the step within K * 1e-5 in free flight and
    CONTACT_MAX_R in contact, rewards 2e-4, flags / phase / counters exact outside near-threshold events."""
    O, H = oracle_mod, _helpers()
    from tvc_ai_b200 import _abi as A
    g = np.load(os.path.join(golden_dir, f"{name}.npz"))
    T, K = len(g["reward"]), 4
    sim, sh = O.OracleSim(O.default_config(O.CONTRACT_R), 1), O.OracleSim(O.default_config(O.CONTRACT_R), 1)
    tw = Twin(A.default_config(A.CONTRACT_R), 1)
    sim.reset(), sh.reset(), tw.reset()
    worst_free = worst_contact = worst_rew = 0.0
    near = bad = 0
    for t in range(T):
        C.memmove(C.byref(sh.env(0)), C.byref(sim.env(0)), C.sizeof(O.Env))
        H._round_body_f32(sh.env(0))
        tw.set_state(H._state_from_oracle(O, sh, tw.get_state()))
        pre = H._body13(sh.env(0))
        a = g["actions"][t:t + 1]
        sim.step(a)
        _, r_o, _, _, outs = sh.step(a)
        o = outs[0]
        obs_d, rew_d, term_d, trunc_d = tw.step(a)
        st = tw.get_state()[0]
        ref = H._body13(sh.env(0))
        dev = np.concatenate([st["pos"], st["quat"], st["vel"], st["omega"]])
        err = float((np.abs(dev - ref) / np.maximum(1.0, np.abs(ref))).max())
        contact = H._in_contact(pre, ref)
        assert err <= (H.CONTACT_MAX_R if contact else K * 1e-5), (name, t, contact, err)
        if contact:
            worst_contact = max(worst_contact, err)
        else:
            worst_free = max(worst_free, err)
        flags_equal = (bool(term_d[0]) == bool(o.terminated) and bool(trunc_d[0]) == bool(o.truncated)
                       and int(st["phase"]) == o.phase and bool(st["success"]) == bool(o.success) and int(st["step"]) == o.step)
        if not flags_equal:
            near += int(H._near_threshold(o))
            bad += int(not H._near_threshold(o))
        elif not H._near_threshold(o) and abs(o.comp[10]) < 900 and tw.comp[0][11] == o.comp[11]:
            rerr = abs(float(rew_d[0]) - r_o[0]) / max(1.0, abs(r_o[0]))
            worst_rew = max(worst_rew, rerr)
            assert rerr <= 2e-4, (name, t, rew_d[0], r_o[0])
        if g["was_reset"][t]:
            sim.reset(), tw.reset()
    print(f"\n[twin {name}] free {worst_free:.2e} contact {worst_contact:.2e} reward {worst_rew:.2e} near {near}")
    assert bad == 0 and near <= max(2, T // 100)
    tw.close()


def test_twin_contract_x_batch(oracle_mod, lib_built):
    """Contract X, 256 envs x 60 steps with DR, sensor noise, delay ring, thrust curve and the in-model Philox actions."""
    O, H = oracle_mod, _helpers()
    from tvc_ai_b200 import _abi as A
    n, K = 256, 10
    over = dict(init_tilt_max=0.2, init_omega_max=0.1, delay_steps=3, thrust_curve=1, propellant_fraction=0.2, cg_burn_shift=0.05,
                autoreset=1, env_id_base=1000)
    sim = O.OracleSim(O.default_config(O.CONTRACT_X, **over), n)
    tw = Twin(A.default_config(A.CONTRACT_X, **over), n)
    o_t, o_o = tw.reset(), sim.reset()
    np.testing.assert_allclose(o_t, o_o, rtol=0, atol=2e-6)
    free, cont = [], []
    bad = near = 0
    for t in range(60):
        for i in range(n):
            H._round_body_f32(sim.env(i))
        tw.set_state(H._state_from_oracle(O, sim, tw.get_state()))
        pre = np.array([H._body13(sim.env(i)) for i in range(n)])
        acts = sim.random_actions(t)
        obs_o, rew_o, term_o, trunc_o, outs = sim.step(acts, threads=4)
        fin_o = np.stack([np.frombuffer(o.final_obs, np.float32) for o in outs])
        obs_d, rew_d, term_d, trunc_d = tw.step(None)
        done_o = term_o | trunc_o
        mism = (term_d != term_o) | (trunc_d != trunc_o)
        for i in np.flatnonzero(mism):
            near += int(H._near_threshold(outs[i]))
            bad += int(not H._near_threshold(outs[i]))
        ok = ~mism
        cmp_d = np.where(done_o[:, None], tw.final, obs_d)
        cmp_o = np.where(done_o[:, None], fin_o, obs_o)
        err = (np.abs(cmp_d - cmp_o) / np.maximum(1.0, np.abs(cmp_o))).max(axis=1)
        contact = np.array([min(H._lowest_gap(pre[i][:3], pre[i][3:7], h=0.6), outs[i].position[2] - 0.65) < 0.06 for i in range(n)])
        free.append(err[ok & ~contact]), cont.append(err[ok & contact])
    fe, ce = np.concatenate(free), np.concatenate(cont)
    print(f"\n[twin X] free max {fe.max():.2e} ({len(fe)}), contact q99 {np.quantile(ce, 0.99):.2e} max {ce.max():.2e} ({len(ce)}), near {near}")
    assert len(ce) > 500 and fe.max() <= K * 1e-5
    assert np.quantile(ce, 0.99) <= K * 1e-5 and ce.max() <= H.CONTACT_MAX_X
    assert bad == 0 and near <= 4
    tw.close()
