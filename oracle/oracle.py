"""ctypes binding of the CPU ORACLE (oracle/tvc_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package (tvc_ai_b200) must never import this.
Parity status: Bullet rows unpinned (PyBullet unavailable), env rows pinned against the
reference's own Python class (tests/golden/make_golden.py); see tvc_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtvc_oracle.so")
_LIB_F32_PATH = os.path.join(_HERE, "libtvc_oracle_f32.so")   # physics layer in float32 (sensitivity probe)
_LIB_F32G_PATH = os.path.join(_HERE, "libtvc_oracle_f32g.so")  # ... with the kernel's precision split (-DORC_GEO_DOUBLE)

MAX_DELAY = 4
HIST = 1000
NCOMP = 12
NSTATS = 16

CONTRACT_R, CONTRACT_X = 0, 1
DIV_OFF, DIV_FAST, DIV_EXACT = 0, 1, 2
Q_DOUBLE_GRAVITY, Q_KEEP_CRITERIA, Q_KEEP_REWARD_HIST, Q_LAGGED_PHASE = 1, 2, 4, 8
Q_THRUST_VECTOR, Q_FROZEN_FORCES, Q_DRAG_CUTOFF, Q_STACKED_DAMPING, Q_EULER_TILT = 1 << 4, 1 << 5, 1 << 6, 1 << 7, 1 << 8
Q_DIVERSITY_BONUS, Q_VARIANCE_PENALTY, Q_CLIP_BEFORE_CURIOSITY = 1 << 9, 1 << 10, 1 << 11
Q_SUCCESS_MASKS_TRUNCATION, Q_CRASH_IS_COM_HEIGHT = 1 << 12, 1 << 13
Q_ALL_REFERENCE = 0x3FFF
Q_CONTRACT_X = Q_ALL_REFERENCE & ~(Q_KEEP_CRITERIA | Q_KEEP_REWARD_HIST)

COMP_NAMES = ("mission_completion", "safety_compliance", "fuel_efficiency", "stability_bonus",
              "control_smoothness", "altitude_maintenance", "crash_penalty", "excessive_tilt",
              "control_saturation", "adjustment", "total_unclipped", "diversity_flag")
STAT_NAMES = ("episodes", "sum_return", "sum_return_sq", "sum_length", "successes", "term_crash",
              "term_tilt", "term_altitude", "term_range", "truncations", "safety_violations",
              "sum_final_altitude", "sum_final_tilt", "sum_fuel_left", "steps", "reserved")
PHASES = ("boost", "coast", "landing", "touchdown", "hover", "complete", "failed")


class BodyParams(C.Structure):
    _fields_ = [("mass", C.c_double), ("inertia", C.c_double * 3), ("lin_damp", C.c_double),
                ("ang_damp", C.c_double), ("use_gyro", C.c_int32), ("substeps", C.c_int32),
                ("gravity", C.c_double * 3), ("dt_step", C.c_double), ("max_vel", C.c_double),
                ("ground", C.c_int32), ("contact_iters", C.c_int32), ("warm_iters", C.c_int32), ("reserved_", C.c_int32),
                ("radius", C.c_double),
                ("half_len", C.c_double), ("cg", C.c_double), ("mu", C.c_double), ("mu_spin", C.c_double),
                ("mu_roll", C.c_double), ("restitution", C.c_double), ("rest_threshold", C.c_double),
                ("erp", C.c_double), ("margin", C.c_double)]


class Body(C.Structure):
    _fields_ = [("pos", C.c_double * 3), ("quat", C.c_double * 4), ("vel", C.c_double * 3),
                ("omega", C.c_double * 3), ("force", C.c_double * 3), ("torque", C.c_double * 3),
                ("thrust_local", C.c_double * 3), ("thrust_arm", C.c_double)]


class Config(C.Structure):
    _fields_ = [("contract", C.c_int32), ("substeps", C.c_int32), ("max_episode_steps", C.c_int32),
                ("autoreset", C.c_int32), ("quirks", C.c_uint32), ("diversity_mode", C.c_int32),
                ("contact_iters", C.c_int32), ("ground", C.c_int32), ("contact_warm_iters", C.c_int32), ("reserved_", C.c_int32),
                ("dt_step", C.c_double),
                ("gradient_penalty", C.c_double), ("diversity_bonus", C.c_double),
                ("mass", C.c_double), ("radius", C.c_double), ("length", C.c_double), ("thrust", C.c_double),
                ("gimbal_max_rad", C.c_double), ("lin_damp", C.c_double), ("ang_damp", C.c_double),
                ("mass_variation", C.c_double), ("thrust_std", C.c_double), ("thrust_lo", C.c_double),
                ("thrust_hi", C.c_double), ("cg_offset_max", C.c_double), ("wind_std", C.c_double),
                ("sensor_noise_std", C.c_double), ("init_tilt_max", C.c_double), ("init_omega_max", C.c_double),
                ("propellant_fraction", C.c_double), ("cg_burn_shift", C.c_double),
                ("delay_steps", C.c_int32), ("thrust_curve", C.c_int32),
                ("seed", C.c_uint64), ("env_id_base", C.c_int64),
                ("contact_mu", C.c_double), ("contact_mu_spin", C.c_double), ("contact_mu_roll", C.c_double),
                ("contact_restitution", C.c_double), ("contact_rest_threshold", C.c_double), ("contact_erp", C.c_double),
                ("contact_margin", C.c_double)]


class Env(C.Structure):
    _fields_ = [("body", Body), ("fuel", C.c_double), ("burn", C.c_int32), ("step", C.c_int32),
                ("phase", C.c_int32), ("success", C.c_int32), ("has_prev", C.c_int32),
                ("prev_action", C.c_float * 2), ("consec", C.c_int32), ("crit_pushes", C.c_int32),
                ("hist_count", C.c_int64), ("hist", C.c_double * HIST),
                ("clip_bits", C.c_uint32 * 32), ("run_bits", C.c_uint32 * 32),
                ("n_clip", C.c_int32), ("n_run", C.c_int32), ("ep_return", C.c_double),
                ("episode", C.c_int32), ("mass_scale", C.c_double), ("thrust_scale", C.c_double),
                ("cg_offset", C.c_double), ("wind", C.c_double * 2),
                ("delay_ring", (C.c_float * 2) * MAX_DELAY)]


class StepOut(C.Structure):
    _fields_ = [("obs", C.c_float * 10), ("final_obs", C.c_float * 10), ("reward", C.c_double),
                ("terminated", C.c_int32), ("truncated", C.c_int32), ("comp", C.c_double * NCOMP),
                ("altitude", C.c_double), ("tilt", C.c_double), ("omega_mag", C.c_double),
                ("fuel", C.c_double), ("vh", C.c_double), ("vv", C.c_double),
                ("position", C.c_double * 3), ("phase", C.c_int32), ("success", C.c_int32),
                ("step", C.c_int32), ("criteria_met", C.c_int32), ("term_reason", C.c_int32),
                ("contact_margin", C.c_double)]


def build(force: bool = False, f32: bool = False, geo_double: bool = False) -> str:
    """Compile oracle/libtvc_oracle.so (building the checker is not using it).  f32=True builds the
    variant whose physics layer evaluates in float32 (-DORC_PHYS_FLOAT); with geo_double=True that variant keeps
    the attitude and the cap-centre heights in double (-DORC_GEO_DOUBLE: the CUDA kernel's precision split)."""
    src = os.path.join(_HERE, "tvc_oracle.c")
    hdr = os.path.join(_HERE, "tvc_oracle.h")
    out = (_LIB_F32G_PATH if geo_double else _LIB_F32_PATH) if f32 else _LIB_PATH
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return out
    base = ["-O2", "-fPIC", "-std=c11", "-ffp-contract=off", "-fno-fast-math", "-shared", "-o", out, src, "-lm"]
    if f32:
        base = ["-DORC_PHYS_FLOAT"] + (["-DORC_GEO_DOUBLE"] if geo_double else []) + base
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True)
                return out
            except (OSError, subprocess.CalledProcessError) as exc:  # try the next compiler / no OpenMP
                last = exc
    raise RuntimeError(f"could not build the oracle: {last}")


_lib = None
_lib_f32 = None
_lib_f32g = None


def lib(f32: bool = False, geo_double: bool = False):
    global _lib, _lib_f32, _lib_f32g
    if f32 and geo_double:
        if _lib_f32g is None:
            _lib_f32g = _declare(C.CDLL(build(f32=True, geo_double=True)))
        return _lib_f32g
    if f32:
        if _lib_f32 is None:
            _lib_f32 = _declare(C.CDLL(build(f32=True)))
        return _lib_f32
    if _lib is not None:
        return _lib
    _lib = _declare(C.CDLL(build()))
    return _lib


def _declare(L):
    dp = C.POINTER(C.c_double)
    L.orc_body_params_default.argtypes = [C.POINTER(BodyParams)]
    L.orc_body_init.argtypes = [C.POINTER(Body), dp, dp]
    L.orc_apply_external_force.argtypes = [C.POINTER(Body), dp, dp]
    L.orc_apply_external_torque.argtypes = [C.POINTER(Body), dp]
    L.orc_step_simulation.argtypes = [C.POINTER(BodyParams), C.POINTER(Body), dp]
    L.orc_reported_quat.argtypes = [dp, dp]
    L.orc_euler_from_quat.argtypes = [dp, dp]
    L.orc_matrix_from_quat.argtypes = [dp, dp]
    L.orc_config_default.argtypes = [C.POINTER(Config), C.c_int]
    L.orc_create.argtypes = [C.POINTER(Config), C.c_int64]
    L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_num_envs.argtypes = [C.c_void_p]
    L.orc_num_envs.restype = C.c_int64
    L.orc_env_ptr.argtypes = [C.c_void_p, C.c_int64]
    L.orc_env_ptr.restype = C.POINTER(Env)
    L.orc_config_ptr.argtypes = [C.c_void_p]
    L.orc_config_ptr.restype = C.POINTER(Config)
    L.orc_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(StepOut), C.c_int]
    L.orc_step_arrays.argtypes = [C.c_void_p] + [C.c_void_p] * 6 + [C.c_int]
    L.orc_random_actions.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    L.orc_stats.argtypes = [C.c_void_p, dp, C.c_int]
    L.orc_last_trace.argtypes = [C.c_void_p]
    L.orc_last_trace.restype = dp
    L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.orc_fuel_table.argtypes = [C.c_int]
    L.orc_fuel_table.restype = C.c_double
    L.orc_thrust_curve.argtypes = [C.c_int, C.c_int]
    L.orc_thrust_curve.restype = C.c_double
    return L


def _d(seq):
    return (C.c_double * len(seq))(*[float(x) for x in seq])


def default_config(contract: int = CONTRACT_R, **over) -> Config:
    cfg = Config()
    lib().orc_config_default(C.byref(cfg), contract)
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def reported_quat(q):
    o = (C.c_double * 4)()
    lib().orc_reported_quat(_d(q), o)
    return tuple(o)


def euler_from_quat(q):
    o = (C.c_double * 3)()
    lib().orc_euler_from_quat(_d(q), o)
    return tuple(o)


def matrix_from_quat(q):
    o = (C.c_double * 9)()
    lib().orc_matrix_from_quat(_d(q), o)
    return tuple(o)


class OracleSim:
    """Batch of oracle envs (fp64).  Mirrors the device engine's reset/step surface."""

    def __init__(self, cfg: Config | None = None, num_envs: int = 1, f32_physics: bool = False, geo_double: bool = False, **over):
        self.L = lib(f32_physics, geo_double)
        self.cfg = cfg if cfg is not None else default_config(**over)
        self.n = int(num_envs)
        self.h = self.L.orc_create(C.byref(self.cfg), self.n)
        if not self.h:
            raise MemoryError("orc_create failed")

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def config(self) -> Config:
        return self.L.orc_config_ptr(self.h).contents

    def env(self, i: int = 0) -> Env:
        return self.L.orc_env_ptr(self.h, i).contents

    def reset(self, mask=None):
        obs = np.zeros((self.n, 10), np.float32)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        self.L.orc_reset(self.h, None if mask is None else mask.ctypes.data, obs.ctypes.data)
        return obs

    def step(self, actions, threads: int = 1):
        """Returns (obs, reward, terminated, truncated, outs) with outs the per-env StepOut array."""
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, 2)
        outs = (StepOut * self.n)()
        self.L.orc_step(self.h, a.ctypes.data, outs, threads)
        obs = np.stack([np.frombuffer(o.obs, np.float32) for o in outs]) if self.n <= 4096 else None
        rew = np.array([o.reward for o in outs])
        term = np.array([o.terminated for o in outs], bool)
        trunc = np.array([o.truncated for o in outs], bool)
        return obs, rew, term, trunc, outs

    def step_arrays(self, actions, threads: int = 1, want_final=False):
        a = np.ascontiguousarray(actions, np.float32).reshape(self.n, 2)
        obs = np.empty((self.n, 10), np.float32)
        rew = np.empty(self.n, np.float64)
        term = np.empty(self.n, np.uint8)
        trunc = np.empty(self.n, np.uint8)
        fin = np.zeros((self.n, 10), np.float32) if want_final else None
        self.L.orc_step_arrays(self.h, a.ctypes.data, obs.ctypes.data, rew.ctypes.data, term.ctypes.data,
                               trunc.ctypes.data, None if fin is None else fin.ctypes.data, threads)
        return obs, rew, term.astype(bool), trunc.astype(bool), fin

    def random_actions(self, t: int):
        a = np.empty((self.n, 2), np.float32)
        self.L.orc_random_actions(self.h, int(t), a.ctypes.data)
        return a

    def stats(self, reset_after=False):
        o = (C.c_double * NSTATS)()
        self.L.orc_stats(self.h, o, int(reset_after))
        return np.array(o)

    def last_trace(self):
        k = self.config.substeps
        p = self.L.orc_last_trace(self.h)
        return np.ctypeslib.as_array(p, shape=(k, 13)).copy()

    # ---- state exchange with the device engine (tests) ----
    def core_state(self, i: int = 0):
        e = self.env(i)
        return dict(pos=np.array(e.body.pos), quat=np.array(e.body.quat), vel=np.array(e.body.vel),
                    omega=np.array(e.body.omega), burn=e.burn, step=e.step, phase=e.phase, success=e.success,
                    has_prev=e.has_prev, prev_action=np.array(e.prev_action), consec=e.consec,
                    hist_count=e.hist_count, episode=e.episode, ep_return=e.ep_return,
                    mass_scale=e.mass_scale, thrust_scale=e.thrust_scale, cg_offset=e.cg_offset,
                    wind=np.array(e.wind))
