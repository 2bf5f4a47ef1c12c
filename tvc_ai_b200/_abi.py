"""ctypes binding of libtvc_b200.so (include/tvc_b200.h).

This is the reference-side stub a TVC-AI maintainer would add: plain
pointers and sizes, no torch types across the boundary.  The library is built in-tree by
`tvc_ai_b200.build.build()`; loading fails loudly if it is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TVC_B200_LIB") or os.path.join(_HERE, "libtvc_b200.so")

ABI_VERSION = 8
OBS_DIM, ACT_DIM, NUM_COMPONENTS, NUM_STATS, MAX_DELAY = 10, 2, 12, 16, 4
CONTRACT_R, CONTRACT_X = 0, 1
DIV_OFF, DIV_FAST, DIV_EXACT = 0, 1, 2
# quirk switches (include/tvc_b200.h TVC_Q_*): a set bit reproduces the reference, see the header for the cleared meaning
Q_DOUBLE_GRAVITY, Q_KEEP_CRITERIA, Q_KEEP_REWARD_HIST, Q_LAGGED_PHASE = 1, 2, 4, 8
Q_THRUST_VECTOR, Q_FROZEN_FORCES, Q_DRAG_CUTOFF, Q_STACKED_DAMPING, Q_EULER_TILT = 1 << 4, 1 << 5, 1 << 6, 1 << 7, 1 << 8
Q_DIVERSITY_BONUS, Q_VARIANCE_PENALTY, Q_CLIP_BEFORE_CURIOSITY = 1 << 9, 1 << 10, 1 << 11
Q_SUCCESS_MASKS_TRUNCATION, Q_CRASH_IS_COM_HEIGHT = 1 << 12, 1 << 13
Q_ALL_REFERENCE = 0x3FFF
Q_CONTRACT_X = Q_ALL_REFERENCE & ~(Q_KEEP_CRITERIA | Q_KEEP_REWARD_HIST)
SEED_KEEP = 0xFFFFFFFFFFFFFFFF   # tvc_reset: keep the current Philox key

COMPONENT_NAMES = ("mission_completion", "safety_compliance", "fuel_efficiency", "stability_bonus",
                   "control_smoothness", "altitude_maintenance", "crash_penalty", "excessive_tilt",
                   "control_saturation", "adjustment", "total_unclipped", "diversity_flag")
STAT_NAMES = ("episodes", "sum_return", "sum_return_sq", "sum_length", "successes", "term_crash",
              "term_tilt", "term_altitude", "term_range", "truncations", "safety_violations",
              "sum_final_altitude", "sum_final_tilt", "sum_fuel_left", "steps", "reserved")

EXPORTS = ("tvc_abi_version", "tvc_last_error", "tvc_config_default", "tvc_create", "tvc_destroy", "tvc_reset",
           "tvc_step", "tvc_step_ex", "tvc_step_host", "tvc_step_host_async", "tvc_host_sync", "tvc_rollout", "tvc_state_bytes", "tvc_get_state",
           "tvc_set_state", "tvc_get_reward_history", "tvc_set_reward_history", "tvc_read_info", "tvc_episode_stats", "tvc_episode_stats_dev", "tvc_set_curriculum",
           "tvc_get_config", "tvc_num_envs", "tvc_lifetime_steps", "tvc_replay_sample", "tvc_curiosity")


class TvcConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("contract", C.c_int32), ("substeps", C.c_int32),
                ("max_episode_steps", C.c_int32), ("autoreset", C.c_int32), ("quirks", C.c_uint32),
                ("diversity_mode", C.c_int32), ("contact_iters", C.c_int32), ("ground", C.c_int32),
                ("delay_steps", C.c_int32), ("thrust_curve", C.c_int32), ("contact_warm_iters", C.c_int32),
                ("dt_step", C.c_double),
                ("gradient_penalty", C.c_float), ("diversity_bonus", C.c_float),
                ("mass", C.c_float), ("radius", C.c_float), ("length", C.c_float), ("thrust", C.c_float),
                ("gimbal_max_rad", C.c_float), ("lin_damp", C.c_float), ("ang_damp", C.c_float),
                ("mass_variation", C.c_float), ("thrust_std", C.c_float), ("thrust_lo", C.c_float),
                ("thrust_hi", C.c_float), ("cg_offset_max", C.c_float), ("wind_std", C.c_float),
                ("sensor_noise_std", C.c_float), ("init_tilt_max", C.c_float), ("init_omega_max", C.c_float),
                ("propellant_fraction", C.c_float), ("cg_burn_shift", C.c_float), ("reserved1", C.c_float),
                ("contact_mu", C.c_float), ("contact_mu_spin", C.c_float), ("contact_mu_roll", C.c_float),
                ("contact_restitution", C.c_float), ("contact_rest_threshold", C.c_float), ("contact_erp", C.c_float),
                ("contact_margin", C.c_float), ("reserved2", C.c_float),
                ("seed", C.c_uint64), ("env_id_base", C.c_int64)]


class TvcStageConditions(C.Structure):
    _fields_ = [("max_initial_tilt", C.c_float), ("max_initial_angular_vel", C.c_float),
                ("domain_randomization", C.c_int32), ("sensor_noise", C.c_int32),
                ("max_gimbal_angle_deg", C.c_float), ("wind_enabled", C.c_int32),
                ("wind_force", C.c_float), ("mass_variation", C.c_float)]


class TvcInfoSoa(C.Structure):
    _fields_ = [("altitude", C.c_void_p), ("tilt_deg", C.c_void_p), ("omega_mag", C.c_void_p),
                ("fuel", C.c_void_p), ("position", C.c_void_p), ("phase", C.c_void_p), ("step", C.c_void_p),
                ("success", C.c_void_p), ("criteria_met", C.c_void_p), ("reward_components", C.c_void_p)]


class TvcStepIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p),
                ("terminated", C.c_void_p), ("truncated", C.c_void_p), ("final_obs", C.c_void_p),
                ("actions_out", C.c_void_p), ("info", TvcInfoSoa)]


class TvcEnvState(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("quat", C.c_float * 4), ("vel", C.c_float * 3), ("omega", C.c_float * 3),
                ("prev_action", C.c_float * 2), ("ep_return", C.c_float),
                ("step", C.c_int32), ("burn", C.c_int32), ("phase", C.c_int32), ("success", C.c_int32),
                ("has_prev", C.c_int32), ("consec", C.c_int32), ("hist_count", C.c_int32), ("episode", C.c_int32),
                ("n_clip", C.c_int32), ("n_run", C.c_int32), ("ring10", C.c_float * 10),
                ("mass_scale", C.c_float), ("thrust_scale", C.c_float), ("cg_offset", C.c_float),
                ("wind", C.c_float * 2), ("delay_ring", (C.c_float * 2) * MAX_DELAY),
                ("clip_bits", C.c_uint32 * 32), ("run_bits", C.c_uint32 * 32)]


class TvcActorWeights(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("w3", C.c_void_p), ("b3", C.c_void_p)]


class TvcForwardModel(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("w3", C.c_void_p), ("b3", C.c_void_p)]


class TvcCuriosityIO(C.Structure):
    _fields_ = [("actions", C.c_void_p), ("obs", C.c_void_p), ("final_obs", C.c_void_p), ("terminated", C.c_void_p),
                ("truncated", C.c_void_p), ("prev_state", C.c_void_p), ("has_prev", C.c_void_p), ("reward_in", C.c_void_p),
                ("reward_out", C.c_void_p), ("intrinsic", C.c_void_p), ("clip_sum", C.c_int32), ("reserved", C.c_int32)]


class TvcReplayRing(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("actions", C.c_void_p), ("reward", C.c_void_p), ("next_obs", C.c_void_p),
                ("terminated", C.c_void_p), ("capacity", C.c_int64)]


class TvcReplayBatch(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("actions", C.c_void_p), ("reward", C.c_void_p), ("next_obs", C.c_void_p),
                ("done", C.c_void_p), ("indices", C.c_void_p)]


class TvcRolloutIO(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("reward_sum", C.c_void_p), ("actions_last", C.c_void_p),
                ("actions_all", C.c_void_p), ("reward_all", C.c_void_p), ("deterministic", C.c_int32),
                ("reserved", C.c_int32), ("obs_all", C.c_void_p), ("next_obs_all", C.c_void_p),
                ("terminated_all", C.c_void_p), ("truncated_all", C.c_void_p)]


_lib = None


def load(path: str | None = None):
    """Load libtvc_b200.so and declare every prototype.  Raises if the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  tvc_ai_b200 has no CPU or PyTorch fallback path.")
    L = C.CDLL(p)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.tvc_abi_version.restype = C.c_int
    L.tvc_last_error.restype = C.c_char_p
    L.tvc_config_default.argtypes = [C.POINTER(TvcConfig), C.c_int]
    L.tvc_create.argtypes = [C.POINTER(TvcConfig), C.c_int, i64, C.POINTER(vp)]
    L.tvc_destroy.argtypes = [vp]
    L.tvc_reset.argtypes = [vp, vp, u64, vp, vp]
    L.tvc_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.tvc_step_ex.argtypes = [vp, C.POINTER(TvcStepIO), vp]
    L.tvc_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.tvc_step_host_async.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.tvc_host_sync.argtypes = [vp]
    L.tvc_rollout.argtypes = [vp, C.POINTER(TvcActorWeights), i32, C.POINTER(TvcRolloutIO), vp]
    L.tvc_state_bytes.argtypes = [vp]
    L.tvc_state_bytes.restype = C.c_size_t
    L.tvc_get_state.argtypes = [vp, vp, C.c_size_t, vp]
    L.tvc_set_state.argtypes = [vp, vp, C.c_size_t, vp]
    L.tvc_get_reward_history.argtypes = [vp, vp, C.c_size_t, vp]
    L.tvc_set_reward_history.argtypes = [vp, vp, C.c_size_t, vp]
    L.tvc_read_info.argtypes = [vp, C.POINTER(TvcInfoSoa), vp]
    L.tvc_episode_stats.argtypes = [vp, C.POINTER(C.c_double), C.c_int, vp]
    L.tvc_episode_stats_dev.argtypes = [vp, vp, C.c_int, vp]
    L.tvc_set_curriculum.argtypes = [vp, C.POINTER(TvcStageConditions)]
    L.tvc_get_config.argtypes = [vp, C.POINTER(TvcConfig)]
    L.tvc_num_envs.argtypes = [vp]
    L.tvc_num_envs.restype = i64
    L.tvc_lifetime_steps.argtypes = [vp]
    L.tvc_lifetime_steps.restype = i64
    L.tvc_curiosity.argtypes = [vp, C.POINTER(TvcForwardModel), C.POINTER(TvcCuriosityIO), vp]
    L.tvc_replay_sample.argtypes = [C.POINTER(TvcReplayRing), i64, i32, u64, u64, vp, C.c_float, C.POINTER(TvcReplayBatch), C.c_int, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("tvc_abi_version",):
            fn.restype = C.c_int
    if L.tvc_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libtvc_b200.so ABI {L.tvc_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    if path is None:
        _lib = L
    return L


def check(rc: int, what: str = "tvc"):
    if rc != 0:
        msg = load().tvc_last_error()
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")


def default_config(contract: int = CONTRACT_R, **over) -> TvcConfig:
    cfg = TvcConfig()
    check(load().tvc_config_default(C.byref(cfg), contract), "tvc_config_default")
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"tvc_config has no field {k!r}")
        setattr(cfg, k, v)
    return cfg
