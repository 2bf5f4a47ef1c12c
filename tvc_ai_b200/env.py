"""Drop-in facades with the reference's names and signatures.

`EnhancedRocketTVCEnv` mirrors /root/reference/env/enhanced_rocket_tvc_env.py:271-753 (same ctor
kwargs, attributes, reset/step/close signatures, info keys and Python types), as an N=1 view of
the batched CUDA engine.  The factories mirror /root/reference/env/__init__.py:28-111.
`install_as_reference_env()` registers this module under the import paths the reference trainer
uses (scripts/train.py:44), so scripts/train.py runs unmodified on the B200 engine.
"""
from __future__ import annotations

import logging
import sys
import types
from collections import deque
from dataclasses import dataclass
from enum import Enum
from typing import Optional

import numpy as np
import torch

from . import _abi as A
from . import spaces
from .engine import BatchedEngine


class MissionPhase(Enum):
    """Same members, values and order as enhanced_rocket_tvc_env.py:21-29."""
    BOOST = "boost"
    COAST = "coast"
    LANDING = "landing"
    TOUCHDOWN = "touchdown"
    HOVER = "hover"
    COMPLETE = "complete"
    FAILED = "failed"


class SuccessCriteria(Enum):
    ATTITUDE = "attitude"
    VELOCITY = "velocity"
    POSITION = "position"
    STABILITY = "stability"
    FUEL = "fuel"


PHASES = tuple(MissionPhase)


@dataclass
class MissionSuccess:
    """Thresholds of enhanced_rocket_tvc_env.py:39-61 (informational: the kernel hard-codes them)."""
    max_tilt_angle: float = 0.087
    max_angular_velocity: float = 0.1
    max_horizontal_velocity: float = 0.5
    max_vertical_velocity: float = 2.0
    min_altitude: float = 0.2
    max_altitude: float = 2.0
    position_tolerance: float = 1.0
    success_duration: int = 100


def engine_config_from_yaml(config: Optional[dict], contract: int, max_episode_steps: int, **over) -> A.TvcConfig:
    """Map the reference's YAML dict onto tvc_config.

    The reference env reads exactly one key, config['reward_function'] and from it only
    gradient_penalty / diversity_bonus (enhanced_rocket_tvc_env.py:301, :83-84).  Contract X also
    honours env.domain_randomization.parameters.* (config/config.yaml:340-349)."""
    config = config or {}
    cfg = A.default_config(contract)
    cfg.max_episode_steps = int(max_episode_steps)
    rf = config.get("reward_function", {}) or {}
    cfg.gradient_penalty = float(rf.get("gradient_penalty", 0.1))
    cfg.diversity_bonus = float(rf.get("diversity_bonus", 0.05))
    if contract == A.CONTRACT_X:
        envc = config.get("env", {}) or {}
        dr = envc.get("domain_randomization", {}) or {}
        if dr and not dr.get("enabled", True):
            cfg.mass_variation = cfg.thrust_std = cfg.cg_offset_max = cfg.wind_std = cfg.sensor_noise_std = 0.0
        par = dr.get("parameters", {}) or {}
        if "mass" in par:
            cfg.mass_variation = float(par["mass"].get("variation", cfg.mass_variation))
        if "thrust" in par:
            cfg.thrust_std = float(par["thrust"].get("variation", cfg.thrust_std))
        if "cg_offset" in par:
            cfg.cg_offset_max = float(par["cg_offset"].get("max", cfg.cg_offset_max))
        if "wind" in par:
            cfg.wind_std = float(par["wind"].get("max_force", cfg.wind_std))
        if "sensor_noise" in par:
            cfg.sensor_noise_std = float(par["sensor_noise"].get("std", cfg.sensor_noise_std))
        seed = (config.get("globals", {}) or {}).get("seed")
        if seed is not None:
            cfg.seed = int(seed)
    for k, v in over.items():
        if not hasattr(cfg, k):
            raise AttributeError(f"tvc_config has no field {k!r}")
        setattr(cfg, k, v)
    return cfg


class CuriosityModule:
    """enhanced_rocket_tvc_env.py:226-269: untrained forward model, 0.01 * MSE (quirk Q19).

    Constructed in the reference's order (inverse model first) so the same torch seed yields the
    same random weights.  Runs in torch on the engine's device; it is outside the hot path."""

    def __init__(self, obs_dim: int, action_dim: int, hidden_dim: int = 256, device="cpu"):
        nn = torch.nn
        self.inverse_model = nn.Sequential(nn.Linear(obs_dim * 2, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim),
                                           nn.ReLU(), nn.Linear(hidden_dim, action_dim))
        self.forward_model = nn.Sequential(nn.Linear(obs_dim + action_dim, hidden_dim), nn.ReLU(),
                                           nn.Linear(hidden_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, obs_dim))
        self.forward_model.to(device)
        self.device = device

    @torch.no_grad()
    def compute_intrinsic_reward(self, state, action, next_state) -> float:
        s = torch.as_tensor(np.asarray(state, np.float32), device=self.device).unsqueeze(0)
        a = torch.as_tensor(np.asarray(action, np.float32), device=self.device).unsqueeze(0)
        n = torch.as_tensor(np.asarray(next_state, np.float32), device=self.device).unsqueeze(0)
        pred = self.forward_model(torch.cat([s, a], dim=1))
        return torch.nn.functional.mse_loss(pred, n).item() * 0.01


class EnhancedRocketTVCEnv(spaces.EnvBase):
    """Single-env, reference-faithful (Contract R) view of the batched CUDA engine."""

    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 60}

    def __init__(self, config: Optional[dict] = None, max_episode_steps: int = 1000, render_mode: Optional[str] = None,
                 enable_hierarchical: bool = True, enable_curiosity: bool = True, enable_physics_informed: bool = True,
                 debug: bool = False, *, device: Optional[int] = None, contract: int = A.CONTRACT_R, **engine_over):
        if spaces.HAVE_GYMNASIUM:
            super().__init__()
        self.config = config or {}
        self.max_episode_steps = max_episode_steps
        self.render_mode = render_mode
        self.enable_hierarchical = enable_hierarchical
        self.enable_curiosity = enable_curiosity
        self.enable_physics_informed = enable_physics_informed
        self.debug = debug
        self.mission_success = MissionSuccess()
        cfg = engine_config_from_yaml(self.config, contract, max_episode_steps, autoreset=0, **engine_over)
        self._engine = BatchedEngine(1, cfg, device=device)
        if enable_curiosity:
            self.curiosity_module = CuriosityModule(obs_dim=8, action_dim=2, device=self._engine.device)
        self.current_phase = MissionPhase.BOOST
        self.mission_successful = False
        self.phase_start_time = 0
        self.current_step = 0
        self.fuel_remaining = 1.0
        self.state_history = deque(maxlen=100)
        self.action_history = deque(maxlen=100)
        self.reward_components_history = deque(maxlen=100)
        self.observation_space = spaces.observation_space()
        self.action_space = spaces.action_space()
        self.np_random = np.random.default_rng()
        self._act = torch.zeros((1, 2), dtype=torch.float32, device=self._engine.device)
        logging.basicConfig(level=logging.DEBUG if debug else logging.INFO)
        self.logger = logging.getLogger(__name__)

    # ---- enhanced_rocket_tvc_env.py:381-407 ----
    def reset(self, seed: Optional[int] = None, options: Optional[dict] = None):
        if spaces.HAVE_GYMNASIUM:
            super().reset(seed=seed)
        elif seed is not None:
            self.np_random = np.random.default_rng(seed)   # seeded and never used (quirk Q15)
        obs = self._engine.reset().cpu().numpy()[0].copy()
        self.current_phase = MissionPhase.BOOST
        self.mission_successful = False
        self.phase_start_time = 0
        self.current_step = 0
        self.fuel_remaining = 1.0
        self.state_history.clear()
        self.action_history.clear()
        self.reward_components_history.clear()
        return obs, self._info_dict(self._engine.read_info())

    def _info_dict(self, t) -> dict:
        pos = t["position"][0].tolist()
        phase = PHASES[int(t["phase"][0])]
        return {
            "position": tuple(float(x) for x in pos),
            "altitude": float(t["altitude"][0]),
            "tilt_angle_deg": np.float64(t["tilt_deg"][0].item()),
            "angular_velocity_mag": np.float64(t["omega_mag"][0].item()),
            "fuel_remaining": float(t["fuel"][0]),
            "mission_phase": phase.value,
            "mission_successful": bool(t["success"][0]),
            "step": int(t["step"][0]),
            "success_criteria_met": bool(t["criteria_met"][0]),
        }

    # ---- enhanced_rocket_tvc_env.py:466-518 ----
    def step(self, action):
        action = np.clip(np.asarray(action, np.float32).reshape(2), -1.0, 1.0)
        self._act.copy_(torch.from_numpy(action).view(1, 2))
        obs_t, rew_t, term_t, trunc_t, info_t = self._engine.step_ex(self._act)
        obs = obs_t.cpu().numpy()[0].copy()
        reward = np.float64(rew_t.item())
        comp = info_t["reward_components"][0].tolist()
        info = self._info_dict(info_t)
        reward_components = {n: comp[i] for i, n in enumerate(A.COMPONENT_NAMES[:6])}
        for i in (6, 7, 8):   # penalties exist only when they fire (ref:189-207)
            if comp[i] != 0.0:
                reward_components[A.COMPONENT_NAMES[i]] = comp[i]
        if self.enable_curiosity and len(self.state_history) > 0:   # ref:496-502, quirks Q14, Q19
            intrinsic = self.curiosity_module.compute_intrinsic_reward(self.state_history[-1], action, obs[:8])
            reward = reward + intrinsic
            reward_components["curiosity"] = intrinsic
        self.state_history.append(obs[:8].copy())
        self.action_history.append(action.copy())
        self.reward_components_history.append(reward_components.copy())
        self.current_step = info["step"]
        self.current_phase = MissionPhase(info["mission_phase"])
        self.mission_successful = info["mission_successful"]
        self.fuel_remaining = info["fuel_remaining"]
        info["reward_components"] = reward_components
        return obs, reward, bool(term_t.item()), bool(trunc_t.item()), info

    def render(self, mode: str = "human"):
        return None

    def close(self):
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng.close()
            self._engine = None


def make_enhanced_tvc_env(**kwargs) -> EnhancedRocketTVCEnv:
    return EnhancedRocketTVCEnv(**kwargs)


# ---- /root/reference/env/__init__.py:66-102 ----
def make_training_env(config=None, **kwargs):
    kw = dict(max_episode_steps=1000, enable_hierarchical=True, enable_curiosity=True, enable_physics_informed=True, debug=False)
    kw.update(kwargs)
    return EnhancedRocketTVCEnv(config=config, **kw)


def make_evaluation_env(config=None, **kwargs):
    kw = dict(max_episode_steps=1000, enable_hierarchical=False, enable_curiosity=False, enable_physics_informed=False, debug=False)
    kw.update(kwargs)
    return EnhancedRocketTVCEnv(config=config, **kw)


def make_debug_env(config=None, **kwargs):
    kw = dict(render_mode="human", max_episode_steps=1000, enable_hierarchical=True, enable_curiosity=True,
              enable_physics_informed=True, debug=True)
    kw.update(kwargs)
    return EnhancedRocketTVCEnv(config=config, **kw)


REGISTERED_IDS = {
    "EnhancedRocketTVC-v0": dict(enable_hierarchical=True, enable_curiosity=True, enable_physics_informed=True, debug=False),
    "EnhancedRocketTVC-Eval-v0": dict(enable_hierarchical=False, enable_curiosity=False, enable_physics_informed=False, debug=False),
    "EnhancedRocketTVC-Debug-v0": dict(enable_hierarchical=True, enable_curiosity=True, enable_physics_informed=True, debug=True),
}


def register_gym_ids():
    """env/__init__.py:28-64: the three ids, each with max_episode_steps=1000 (quirk Q22)."""
    if not spaces.HAVE_GYMNASIUM:
        return False
    from gymnasium.envs.registration import register, registry  # pragma: no cover
    for env_id, kw in REGISTERED_IDS.items():  # pragma: no cover
        if env_id not in registry:
            register(id=env_id, entry_point="tvc_ai_b200.env:EnhancedRocketTVCEnv", max_episode_steps=1000, kwargs=kw)
    return True  # pragma: no cover


def install_as_reference_env():
    """Make `from env.enhanced_rocket_tvc_env import EnhancedRocketTVCEnv, MissionPhase` and
    `from env import make_training_env, ...` resolve to this module (scripts/train.py:44)."""
    me = sys.modules[__name__]
    pkg = types.ModuleType("env")
    pkg.__path__ = []
    for name in ("EnhancedRocketTVCEnv", "MissionPhase", "make_training_env", "make_evaluation_env", "make_debug_env"):
        setattr(pkg, name, getattr(me, name))
    pkg.enhanced_rocket_tvc_env = me
    sys.modules["env"] = pkg
    sys.modules["env.enhanced_rocket_tvc_env"] = me
    return pkg
