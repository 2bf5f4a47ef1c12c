"""RocketTVCVectorEnv: Gymnasium-VectorEnv-shaped batched env over the CUDA engine.

The reference has no vectorised env (training.num_envs is never read); this
is the drop-in boundary SURVEY.md section 8(b) specifies: `num_envs`, single/batched spaces,
`reset(seed=, options=) -> (obs[N,10], infos)`, `step(actions[N,2]) -> (obs, rewards, terminations,
truncations, infos)` with Gymnasium 0.26-0.29 same-step autoreset (`final_observation` /
`final_info` plus `_key` masks).  numpy in -> numpy out (through tvc_step_host: host buffers, one C
call); torch CUDA in -> torch CUDA out, zero-copy views of the engine's buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _abi as A
from . import spaces
from .engine import BatchedEngine
from .env import PHASES, engine_config_from_yaml


class RocketTVCVectorEnv:
    metadata = {"render_modes": [], "autoreset_mode": "same_step"}
    OBJECT_INFO_MAX = 4096   # above this, final_observation is a dense array + mask instead of an object array

    def __init__(self, num_envs: int, config: Optional[dict] = None, max_episode_steps: int = 1000,
                 contract: str | int = "R", device: Optional[int] = None, env_id_base: int = 0,
                 final_info: bool = True, copy_outputs: bool = True, enable_curiosity: bool = False,
                 curiosity_module=None, curiosity_impl: str = "tcgen05", **engine_over):
        if isinstance(contract, str):
            contract = {"R": A.CONTRACT_R, "X": A.CONTRACT_X}[contract.upper()]
        self.num_envs = int(num_envs)
        self.contract = contract
        self.max_episode_steps = int(max_episode_steps)
        cfg = engine_config_from_yaml(config, contract, max_episode_steps, autoreset=1, env_id_base=int(env_id_base),
                                      **engine_over)
        self.engine = BatchedEngine(self.num_envs, cfg, device=device)
        self.single_observation_space = spaces.observation_space()
        self.single_action_space = spaces.action_space()
        self.observation_space = spaces.batch_space(self.single_observation_space, self.num_envs)
        self.action_space = spaces.batch_space(self.single_action_space, self.num_envs)
        self.render_mode = None
        self.closed = False
        self._want_final_info = bool(final_info)
        # numpy path: True returns fresh arrays every step (Gymnasium semantics); False returns views of the pinned
        # host buffers tvc_step_host writes into (valid until the next step) and float32 rewards -- no host copies
        self._copy_outputs = bool(copy_outputs)
        self._done_buf = None
        # Row S14 / quirks Q14, Q19 for the batch (torch path): the reference's never-trained forward model adds
        # 0.01 * mean((f([s8, a]) - s8')^2) to the clipped reward, skipped on the first step of every episode
        # (ref:257-269, :496-502).  Off by default, as in the reference's evaluation env (scripts/train.py:329).
        # curiosity_impl: "tcgen05" = the batched tensor-core kernel tvc_curiosity (bf16 operands, fp32 accumulation: the
        # term agrees with the fp32 module to ~3e-3 relative, the reward to < 1e-6); "torch" = the fp32 torch module.
        self.enable_curiosity = bool(enable_curiosity)
        self.curiosity_module = None
        if curiosity_impl not in ("tcgen05", "torch"):
            raise ValueError("curiosity_impl must be 'tcgen05' or 'torch'")
        self.curiosity_impl = curiosity_impl
        if self.enable_curiosity:
            from .env import CuriosityModule
            self.curiosity_module = curiosity_module or CuriosityModule(obs_dim=8, action_dim=2, device=self.engine.device)
            self._prev_s8 = torch.zeros((self.num_envs, 8), device=self.engine.device)
            self._has_prev = torch.zeros(self.num_envs, dtype=torch.bool, device=self.engine.device)
            self._has_prev_u8 = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.engine.device)
            self._rew_total = torch.zeros(self.num_envs, dtype=torch.float32, device=self.engine.device)
            self.intrinsic = torch.zeros(self.num_envs, dtype=torch.float32, device=self.engine.device)
            self._fm_packed = False

    # ------------------------------------------------------------------
    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        """Resets every env.  `seed` re-keys the Philox streams in Contract X; in Contract R the
        reference's reset is deterministic and ignores it (quirk Q15)."""
        as_torch = bool(options and options.get("return_torch"))
        obs = self.engine.reset(seed=seed)
        if self.enable_curiosity:
            self._has_prev.zero_()          # ref:399-401: reset() clears state_history
            self._has_prev_u8.zero_()
        return (obs if as_torch else obs.cpu().numpy().copy()), {}

    def step(self, actions):
        if isinstance(actions, torch.Tensor):
            return self._step_torch(actions)
        return self._step_numpy(np.asarray(actions, np.float32))

    def step_random(self):
        """Step with in-kernel Philox U(-1,1) actions (synthetic-workload path used by bench.py)."""
        return self._step_torch(None)

    def _step_torch(self, actions):
        if actions is not None and actions.dtype != torch.float32:
            actions = actions.float()
        if self._want_final_info:
            obs, rew, term, trunc, info = self.engine.step_ex(actions)
        else:
            obs, rew, term, trunc = self.engine.step(actions)
            info = None
        term_b, trunc_b = term.bool(), trunc.bool()
        done = term_b | trunc_b
        if self.enable_curiosity:
            if actions is None:
                raise ValueError("enable_curiosity needs the actions (in-kernel random actions are not visible to the forward model)")
            if self.curiosity_impl == "tcgen05":
                # one launch: forward model on the tensor cores + MSE + reward add + history update (csrc/tvc_curiosity.cu)
                rew = self.engine.curiosity(actions.reshape(self.num_envs, 2), self._prev_s8, self._has_prev_u8, self._rew_total,
                                            intrinsic=self.intrinsic,
                                            forward_model=None if self._fm_packed else self.curiosity_module.forward_model)
                self._fm_packed = True
            else:
                with torch.no_grad():
                    a = actions.reshape(self.num_envs, 2).clamp(-1.0, 1.0)
                    s8_next = torch.where(done[:, None], self.engine.final_obs[:, :8], obs[:, :8])   # the env's own next state
                    pred = self.curiosity_module.forward_model(torch.cat([self._prev_s8, a], dim=1))
                    intrinsic = 0.01 * ((pred - s8_next) ** 2).mean(dim=1)
                    self.intrinsic = torch.where(self._has_prev, intrinsic, torch.zeros_like(intrinsic))
                    rew = rew + self.intrinsic   # a new tensor: the engine's buffer keeps the extrinsic reward
                    self._prev_s8.copy_(obs[:, :8])
                    self._has_prev.copy_(~done)
        infos = {"final_observation": self.engine.final_obs, "_final_observation": done}
        if info is not None:
            infos["final_info"] = {k: info[k] for k in ("altitude", "tilt_deg", "omega_mag", "fuel", "phase", "step",
                                                        "success", "criteria_met", "reward_components")}
            infos["_final_info"] = done
        return obs, rew, term_b, trunc_b, infos

    def pinned_actions(self) -> np.ndarray:
        """Pinned [N,2] float32 buffer: a policy that writes its actions here and passes it to `step` avoids the staging
        copy a pageable array needs (the copy engine reads pinned host memory directly)."""
        return self.engine.pinned_actions()

    def _step_numpy(self, actions):
        if actions.shape != (self.num_envs, 2):
            raise ValueError(f"actions must have shape {(self.num_envs, 2)}, got {actions.shape}")
        obs, rew, term, trunc, final = self.engine.step_host(actions, want_final=True)
        if self._copy_outputs:
            obs, rew, term, trunc = obs.copy(), rew.astype(np.float64), term.copy(), trunc.copy()
        if self._copy_outputs:
            done = term | trunc
        else:       # views path: the mask lives in a reused buffer as well
            if self._done_buf is None:
                self._done_buf = np.empty(self.num_envs, np.bool_)
            done = np.bitwise_or(term, trunc, out=self._done_buf)
        infos = {}
        if done.any():
            if self.num_envs <= self.OBJECT_INFO_MAX:
                # Gymnasium 0.26-0.29 layout: object array holding the terminal observation of finished envs
                fo = np.full(self.num_envs, None, dtype=object)
                for i in np.flatnonzero(done):
                    fo[i] = final[i].copy()
            else:
                # large batches: dense [N,10] array, valid where the mask is set (no per-env Python loop)
                fo = final.copy() if self._copy_outputs else final
            infos["final_observation"] = fo
            infos["_final_observation"] = done
        return obs, rew, term, trunc, infos

    # ------------------------------------------------------------------
    def episode_stats(self, reset_after: bool = False) -> dict:
        s = self.engine.stats(reset_after)
        return dict(zip(A.STAT_NAMES, s.tolist()))

    def set_curriculum(self, conditions: dict):
        self.engine.set_curriculum(conditions)

    def call_info(self) -> dict:
        """Current (post-autoreset) per-env info as numpy arrays, keys as in the reference's info dict."""
        t = self.engine.read_info()
        phase = t["phase"].cpu().numpy()
        return {"altitude": t["altitude"].cpu().numpy(), "tilt_angle_deg": t["tilt_deg"].cpu().numpy(),
                "angular_velocity_mag": t["omega_mag"].cpu().numpy(), "fuel_remaining": t["fuel"].cpu().numpy(),
                "mission_phase": np.array([PHASES[p].value for p in phase]),
                "mission_successful": t["success"].cpu().numpy().astype(bool), "step": t["step"].cpu().numpy(),
                "success_criteria_met": t["criteria_met"].cpu().numpy().astype(bool),
                "position": t["position"].cpu().numpy()}

    def close(self, **kwargs):
        if not self.closed:
            self.engine.close()
            self.closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class RocketTVCHostPipelineEnv:
    """The numpy (host-buffer) side of `RocketTVCVectorEnv` for large batches, pipelined over env slabs.

    The envs are split into `slabs` contiguous slabs, each with its own engine (handle, state, stream; `env_id_base`
    keeps the global env ids, so the trajectories are those of one big engine, bit for bit).  `step()` enqueues every
    slab's host step (`tvc_step_host_async`: H2D of its actions, kernels, D2H of its results) and then waits for all of
    them: the copy engines move slab k's results to the host while the SMs step slab k+1.  Results land in pinned arrays
    owned by this object and are returned as views (valid until the next step), Gymnasium `VectorEnv.step` shapes.
    """

    def __init__(self, num_envs: int, config: Optional[dict] = None, max_episode_steps: int = 1000,
                 contract: str | int = "X", device: Optional[int] = None, env_id_base: int = 0, slabs: int = 2,
                 **engine_over):
        if isinstance(contract, str):
            contract = {"R": A.CONTRACT_R, "X": A.CONTRACT_X}[contract.upper()]
        self.num_envs = int(num_envs)
        slabs = max(1, min(int(slabs), self.num_envs))
        edges = [round(k * self.num_envs / slabs) for k in range(slabs + 1)]
        self._ranges = [(edges[k], edges[k + 1]) for k in range(slabs) if edges[k + 1] > edges[k]]
        self.engines = [BatchedEngine(hi - lo, engine_config_from_yaml(config, contract, max_episode_steps, autoreset=1,
                                                                       env_id_base=int(env_id_base) + lo, **engine_over),
                                      device=device) for lo, hi in self._ranges]
        self.single_observation_space = spaces.observation_space()
        self.single_action_space = spaces.action_space()
        self.observation_space = spaces.batch_space(self.single_observation_space, self.num_envs)
        self.action_space = spaces.batch_space(self.single_action_space, self.num_envs)
        n = self.num_envs
        pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()  # noqa: E731
        self._t = dict(act=pin((n, 2), torch.float32), obs=pin((n, 10), torch.float32), rew=pin((n,), torch.float32),
                       term=pin((n,), torch.uint8), trunc=pin((n,), torch.uint8), final=pin((n, 10), torch.float32))
        self._np = {k: v.numpy() for k, v in self._t.items()}
        self.closed = False

    def pinned_actions(self) -> np.ndarray:
        """Pinned [N,2] float32 buffer; pass it to `step` after writing the actions into it (no staging copy)."""
        return self._np["act"]

    def reset(self, *, seed: Optional[int] = None, options: Optional[dict] = None):
        obs = self._np["obs"]
        for eng, (lo, hi) in zip(self.engines, self._ranges):
            obs[lo:hi] = eng.reset(seed=seed).cpu().numpy()
        return obs.copy(), {}

    def step(self, actions):
        a = np.ascontiguousarray(actions, np.float32)
        if a.shape != (self.num_envs, 2):
            raise ValueError(f"actions must have shape {(self.num_envs, 2)}, got {a.shape}")
        b = self._np
        vp = lambda arr, lo: C.c_void_p(arr.ctypes.data + lo * arr.strides[0])  # noqa: E731
        for eng, (lo, hi) in zip(self.engines, self._ranges):
            A.check(eng.L.tvc_step_host_async(eng.h, vp(a, lo), vp(b["obs"], lo), vp(b["rew"], lo), vp(b["term"], lo),
                                              vp(b["trunc"], lo), vp(b["final"], lo)), "tvc_step_host_async")
        for eng in self.engines:
            A.check(eng.L.tvc_host_sync(eng.h), "tvc_host_sync")
        term, trunc = b["term"].view(np.bool_), b["trunc"].view(np.bool_)
        done = term | trunc
        infos = {"final_observation": b["final"], "_final_observation": done} if done.any() else {}
        return b["obs"], b["rew"], term, trunc, infos

    def episode_stats(self, reset_after: bool = False) -> dict:
        s = sum(eng.stats(reset_after) for eng in self.engines)
        return dict(zip(A.STAT_NAMES, s.tolist()))

    def set_curriculum(self, conditions: dict):
        for eng in self.engines:
            eng.set_curriculum(conditions)

    def close(self, **kwargs):
        if not self.closed:
            for eng in self.engines:
                eng.close()
            self.closed = True
