"""BatchedEngine: one libtvc_b200 handle + the torch CUDA buffers it writes into.

PyTorch is plumbing here (device memory, streams, torch.distributed); every number is produced by
the sm_100a kernels behind the C ABI.  There is no CPU path: constructing an engine without a
B200-class device raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi as A

STATE_DTYPE = np.dtype([
    ("pos", np.float32, 3), ("quat", np.float32, 4), ("vel", np.float32, 3), ("omega", np.float32, 3),
    ("prev_action", np.float32, 2), ("ep_return", np.float32),
    ("step", np.int32), ("burn", np.int32), ("phase", np.int32), ("success", np.int32), ("has_prev", np.int32),
    ("consec", np.int32), ("hist_count", np.int32), ("episode", np.int32), ("n_clip", np.int32), ("n_run", np.int32),
    ("ring10", np.float32, 10), ("mass_scale", np.float32), ("thrust_scale", np.float32), ("cg_offset", np.float32),
    ("wind", np.float32, 2), ("delay_ring", np.float32, (A.MAX_DELAY, 2)),
    ("clip_bits", np.uint32, 32), ("run_bits", np.uint32, 32)])
assert STATE_DTYPE.itemsize == C.sizeof(A.TvcEnvState), (STATE_DTYPE.itemsize, C.sizeof(A.TvcEnvState))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchedEngine:
    def __init__(self, num_envs: int, config: A.TvcConfig | None = None, device: int | None = None,
                 contract: int = A.CONTRACT_R, **over):
        self.L = A.load()
        if not torch.cuda.is_available():
            raise RuntimeError("tvc_ai_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        cfg = config if config is not None else A.default_config(contract)
        for k, v in over.items():
            if not hasattr(cfg, k):
                raise AttributeError(f"tvc_config has no field {k!r}")
            setattr(cfg, k, v)
        self.n = int(num_envs)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            A.check(self.L.tvc_create(C.byref(cfg), self.device_index, self.n, C.byref(h)), "tvc_create")
        self.h = h
        n, dev = self.n, self.device
        self.obs = torch.zeros((n, A.OBS_DIM), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.terminated = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.final_obs = torch.zeros((n, A.OBS_DIM), dtype=torch.float32, device=dev)
        self._info = None
        self._host = None
        self._pinned_act = None
        self._stats_dev = torch.zeros(A.NUM_STATS, dtype=torch.float64, device=dev)

    # ------------------------------------------------------------------ lifecycle
    def close(self):
        if getattr(self, "h", None):
            self.L.tvc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def config(self) -> A.TvcConfig:
        cfg = A.TvcConfig()
        A.check(self.L.tvc_get_config(self.h, C.byref(cfg)), "tvc_get_config")
        return cfg

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_actions(self, actions):
        if actions is None:
            return None
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.float32
                and actions.device == self.device):
            raise TypeError("actions must be a float32 CUDA tensor on the engine's device")
        if tuple(actions.shape) != (self.n, A.ACT_DIM):
            raise ValueError(f"actions must have shape {(self.n, A.ACT_DIM)}, got {tuple(actions.shape)}")
        return actions if actions.is_contiguous() else actions.contiguous()

    # ------------------------------------------------------------------ hot path
    def reset(self, mask: torch.Tensor | None = None, seed: int | None = None):
        """seed=None keeps the current Philox key; any integer (0 included) re-keys the streams."""
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        key = A.SEED_KEEP if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        A.check(self.L.tvc_reset(self.h, _ptr(mask), key, _ptr(self.obs), self._stream()), "tvc_reset")
        return self.obs

    def step(self, actions: torch.Tensor | None, want_final: bool = True):
        """One env step for all envs.  actions=None draws Philox U(-1,1) actions in-kernel."""
        a = self._check_actions(actions)
        if a is not None:
            A.check(self.L.tvc_step(self.h, _ptr(a), _ptr(self.obs), _ptr(self.reward), _ptr(self.terminated),
                                    _ptr(self.truncated), _ptr(self.final_obs) if want_final else None,
                                    self._stream()), "tvc_step")
        else:
            io = A.TvcStepIO()
            io.obs, io.reward = self.obs.data_ptr(), self.reward.data_ptr()
            io.terminated, io.truncated = self.terminated.data_ptr(), self.truncated.data_ptr()
            io.final_obs = self.final_obs.data_ptr() if want_final else None
            A.check(self.L.tvc_step_ex(self.h, C.byref(io), self._stream()), "tvc_step_ex")
        return self.obs, self.reward, self.terminated, self.truncated

    def _info_buffers(self):
        if self._info is None:
            n, dev = self.n, self.device
            f = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
            self._info = dict(altitude=f(n), tilt_deg=f(n), omega_mag=f(n), fuel=f(n), position=f(n, 3),
                              phase=torch.zeros(n, dtype=torch.int32, device=dev),
                              step=torch.zeros(n, dtype=torch.int32, device=dev),
                              success=torch.zeros(n, dtype=torch.uint8, device=dev),
                              criteria_met=torch.zeros(n, dtype=torch.uint8, device=dev),
                              reward_components=f(n, A.NUM_COMPONENTS),
                              actions=f(n, A.ACT_DIM))
        return self._info

    def _info_struct(self, with_components=True):
        b = self._info_buffers()
        s = A.TvcInfoSoa()
        for k in ("altitude", "tilt_deg", "omega_mag", "fuel", "position", "phase", "step", "success", "criteria_met"):
            setattr(s, k, b[k].data_ptr())
        s.reward_components = b["reward_components"].data_ptr() if with_components else None
        return s

    def step_ex(self, actions: torch.Tensor | None):
        """Step and also return the terminal (pre-autoreset) info tensors and reward components."""
        a = self._check_actions(actions)
        io = A.TvcStepIO()
        io.actions = a.data_ptr() if a is not None else None
        io.obs, io.reward = self.obs.data_ptr(), self.reward.data_ptr()
        io.terminated, io.truncated = self.terminated.data_ptr(), self.truncated.data_ptr()
        io.final_obs = self.final_obs.data_ptr()
        io.info = self._info_struct()
        io.actions_out = self._info["actions"].data_ptr()
        A.check(self.L.tvc_step_ex(self.h, C.byref(io), self._stream()), "tvc_step_ex")
        return self.obs, self.reward, self.terminated, self.truncated, self._info

    def step_host(self, actions_np: np.ndarray | None, want_final: bool = False):
        """End-to-end step through HOST buffers (H2D + kernel + D2H + sync inside the C call)."""
        if self._host is None:
            n = self.n
            pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()  # noqa: E731
            # obs | reward | terminated | truncated share one pinned slab (46 B per env, the library's own staging
            # layout): tvc_step_host then needs a single device-to-host copy
            slab = pin((46 * n,), torch.uint8)
            self._host = dict(slab=slab,
                              obs=slab[:40 * n].view(torch.float32).view(n, 10),
                              rew=slab[40 * n:44 * n].view(torch.float32),
                              term=slab[44 * n:45 * n], trunc=slab[45 * n:46 * n],
                              final=pin((n, 10), torch.float32))
            h_ = self._host
            # numpy views and raw pointers are made once: the step path does no per-call tensor bookkeeping
            h_["views"] = (h_["obs"].numpy(), h_["rew"].numpy(), h_["term"].numpy().view(np.bool_), h_["trunc"].numpy().view(np.bool_),
                           h_["final"].numpy())
            h_["ptrs"] = tuple(C.c_void_p(h_[k].data_ptr()) for k in ("obs", "rew", "term", "trunc", "final"))
        hb = self._host
        ap = None
        if actions_np is not None:
            # the library reads pinned memory (e.g. `pinned_actions()`) in place and stages pageable arrays itself
            act = actions_np if (actions_np.dtype == np.float32 and actions_np.flags.c_contiguous) else np.ascontiguousarray(actions_np, np.float32)
            if act.size != 2 * self.n:
                raise ValueError(f"actions must have shape ({self.n}, 2)")
            ap = act.ctypes.data
        p = hb["ptrs"]
        A.check(self.L.tvc_step_host(self.h, ap, p[0], p[1], p[2], p[3], p[4] if want_final else None), "tvc_step_host")
        v = hb["views"]
        return v[0], v[1], v[2], v[3], (v[4] if want_final else None)

    def pinned_actions(self) -> np.ndarray:
        """A pinned (page-locked) [N,2] float32 array: actions written here go to the device without a staging copy."""
        if self._pinned_act is None:
            self._pinned_act = torch.zeros((self.n, 2), dtype=torch.float32).pin_memory()
        return self._pinned_act.numpy()

    # ------------------------------------------------------------------ state / info / stats
    def get_state(self) -> np.ndarray:
        blob = torch.empty(self.n * STATE_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        A.check(self.L.tvc_get_state(self.h, _ptr(blob), blob.numel(), self._stream()), "tvc_get_state")
        return blob.cpu().numpy().view(STATE_DTYPE).copy()

    def set_state(self, state: np.ndarray):
        state = np.ascontiguousarray(state, dtype=STATE_DTYPE)
        if state.shape != (self.n,):
            raise ValueError(f"state must have shape ({self.n},)")
        blob = torch.from_numpy(state.view(np.uint8).copy()).to(self.device)
        A.check(self.L.tvc_set_state(self.h, _ptr(blob), blob.numel(), self._stream()), "tvc_set_state")
        torch.cuda.current_stream(self.device).synchronize()

    def get_reward_history(self) -> torch.Tensor:
        """TVC_DIV_EXACT only: the [N, 1000] reward window (checkpoints; the rest of the state is in get_state())."""
        out = torch.empty((self.n, 1000), dtype=torch.float32, device=self.device)
        A.check(self.L.tvc_get_reward_history(self.h, _ptr(out), out.numel() * 4, self._stream()), "tvc_get_reward_history")
        return out

    def set_reward_history(self, hist: torch.Tensor):
        hist = hist.to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(hist.shape) != (self.n, 1000):
            raise ValueError(f"history must have shape ({self.n}, 1000)")
        A.check(self.L.tvc_set_reward_history(self.h, _ptr(hist), hist.numel() * 4, self._stream()), "tvc_set_reward_history")
        torch.cuda.current_stream(self.device).synchronize()

    def read_info(self):
        s = self._info_struct(with_components=False)
        A.check(self.L.tvc_read_info(self.h, C.byref(s), self._stream()), "tvc_read_info")
        return self._info

    def stats(self, reset_after: bool = False) -> np.ndarray:
        out = (C.c_double * A.NUM_STATS)()
        A.check(self.L.tvc_episode_stats(self.h, out, int(reset_after), self._stream()), "tvc_episode_stats")
        return np.array(out)

    def stats_device(self, reset_after: bool = False) -> torch.Tensor:
        """Per-GPU statistics vector left on the device (input of the NCCL all-reduce)."""
        A.check(self.L.tvc_episode_stats_dev(self.h, _ptr(self._stats_dev), int(reset_after), self._stream()),
                "tvc_episode_stats_dev")
        return self._stats_dev

    def set_curriculum(self, conditions: dict):
        c = A.TvcStageConditions()
        c.max_initial_tilt = float(conditions.get("max_initial_tilt", 0.0))
        c.max_initial_angular_vel = float(conditions.get("max_initial_angular_vel", 0.0))
        c.domain_randomization = int(bool(conditions.get("domain_randomization", False)))
        c.sensor_noise = int(bool(conditions.get("sensor_noise", False)))
        c.max_gimbal_angle_deg = float(conditions.get("max_gimbal_angle", 0.0) or 0.0)
        c.wind_enabled = int(bool(conditions.get("wind_enabled", False)))
        c.wind_force = float(conditions.get("wind_force", 0.0))
        c.mass_variation = float(conditions.get("mass_variation", 0.0))
        A.check(self.L.tvc_set_curriculum(self.h, C.byref(c)), "tvc_set_curriculum")

    @property
    def lifetime_steps(self) -> int:
        return int(self.L.tvc_lifetime_steps(self.h))

    # ------------------------------------------------------------------ fused rollout
    def curiosity(self, actions: torch.Tensor, prev_state: torch.Tensor, has_prev: torch.Tensor, reward_out: torch.Tensor,
                  intrinsic: torch.Tensor | None = None, forward_model=None, clip_sum: bool = False):
        """Row S14 for the batch on the tensor cores (tvc_curiosity): call after `step` with that step's actions.

        prev_state [N,8] f32 and has_prev [N] u8 are the kernel's in/out history; reward_out [N] receives reward + intrinsic
        (the engine's own reward buffer keeps the extrinsic value).  forward_model: a torch nn.Sequential
        Linear(10,256)-ReLU-Linear(256,256)-ReLU-Linear(256,8) to (re)pack, or None to reuse the packed weights."""
        w = None
        keep = []
        if forward_model is not None:
            w = A.TvcForwardModel()
            lin = [m for m in forward_model if isinstance(m, torch.nn.Linear)]
            if [tuple(m.weight.shape) for m in lin] != [(256, 10), (256, 256), (8, 256)]:
                raise ValueError("tvc_curiosity is built for the reference's forward model 10-256-256-8")
            for k, m in zip(("1", "2", "3"), lin):
                wt = m.weight.detach().to(device=self.device, dtype=torch.float32).contiguous()
                bt = m.bias.detach().to(device=self.device, dtype=torch.float32).contiguous()
                keep += [wt, bt]
                setattr(w, "w" + k, wt.data_ptr()), setattr(w, "b" + k, bt.data_ptr())
        a = actions.to(device=self.device, dtype=torch.float32).contiguous()
        io = A.TvcCuriosityIO()
        io.actions, io.obs, io.final_obs = a.data_ptr(), self.obs.data_ptr(), self.final_obs.data_ptr()
        io.terminated, io.truncated = self.terminated.data_ptr(), self.truncated.data_ptr()
        io.prev_state, io.has_prev = prev_state.data_ptr(), has_prev.data_ptr()
        io.reward_in, io.reward_out = self.reward.data_ptr(), reward_out.data_ptr()
        io.intrinsic = intrinsic.data_ptr() if intrinsic is not None else None
        io.clip_sum = int(clip_sum)
        A.check(self.L.tvc_curiosity(self.h, C.byref(w) if w is not None else None, C.byref(io), self._stream()), "tvc_curiosity")
        if keep:
            torch.cuda.current_stream(self.device).synchronize()   # the pack kernel read the temporaries
        return reward_out

    def rollout(self, weights: dict, T: int, deterministic: bool = False, record: bool = False, transitions: dict | None = None):
        """T env steps per launch with the 2x256 SAC actor evaluated in-kernel (tvc_rollout).

        transitions: optional dict of preallocated CUDA tensors the kernel fills directly (on-device replay feed):
        obs [T,N,10], actions [T,N,2], reward [T,N], next_obs [T,N,10], terminated [T,N] u8, truncated [T,N] u8."""
        w = A.TvcActorWeights()
        keep = []
        for k in ("w1", "b1", "w2", "b2", "w3", "b3"):
            t = weights[k].to(device=self.device, dtype=torch.float32).contiguous()
            keep.append(t)
            setattr(w, k, t.data_ptr())
        io = A.TvcRolloutIO()
        out = dict(obs=self.obs, reward_sum=torch.zeros(self.n, dtype=torch.float32, device=self.device),
                   actions_last=torch.zeros((self.n, 2), dtype=torch.float32, device=self.device))
        io.obs, io.reward_sum, io.actions_last = out["obs"].data_ptr(), out["reward_sum"].data_ptr(), out["actions_last"].data_ptr()
        if record:
            out["actions_all"] = torch.zeros((T, self.n, 2), dtype=torch.float32, device=self.device)
            out["reward_all"] = torch.zeros((T, self.n), dtype=torch.float32, device=self.device)
            io.actions_all, io.reward_all = out["actions_all"].data_ptr(), out["reward_all"].data_ptr()
        if transitions is not None:
            io.obs_all, io.next_obs_all = transitions["obs"].data_ptr(), transitions["next_obs"].data_ptr()
            io.actions_all, io.reward_all = transitions["actions"].data_ptr(), transitions["reward"].data_ptr()
            io.terminated_all, io.truncated_all = transitions["terminated"].data_ptr(), transitions["truncated"].data_ptr()
        io.deterministic = int(deterministic)
        A.check(self.L.tvc_rollout(self.h, C.byref(w), int(T), C.byref(io), self._stream()), "tvc_rollout")
        out["_keep"] = keep
        return out
