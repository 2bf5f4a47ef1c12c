"""On-device replay ring fed by the fused rollout kernel.

Replaces the reference's batch-of-1 feed (scripts/train.py:574-584: one dict of five 1-row tensors per env step, built on the
host and copied to the device) by a ring of transitions that never leaves HBM:

* `DeviceReplay.write_block(T)` hands `BatchedEngine.rollout(..., transitions=...)` views of the ring at its head, so
  `tvc_rollout` stores (obs, action, reward, next_obs, terminated, truncated) of T steps x N envs IN PLACE -- no copy;
* `DeviceReplay.sample_into(batch)` draws uniform indices from Philox4x32-10 (key = seed, counter = (row, draw)) and gathers
  the five arrays into the learner's static batch tensors in ONE launch of `replay_gather_kernel` (`tvc_replay_sample`,
  csrc/tvc_replay.cu): reward * reward_scale, done = terminated as float -- the inputs of
  agent/multi_algorithm_agent.py:950-1016 (`_update_sac`).  Static output tensors make the call CUDA-graph capturable.

The ring is a whole number of rollout blocks (capacity = blocks * T * N), so a block never wraps.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi as A


class DeviceReplay:
    def __init__(self, num_envs: int, block_steps: int, capacity: int = 1 << 20, device: int | torch.device = 0, seed: int = 0,
                 reward_scale: float = 1.0):
        if num_envs < 1 or block_steps < 1:
            raise ValueError("num_envs and block_steps must be positive")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceReplay lives in HBM: a CUDA device is required (no CPU fallback)")
        self.n, self.T = int(num_envs), int(block_steps)
        self.block = self.n * self.T
        self.blocks = max(1, int(capacity) // self.block)
        self.capacity = self.blocks * self.block
        self.seed, self.reward_scale = int(seed), float(reward_scale)
        self.L = A.load()
        z = lambda *shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=self.device)  # noqa: E731
        self.obs, self.actions, self.reward = z(self.capacity, 10), z(self.capacity, 2), z(self.capacity)
        self.next_obs = z(self.capacity, 10)
        self.terminated, self.truncated = z(self.capacity, dt=torch.uint8), z(self.capacity, dt=torch.uint8)
        self.head_block = 0     # next block to be written
        self.filled = 0         # transitions available to the sampler
        self.draws = 0          # Philox counter: one per sample_into call
        # device-side control words {filled, draw_base}: read by the gather kernel, so that launches captured into a CUDA
        # graph follow the ring as it fills and draw fresh indices at every replay (`tick()` advances draw_base on the stream)
        self.ctl = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._ring = A.TvcReplayRing(self.obs.data_ptr(), self.actions.data_ptr(), self.reward.data_ptr(), self.next_obs.data_ptr(),
                                     self.terminated.data_ptr(), self.capacity)

    # ------------------------------------------------------------------ writer side
    def write_block(self) -> dict:
        """Views of the ring at its head, shaped for `BatchedEngine.rollout(transitions=...)`; call `commit()` afterwards."""
        lo, T, n = self.head_block * self.block, self.T, self.n
        v = lambda t, *tail: t[lo:lo + self.block].view(T, n, *tail)  # noqa: E731
        return dict(obs=v(self.obs, 10), actions=v(self.actions, 2), reward=v(self.reward), next_obs=v(self.next_obs, 10),
                    terminated=v(self.terminated), truncated=v(self.truncated))

    def commit(self):
        self.head_block = (self.head_block + 1) % self.blocks
        self.filled = min(self.filled + self.block, self.capacity)
        self.ctl[0:1].fill_(self.filled)     # stream-ordered behind the rollout that wrote the block

    def collect(self, engine, weights: dict, deterministic: bool = False):
        """One fused-rollout launch (T steps of every env with the actor in the loop) written straight into the ring."""
        if engine.n != self.n:
            raise ValueError(f"engine has {engine.n} envs, the ring was sized for {self.n}")
        out = engine.rollout(weights, self.T, deterministic=deterministic, transitions=self.write_block())
        self.commit()
        return out

    # ------------------------------------------------------------------ sampler side
    def new_batch(self, batch: int, with_indices: bool = False) -> dict:
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=self.device)  # noqa: E731
        b = dict(obs=z(batch, 10), actions=z(batch, 2), reward=z(batch), next_obs=z(batch, 10), done=z(batch))
        if with_indices:
            b["indices"] = torch.zeros(batch, dtype=torch.int64, device=self.device)
        return b

    def tick(self, n: int = 1):
        """Advance the device-side draw base by n (a stream-ordered torch op: capturable)."""
        self.ctl[1:2].add_(n)

    def sample_into(self, batch: dict, draw: int | None = None, filled: int | None = None, device_ctl: bool = False):
        """Gather a uniform sample of the filled part of the ring into `batch` (tensors from `new_batch`); one kernel launch on
        the current stream.  `draw` fixes the Philox counter (default: a running count).  device_ctl=True: `filled` and the
        draw base come from the device-side control words (for CUDA-graph capture; `draw` is then an offset)."""
        filled = self.filled if filled is None else int(filled)
        if filled < 1:
            raise RuntimeError("the replay ring is empty")
        if draw is None:
            draw = 0 if device_ctl else self.draws
            if not device_ctl:
                self.draws += 1
        out = A.TvcReplayBatch(batch["obs"].data_ptr(), batch["actions"].data_ptr(), batch["reward"].data_ptr(),
                               batch["next_obs"].data_ptr(), batch["done"].data_ptr(),
                               batch["indices"].data_ptr() if "indices" in batch else None)
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        A.check(self.L.tvc_replay_sample(C.byref(self._ring), filled, int(batch["reward"].numel()), self.seed, int(draw),
                                         C.c_void_p(self.ctl.data_ptr()) if device_ctl else None, self.reward_scale, C.byref(out), self.device.index or 0, stream), "tvc_replay_sample")
        return batch
