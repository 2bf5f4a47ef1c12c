#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched TVC step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2]'s per-GPU slab -- 262,144 envs per GPU (weak
scaling), Contract X with full domain randomisation, sensor noise, K=10 substeps per step, same-step
autoreset, pre-generated U(-1,1) actions resident in HBM -- because the metric is quoted "at 1/2/4/8
B200".  configs[1] (4,096 envs on one GPU, latency-bound) is reported beside it under "small_batch".
One "step" = one launch of the step kernel over the rank's env slab.

Timing: W warm-up steps, then K steps each bracketed by CUDA events on the launching stream; L2 is
flushed between timed steps by zeroing a 256 MiB buffer (not timed); ms_per_step is the MAX over
ranks of the mean per-step device time.  `value` = envs over all ranks / that time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 262144
ALGO_BYTES_PER_ENV_STEP_X = 366   # Contract X state r/w + DR params + delay ring + I/O
ALGO_BYTES_PER_ENV_STEP_R = 294
HBM_FALLBACK_GBS = 6650.0


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def _bf16_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["bf16_tflops_sustained"])
    except Exception:  # noqa: BLE001
        return 1400.0


def _issue_roofline(ms, clocks):
    """Second roofline of the step path: warp-instructions issued per clock and SM sub-partition (count from the committed
    ncu capture, time and clock measured live) against the issue limit of 1.0 and the measured FP32 ceilings."""
    wi = _traffic("warp_instructions_per_step")
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    if not wi:
        return None
    ipc = wi / (148 * 4 * ms * 1e-3 * mhz * 1e6)
    return {"bound": "issue", "achieved": ipc, "peak": 1.0, "unit": "warp-instructions/clk/SM sub-partition", "frac": ipc,
            "warp_instructions_per_step": wi,
            "fp32_ceilings": {"ffma_three_register_sources": 0.65, "ffma_two_shared_sources": 0.93,
                              "source": "tools/micro/fp32_rate.cu on this pool's B200 (profiles/r01_fp32_rate_microbench.txt)"},
            "note": "the step is a single wave of dependent chains (near-ground groups and airborne groups on 2,368 warp slots): "
                    "the instruction count and latency along each warp's chain bound it; see DESIGN.md section 6"}


STEP_PATH_SOURCES = ("tvc_abi.cu", "tvc_device.cuh", "tvc_internal.h")   # what step_kernel_v2 / close_kernel are compiled from


def csrc_hash() -> str:
    """sha256 (first 16 hex digits) over the step path's kernel sources: the committed ncu figures of the step kernel are only
    quoted for the build they were captured from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "tvc_ai_b200", "csrc")
    for f in STEP_PATH_SOURCES:
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def _traffic(key="dram_bytes_per_launch"):
    """A per-launch figure of the step path from the committed ncu --set full capture (DRAM bytes, warp-instructions), or
    None -- also None when the capture was taken from other kernel sources than the ones being measured (stale)."""
    p = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if d.get("csrc_sha16") != csrc_hash():
                return None
            return d.get(key)
        except Exception:  # noqa: BLE001
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


OVERRIDES = {}   # --override key=value (diagnostic what-if runs; the default line uses none)


def _workload_cfg(A, env_id_base=0, contract=None):
    contract = A.CONTRACT_X if contract is None else contract
    return A.default_config(contract, autoreset=1, env_id_base=env_id_base, **OVERRIDES)


def _oracle_cfg(O):
    return O.default_config(O.CONTRACT_X, autoreset=1)


def cpu_baseline(budget_s: float = 12.0):
    """The fp64 oracle port on the host cores (OpenMP over envs), same workload, bounded sample."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n = 4096
    sim = O.OracleSim(_oracle_cfg(O), n)
    sim.reset()
    acts = sim.random_actions(0)
    for w in range(2):
        sim.step_arrays(acts, threads=cores)
    t0 = time.perf_counter()
    sim.step_arrays(acts, threads=cores)
    one = max(time.perf_counter() - t0, 1e-6)
    steps = int(max(5, min(400, budget_s / one)))
    t0 = time.perf_counter()
    for t in range(steps):
        sim.step_arrays(sim.random_actions(t + 3), threads=cores)
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n} envs x {steps} steps, Contract X K=10, fp64 C oracle (oracle/tvc_oracle.c), OpenMP {cores} threads; "
                      "PyBullet reference not installable (SURVEY.md F2)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference itself
    (Python + PyBullet) cannot run in this image, so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n = 16384   # bounded sample of the 262144-env slab per "step"
    sim = O.OracleSim(_oracle_cfg(O), n)
    sim.reset()
    for w in range(args.warmup):
        sim.step_arrays(sim.random_actions(w), threads=cores)
    t0 = time.perf_counter()
    for t in range(args.steps):
        sim.step_arrays(sim.random_actions(args.warmup + t), threads=cores)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    line = {"impl": "reference", "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"bounded sample: {n} of the {ENVS_PER_GPU}-env slab per step, Contract X full DR, K=10, "
                                   "autoreset, Philox actions", "substeps": 10},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{n} envs x {args.steps} steps; fp64 C oracle, OpenMP {cores} threads "
                                       "(PyBullet reference not installable here)"},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist

    from tvc_ai_b200 import _abi as A
    from tvc_ai_b200 import dist as D
    from tvc_ai_b200.engine import BatchedEngine
    from tvc_ai_b200.vector_env import RocketTVCVectorEnv

    rank, world, local = D.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    affinity = D.bind_to_gpu_cpus(local)     # before any pinned allocation (first touch decides the NUMA node of the slabs)
    dev = torch.device("cuda", local)
    n = args.envs_per_gpu
    cfg_used = _workload_cfg(A, env_id_base=rank * n)
    eng = BatchedEngine(n, cfg_used, device=local)
    eng.reset()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    pool = [torch.rand((n, 2), generator=gen, device=dev) * 2 - 1 for _ in range(16)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream(device=dev)
    K, W = args.steps, max(args.warmup, 3)
    launches = 0
    # episode-statistics reduction (+ NCCL all-reduce over NVLink when N > 1): every 64 steps in production; a
    # short run still gets at least one inside the timed region
    stat_every = min(64, max(1, K // 2))
    stat_events = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # burn-in: every env starts at z = 1 m in lock-step; run until episodes have desynchronised so the
    # timed region sees the steady-state mix of free flight, ground contact and resets
    for b in range(args.burn_in):
        eng.step(pool[b % 16], want_final=False)
    eng.stats(reset_after=True)
    for w in range(W):
        eng.step(pool[w % 16], want_final=False)
    if world > 1:   # warm the NCCL communicator outside the timed region
        D.allreduce_stats(eng.stats_device(False))
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    wall0 = time.perf_counter()
    for k in range(K):
        flush.zero_()
        starts[k].record()
        eng.step(pool[k % 16], want_final=False)
        launches += 2            # step_kernel_v2 + the closing sort kernel (classify_kernel with the deferred resets)
        ends[k].record()
        if (k + 1) % stat_every == 0:   # side stream: does not delay the next step's launch
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0.record(side)
                D.allreduce_stats(eng.stats_device(False))
                s1.record(side)
                stat_events.append((s0, s1))
            launches += 1        # stats_reduce_kernel
    barrier()
    wall = time.perf_counter() - wall0
    per = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    stat_us = [1e3 * a.elapsed_time(b) for a, b in stat_events]
    ms = sum(per) / K
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_envs = n * world
    value = total_envs / (ms_max * 1e-3)

    # the same loop with the kernel's own Philox action stream (actions U(-1,1) from Philox4x32-10, counter =
    # (env id, step)) instead of the pre-generated torch.rand pool -- reported beside the headline
    torch.cuda.synchronize(dev)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ph_ms = 0.0
    for k in range(min(K, 20)):
        flush.zero_()
        pe0.record()
        eng.step(None, want_final=False)
        pe1.record()
        torch.cuda.synchronize(dev)
        ph_ms += pe0.elapsed_time(pe1)
    ph_ms /= min(K, 20)
    launches += 2 * min(K, 20)

    # warm-L2 number (no flush, back to back) -- reported beside the headline, not instead of it
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        eng.step(pool[k % 16], want_final=False)
    e1.record()
    torch.cuda.synchronize(dev)
    warm_ms = e0.elapsed_time(e1) / K

    stats = D.stats_dict(D.allreduce_stats(eng.stats_device(False)))
    torch.cuda.synchronize(dev)
    eng.close()

    # end-to-end through the public VectorEnv API with HOST (numpy) buffers, every rank (its own env slab): the
    # timed region contains, per step, the H2D copy of the actions and the D2H copies of obs / reward / flags /
    # final observations (tvc_step_host: pinned host buffers, stream sync inside the call)
    import numpy as np
    venv = RocketTVCVectorEnv(n, config={}, contract="X", device=local, final_info=False, copy_outputs=False,
                              env_id_base=rank * n, **OVERRIDES)
    venv.reset(seed=42)
    for b in range(args.burn_in):          # same steady-state mix as the device-timed region (device-resident actions, untimed)
        venv.engine.step(pool[b % 16], want_final=False)
    torch.cuda.synchronize(dev)            # tvc_step_host runs on the handle's own stream: drain the burn-in first
    # the step's inputs live in pinned host memory (what a host-side policy would write into `venv.pinned_actions()`)
    host_actions = [torch.from_numpy(np.random.default_rng(100 * rank + i).uniform(-1, 1, (n, 2)).astype(np.float32)).pin_memory().numpy()
                    for i in range(4)]
    for w in range(3):
        venv.step(host_actions[w % 4])
    ke = max(5, min(K, 30))
    barrier()
    t0 = time.perf_counter()
    for k in range(ke):
        o_h, r_h, te_h, tr_h, _ = venv.step(host_actions[k % 4])
    torch.cuda.synchronize(dev)
    e2e_dt = (time.perf_counter() - t0) / ke
    e2e_check = float(r_h[:16].sum())      # touch the result on the host
    done_rows = int((te_h | tr_h).sum())   # final-observation rows the last step wrote into the pinned host buffer
    venv.close()
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_dt = float(te.item())
    launches += 2 * ke
    e2e_rec = {"value": total_envs / e2e_dt, "unit": "env-steps/s", "h2d_bytes_per_step": n * 8,
                    "d2h_bytes_per_step": n * (40 + 4 + 1 + 1) + done_rows * 40, "ms_per_step": 1e3 * e2e_dt, "n_gpus_measured": world,
                    "bytes_are": "per GPU", "result_checksum": e2e_check, "cpu_affinity": affinity,
                    "d2h_note": f"obs 40 B + reward 4 B + 2 flag bytes per env by copy engine; the final observations of the {done_rows} "
                                "envs that ended an episode in the (last) step are stored by the kernel straight into the pinned host buffer",
                    "api": "RocketTVCVectorEnv.step(pinned numpy actions, copy_outputs=False) -> tvc_step_host (H2D from the pinned actions, one 46 B/env D2H into the pinned result slab, stream sync inside)"}

    clocks = sampler.stop()   # sampled from the start of the timed region to the end of the e2e measurement (all under load)
    extra = {"e2e": e2e_rec}
    if rank == 0:
        # configs[1]: 4,096 envs on one GPU (latency-bound), same Contract X settings
        small = BatchedEngine(4096, _workload_cfg(A), device=local)
        small.reset()
        acts = [torch.rand((4096, 2), device=dev) * 2 - 1 for _ in range(4)]
        for w in range(args.burn_in):
            small.step(acts[w % 4], want_final=False)
        torch.cuda.synchronize(dev)
        e0.record()
        for k in range(200):
            small.step(acts[k % 4], want_final=False)
        e1.record()
        torch.cuda.synchronize(dev)
        us = 1e3 * e0.elapsed_time(e1) / 200
        extra["small_batch"] = {"workload": "configs[1]: 4096 envs, 1 GPU, full DR, K=10, L2-resident by size",
                                "us_per_step": us, "env_steps_per_sec": 4096 / (us * 1e-6)}
        small.close()

        # the reference's own constants (Contract R: every quirk, K = 4 substeps of 0.005 s, no DR) on a slab of the headline
        # size, bit-ring diversity window, cold L2 -- what the path costs when it reproduces the shipped env instead of the
        # north_star's extension; 294 algorithmic bytes per env-step
        try:
            cr = BatchedEngine(n, A.default_config(A.CONTRACT_R, autoreset=1, diversity_mode=A.DIV_FAST), device=local)
            cr.reset()
            for w in range(1100):                      # past 1,000 pushes: the reward window is full
                cr.step(pool[w % len(pool)], want_final=False)
            torch.cuda.synchronize(dev)
            cr_ms = 0.0
            for k in range(30):
                flush.zero_()
                e0.record()
                cr.step(pool[k % len(pool)], want_final=False)
                e1.record()
                torch.cuda.synchronize(dev)
                cr_ms += e0.elapsed_time(e1) / 30
            peak_cr, _ = _peaks()
            extra["contract_r"] = {"workload": f"Contract R (reference constants, all quirks, K=4, no DR), {n} envs, same-step autoreset, "
                                               "bit-ring diversity window, L2 flushed between steps",
                                   "ms_per_step": cr_ms, "env_steps_per_sec": n / (cr_ms * 1e-3),
                                   "algorithmic_bytes_per_env_step": 294,
                                   "hbm_frac": 294 * n / (cr_ms * 1e-3) / 1e9 / peak_cr}
            cr.close()
        except Exception as exc:  # noqa: BLE001 -- an extra leg must not take the headline line down
            extra["contract_r"] = {"error": repr(exc)}

        # row S14 (reference training default enable_curiosity=True): the batch's curiosity term on the tensor cores, timed
        # alone on a slab of the headline size in its steady mix (cold L2) next to the fp32 torch path it replaces
        try:
            from tvc_ai_b200.env import CuriosityModule
            torch.manual_seed(0)
            fm = CuriosityModule(obs_dim=8, action_dim=2, device=dev).forward_model
            ce = BatchedEngine(n, _workload_cfg(A), device=local)
            ce.reset()
            cact = pool[0]
            for w in range(100):
                ce.step(pool[w % 16], want_final=True)
            ce.step(cact, want_final=True)
            prev_s8 = ce.obs[:, :8].clone()
            hasp = torch.ones(n, dtype=torch.uint8, device=dev)
            rew_c, intr_c = torch.empty(n, device=dev), torch.empty(n, device=dev)
            ce.curiosity(cact, prev_s8, hasp, rew_c, intrinsic=intr_c, forward_model=fm)
            cms, tms = [], []
            hb = hasp.bool()
            for k in range(20):
                flush.zero_()
                e0.record()
                ce.curiosity(cact, prev_s8, hasp, rew_c, intrinsic=intr_c)
                e1.record()
                torch.cuda.synchronize(dev)
                cms.append(e0.elapsed_time(e1))
                flush.zero_()
                e0.record()
                with torch.no_grad():
                    dn = (ce.terminated | ce.truncated).bool()
                    nx8 = torch.where(dn[:, None], ce.final_obs[:, :8], ce.obs[:, :8])
                    ii = 0.01 * ((fm(torch.cat([prev_s8, cact.clamp(-1.0, 1.0)], dim=1)) - nx8) ** 2).mean(dim=1)
                    _ = ce.reward + torch.where(hb, ii, torch.zeros_like(ii))
                    prev_s8.copy_(ce.obs[:, :8]); hb.copy_(~dn)
                e1.record()
                torch.cuda.synchronize(dev)
                tms.append(e0.elapsed_time(e1))
            cms.sort(); tms.sort()
            cflops = 2.0 * n * (16 * 256 + 256 * 256 + 256 * 16)
            extra["curiosity"] = {"workload": f"row S14 for {n} envs: forward model 10-256-256-8 (bf16 tcgen05.mma, fp32 accumulate in TMEM) + MSE + "
                                              "reward add + history update in one launch (tvc_curiosity), after the step",
                                  "ms_per_step": cms[len(cms) // 2], "mma_tflops": cflops / (cms[len(cms) // 2] * 1e-3) / 1e12,
                                  "torch_fp32_ms_per_step": tms[len(tms) // 2],
                                  "torch_what": "eager fp32 nn.Sequential forward (cuBLAS) + the elementwise kernels around it"}
            ce.close()
        except Exception as exc:  # noqa: BLE001
            extra["curiosity"] = {"error": f"{type(exc).__name__}: {exc}"}

        # configs[3]: fused rollout, 65,536 envs, SAC actor MLP 2x256 (bf16 tcgen05) inside the step loop, T=64
        if not args.no_rollout:
            nr, T = 65536, 64
            torch.manual_seed(0)
            nn = torch.nn
            net = nn.Sequential(nn.Linear(10, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 4)).to(dev)
            w = dict(w1=net[0].weight.detach(), b1=net[0].bias.detach(), w2=net[2].weight.detach(), b2=net[2].bias.detach(),
                     w3=net[4].weight.detach(), b3=net[4].bias.detach())
            ro = BatchedEngine(nr, _workload_cfg(A), device=local)
            ro.reset()
            for wu in range(4):                     # 256 steps: warm-up + episode desynchronisation
                ro.rollout(w, T)
            torch.cuda.synchronize(dev)
            reps = 5
            e0.record()
            for r in range(reps):
                ro.rollout(w, T)
            e1.record()
            torch.cuda.synchronize(dev)
            rms = e0.elapsed_time(e1) / reps
            flops = 138240.0 * nr * T
            extra["fused_rollout"] = {"workload": "configs[3]: 65536 envs, T=64 steps/launch, actor 10-256-256-4 (bf16 tcgen05.mma, "
                                                  "fp32 accumulate in TMEM), Contract X full DR, K=10",
                                      "ms_per_launch": rms, "env_steps_per_sec": nr * T / (rms * 1e-3),
                                      "actor_tflops": flops / (rms * 1e-3) / 1e12,
                                      "tensor_peak_tflops": _bf16_peak(), "note": "physics-bound: the MLP is a small share"}
            launches += 2 * (reps + 4)   # pack_actor_kernel + rollout_kernel per call
            # the same T steps UNFUSED, at its best: step path + one batched bf16 actor forward (cuBLAS) + the sampling
            # elementwise kernels per step, captured into ONE CUDA graph (no launch gaps) -- what the fusion has to beat
            try:
                netb = net.to(torch.bfloat16)
                act_buf = torch.zeros((nr, 2), device=dev)

                def unfused_step():
                    out4 = netb(ro.obs.to(torch.bfloat16)).float()
                    act_buf.copy_(torch.tanh(out4[:, :2] + out4[:, 2:].clamp(-20, 2).exp() * torch.randn((nr, 2), device=dev)))
                    ro.step(act_buf, want_final=False)
                cs = torch.cuda.Stream(device=dev)
                cs.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(cs), torch.no_grad():
                    for wu in range(4):
                        unfused_step()
                torch.cuda.current_stream(dev).wait_stream(cs)
                torch.cuda.synchronize(dev)
                ug = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ug), torch.no_grad():
                    unfused_step()
                for wu in range(8):
                    ug.replay()
                torch.cuda.synchronize(dev)
                e0.record()
                for k in range(T):
                    ug.replay()
                e1.record()
                torch.cuda.synchronize(dev)
                ums = e0.elapsed_time(e1)
                extra["fused_rollout"]["unfused_ms_per_T_steps"] = ums
                extra["fused_rollout"]["unfused_env_steps_per_sec"] = nr * T / (ums * 1e-3)
                extra["fused_rollout"]["unfused_what"] = ("per step, replayed as one CUDA graph: torch bf16 nn.Sequential forward (cuBLAS) + tanh / exp / randn "
                                                          "elementwise kernels + tvc_step (step_kernel_v2 + close_kernel)")
            except Exception as exc:  # noqa: BLE001
                extra["fused_rollout"]["unfused_error"] = f"{type(exc).__name__}: {exc}"
            launches += 2 * (T + 8)
            ro.close()

            # configs[4]: curriculum stage 6 with the end-to-end SAC loop -- fused rollout -> on-device replay ring ->
            # graph-replayed batched SAC updates (tvc_ai_b200/replay.py, sac.py), 4,096 envs, one GPU
            try:
                from tvc_ai_b200.curriculum import stage6_conditions
                from tvc_ai_b200.sac import SACConfig, train_sac
                e5 = BatchedEngine(4096, A.default_config(A.CONTRACT_X, autoreset=1, delay_steps=3, thrust_curve=1, propellant_fraction=0.2,
                                                          cg_burn_shift=0.05), device=local)
                e5.set_curriculum(stage6_conditions())
                e5.reset()
                scfg = SACConfig(batch_size=4096, learning_starts=4096 * 8, lr_actor=3e-4, lr_critic=3e-4)
                _, rp5, tm5 = train_sac(e5, 24, rollout_steps=8, config=scfg, seed=0)
                extra["config5"] = {"workload": "configs[4]: stage 6 (wind 3 N, mass +-30 %, tilt 0.7 rad, sensor noise, actuator delay 3, thrust curve), "
                                                "4096 envs, per iteration: tvc_rollout T=8 (transitions stored into the replay ring by the kernel) + 8 SAC "
                                                "updates of batch 4096 (tvc_replay_sample gather + PyTorch SAC update, one CUDA-graph replay)",
                                    "env_steps_per_sec_e2e": tm5["env_steps_per_sec_e2e"], "learner_share": tm5["learner_share"],
                                    "env_ms_per_iter": tm5["env_ms_per_iter"], "learner_ms_per_iter": tm5["learner_ms_per_iter"],
                                    "updates": tm5["updates"], "replay_filled": rp5.filled}
                launches += 24 * (2 + 8)     # pack + rollout + 8 gather kernels per iteration
                # the same loop with the reference's own update arithmetic (sac.reference_update: agent/multi_algorithm_agent.py:950-1016)
                e5.reset()
                _, rp5r, tm5r = train_sac(e5, 12, rollout_steps=8, config=SACConfig.reference_rule(batch_size=4096, learning_starts=4096 * 8), seed=0)
                extra["config5"]["reference_rule"] = {"env_steps_per_sec_e2e": tm5r["env_steps_per_sec_e2e"], "learner_share": tm5r["learner_share"],
                                                      "learner_ms_per_iter": tm5r["learner_ms_per_iter"], "updates": tm5r["updates"]}
                launches += 12 * (2 + 8)
                e5.close()
            except Exception as exc:  # noqa: BLE001 -- the learner is adjacent to the hot path: never lose the headline line over it
                extra["config5"] = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        peak, peak_src = _peaks()
        per_launch_bytes = ALGO_BYTES_PER_ENV_STEP_X * n
        achieved = per_launch_bytes / (ms * 1e-3) / 1e9      # this rank's kernel, GB/s
        line = {
            "metric": "env_steps_per_sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"configs[2] per-GPU slab: {n} envs/GPU x {world} GPU, Contract X (mass/thrust/cg/wind DR, "
                                   "sensor noise), K=10 substeps/step, same-step autoreset (final observations not requested in the "
                                   "device-timed loop), U(-1,1) actions pre-generated with torch.rand and resident in HBM, "
                                   f"episode-stat reduction every {stat_every} steps on a side stream ({len(stat_events)} inside the timed "
                                   "region; NCCL all-reduce when N>1)",
                       "envs_per_gpu": n, "substeps": int(cfg_used.substeps), "contact_iters": [int(cfg_used.contact_iters), int(cfg_used.contact_warm_iters)],
                       "quirks": hex(int(cfg_used.quirks)), "parallelism": f"env-slab x{world}",
                       "l2": "flushed between timed steps (256 MiB zero-fill, not timed)",
                       "arithmetic": "f32; envs within the contact margin of the ground also carry the attitude and the two "
                                     "cap-centre heights in f64 inside a step (DESIGN.md section 3), state in HBM is f32",
                       "burn_in_steps": args.burn_in, "overrides": dict(OVERRIDES)},
            "env_substeps_per_sec": value * 10,
            "stats_allreduce": {"count": len(stat_events), "us_mean": (sum(stat_us) / len(stat_us)) if stat_us else None,
                                "bytes": 128, "every_steps": stat_every,
                                "what": "stats_reduce_kernel + " + ("ncclAllReduce(16 x f64) over NVLink" if world > 1 else "no collective at N=1")
                                        + ", CUDA events on the side stream (rank 0)"},
            "philox_actions": {"ms_per_step": ph_ms, "env_steps_per_sec_per_gpu": n / (ph_ms * 1e-3),
                               "note": "same workload with actions=NULL: the kernel draws U(-1,1) actions from Philox4x32-10 (rank 0)"},
            "csrc_sha16": csrc_hash(),
            "warm_l2": {"ms_per_step": warm_ms, "env_steps_per_sec_per_gpu": n / (warm_ms * 1e-3),
                        "note": "back-to-back launches, state L2-resident (62 MB < 126 MB)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic(), "peak_source": peak_src,
                         "kernel": "step_kernel_v2<X=true,DIV=fast> (+ the closing sort / deferred-reset kernel)",
                         "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP_X,
                         "traffic_note": None if _traffic() is not None else "no ncu capture of this build is committed (profiles/step_kernel_traffic.json csrc_sha16 differs)",
                         "note": "not DRAM-bound: thousands of thread-instructions per env-step in dependent chains (block contact solve for "
                                 "the in-contact quarter of the envs, ten substeps, Euler angles, reward, Philox noise for all); "
                                 "see issue_roofline, DESIGN.md section 6 and profiles/"},
            "issue_roofline": _issue_roofline(ms, clocks),
            "gpu_launches": launches,
            "clocks": clocks,
            "region_wall_ms": 1e3 * wall,
            "episode_stats": {k: stats[k] for k in ("episodes", "successes", "steps", "term_crash", "term_tilt")},
        }
        line.update(extra)
        if "e2e" not in line:
            line["e2e"] = None
        if world == 1 or True:
            os.sched_setaffinity(0, all_cpus)      # the CPU baseline uses every host core
            line["cpu_baseline"] = cpu_baseline() if world == 1 else None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--no-rollout", action="store_true", help="skip the fused-rollout (config 4) measurement")
    ap.add_argument("--burn-in", type=int, default=400, help="untimed steps before warm-up (episode desynchronisation)")
    ap.add_argument("--override", action="append", default=[], help="tvc_config field=value (diagnostics only)")
    args = ap.parse_args()
    for kv in args.override:
        k, v = kv.split("=")
        OVERRIDES[k] = float(v) if "." in v else int(v)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
