"""Curriculum stage index logic and stage -> env-conditions schema (host side).

Restates /root/reference/scripts/curriculum_manager.py:61-163 (stage table from the YAML
`curriculum.stages` dict, conditions schema :76-94) and :191-246 (advance when at least half the
stage duration has elapsed and the evaluation success-rate / mean-reward thresholds are met).
Unlike the reference -- whose trainer mis-calls update() so the stage never advances and whose env
has no setter for the conditions -- `apply()` pushes the active stage's
conditions into the CUDA engine through tvc_set_curriculum (Contract X).

Under N>1 ranks every rank feeds the all-reduced statistics to `update_from_stats`, so the stage
index is bit-identical on all ranks.
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import Dict, List, Optional

STAGE6_NAME = "stage_6_full_realism"


@dataclass
class Stage:
    name: str
    duration_steps: int
    conditions: Dict
    success_criteria: Dict
    completed: bool = False
    performance_history: List[float] = field(default_factory=list)


def stages_from_config(curriculum_config: Dict) -> List[Stage]:
    """curriculum_manager.py:61-163.  PyYAML keeps the last of stage_1's duplicate `episodes`
    keys (200); durations are episodes * 1000 steps."""
    if not curriculum_config.get("enabled", False):
        return []
    raw = curriculum_config.get("stages", {})
    out: List[Stage] = []
    if isinstance(raw, dict):
        for key, data in raw.items():
            if not isinstance(data, dict):
                continue
            envc = data.get("environment", {}) or {}
            out.append(Stage(
                name=data.get("name", key),
                duration_steps=int(data.get("episodes", 200)) * 1000,
                conditions={
                    "max_initial_tilt": envc.get("initial_tilt_max", 0.1),
                    "max_initial_angular_vel": 0.1,
                    "domain_randomization": envc.get("mass_variation", 0) > 0,
                    "sensor_noise": False,
                    "max_gimbal_angle": 10.0,
                    "wind_enabled": envc.get("wind_force", 0) > 0,
                    "wind_force": envc.get("wind_force", 0),
                    "mass_variation": envc.get("mass_variation", 0),
                },
                success_criteria={"min_success_rate": envc.get("success_threshold", 0.7), "min_avg_reward": 100.0,
                                  "evaluation_episodes": 50}))
    elif isinstance(raw, list):
        for d in raw:
            out.append(Stage(d["name"], int(d["duration_steps"]), dict(d["conditions"]), dict(d["success_criteria"])))
    return out


def stage6_conditions(stage5: Optional[Dict] = None) -> Dict:
    """BASELINE config 5 "curriculum stage 6" does not exist in the reference (5 stages in the YAML).
    Contract X defines it as stage_5's conditions (config/config.yaml:279-286) + sensor noise on
    (+ actuator delay 3 control steps and per-episode thrust-curve scale, set on the engine config)."""
    c = dict(stage5 or {"max_initial_tilt": 0.7, "max_initial_angular_vel": 0.1, "domain_randomization": True,
                        "max_gimbal_angle": 10.0, "wind_enabled": True, "wind_force": 3.0, "mass_variation": 0.3})
    c["sensor_noise"] = True
    return c


class CurriculumManager:
    """Same public surface as the reference class for the parts on the path: update(),
    should_advance_stage(), advance_stage(), get_current_stage(), get_environment_config()."""

    def __init__(self, curriculum_config: Dict, logger: Optional[logging.Logger] = None):
        self.config = curriculum_config or {}
        self.logger = logger or logging.getLogger(__name__)
        self.stages = stages_from_config(self.config)
        self.current_stage_idx = 0
        self.current_step = 0
        self.evaluation_history: List[Dict] = []
        self.stage_transition_steps: List[int] = []

    def get_current_stage(self) -> Optional[Stage]:
        return self.stages[self.current_stage_idx] if self.current_stage_idx < len(self.stages) else None

    def get_environment_config(self) -> Dict:
        st = self.get_current_stage()
        return {} if st is None else st.conditions

    def _stage_start(self) -> int:
        return sum(s.duration_steps for s in self.stages[:self.current_stage_idx])

    def should_advance_stage(self, eval_metrics: Dict) -> bool:   # curriculum_manager.py:191-222
        st = self.get_current_stage()
        if st is None or st.completed:
            return False
        if self.current_step - self._stage_start() < st.duration_steps * 0.5:
            return False
        crit = st.success_criteria
        return (eval_metrics.get("eval_success_rate", 0.0) >= crit["min_success_rate"]
                and eval_metrics.get("eval_reward_mean", -float("inf")) >= crit["min_avg_reward"])

    def advance_stage(self) -> bool:   # curriculum_manager.py:224-246
        st = self.get_current_stage()
        if st:
            st.completed = True
            self.stage_transition_steps.append(self.current_step)
        self.current_stage_idx += 1
        return self.get_current_stage() is not None

    def update(self, step: int, eval_metrics: Optional[Dict] = None) -> Dict:   # curriculum_manager.py:248-291
        self.current_step = step
        st = self.get_current_stage()
        if st is None:
            return {}
        if eval_metrics:
            st.performance_history.append(eval_metrics.get("eval_reward_mean", 0.0))
            self.evaluation_history.append({"step": step, "stage": st.name, "metrics": dict(eval_metrics)})
            if self.should_advance_stage(eval_metrics):
                self.advance_stage()
        # as in the reference, the progress block below still refers to the stage that was current
        # on entry (name, duration) but to the *new* index and start step
        cfg = dict(self.get_environment_config())
        progress = (step - self._stage_start()) / st.duration_steps
        cfg["_curriculum_info"] = {"stage_name": st.name, "stage_index": self.current_stage_idx,
                                   "stage_progress": min(progress, 1.0), "total_stages": len(self.stages)}
        return cfg

    def is_curriculum_complete(self) -> bool:
        return self.current_stage_idx >= len(self.stages)

    # ---- coupling to the batched engine (what the reference never had) ----
    def update_from_stats(self, step: int, stats: Dict) -> Dict:
        """Drive update() from an (all-reduced) episode-statistics vector."""
        ep = max(stats.get("episodes", 0.0), 1.0)
        metrics = {"eval_success_rate": stats.get("successes", 0.0) / ep,
                   "eval_reward_mean": stats.get("sum_return", 0.0) / ep}
        return self.update(step, metrics if stats.get("episodes", 0.0) > 0 else None)

    def apply(self, target) -> Dict:
        """Push the active stage's conditions into a BatchedEngine / RocketTVCVectorEnv."""
        cond = self.get_environment_config()
        if cond:
            target.set_curriculum(cond)
        return cond
