"""Freeze the reference CurriculumManager's stage decisions -- run in the BUILD CONTAINER only.

Imports /root/reference/scripts/curriculum_manager.py (pure Python) and /root/reference/config/
config.yaml, drives update(step, eval_metrics) with a fixed schedule and records the stage table
and the stage index after every call.  Output: tests/golden/curriculum.json (+ the YAML's
curriculum section, so the GPU box does not need /root/reference)."""
import importlib.util
import json
import os

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_curriculum", "/root/reference/scripts/curriculum_manager.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = yaml.safe_load(open("/root/reference/config/config.yaml"))
    cur = cfg["curriculum"]
    cm = mod.CurriculumManager(cur)
    table = [dict(name=s.name, duration_steps=s.duration_steps, conditions=s.conditions, success_criteria=s.success_criteria)
             for s in cm.stages]
    schedule, trace = [], []
    step = 0
    rates = [0.1, 0.5, 0.72, 0.9, 0.95, 0.6, 0.99, 0.8, 0.85, 0.92, 0.97, 0.99]
    for i in range(60):
        step += 50_000
        metrics = {"eval_success_rate": rates[i % len(rates)], "eval_reward_mean": 50.0 + 15.0 * (i % 9)} if i % 2 == 0 else None
        out = cm.update(step, metrics)
        info = out.get("_curriculum_info", {})
        schedule.append(dict(step=step, metrics=metrics))
        trace.append(dict(stage_index=cm.current_stage_idx, info=info,
                          conditions={k: v for k, v in out.items() if k != "_curriculum_info"}))
    json.dump(dict(curriculum_config=cur, stages=table, schedule=schedule, trace=trace),
              open(os.path.join(HERE, "curriculum.json"), "w"), indent=1)
    print("stages:", [(s["name"], s["duration_steps"]) for s in table], "final index", cm.current_stage_idx)


if __name__ == "__main__":
    main()
