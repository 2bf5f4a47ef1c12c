"""Print the key numbers of a bench.py JSON line read from stdin (helper for gpurun one-liners)."""
import json
import sys

d = json.loads(sys.stdin.read().strip().splitlines()[-1])
tag = sys.argv[1] if len(sys.argv) > 1 else ""
out = [tag, "ms/step", round(d["ms_per_step"], 4), "warm", round(d["warm_l2"]["ms_per_step"], 4),
       "small_us", round(d["small_batch"]["us_per_step"], 1)]
if d.get("e2e"):
    out += ["e2e_ms", round(d["e2e"]["ms_per_step"], 2)]
if d.get("fused_rollout"):
    out += ["rollout_ms", round(d["fused_rollout"]["ms_per_launch"], 2)]
out += ["frac", round(d["roofline"]["frac"], 4), "episodes", d["episode_stats"]["episodes"]]
print(*out)
