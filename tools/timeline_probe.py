"""Per-launch timeline of one outer update in the engine's own (overlapped) schedule: what bounds the update -- the 21 step launches on
the main stream or the decision -> catch-up chain on the side stream?  (trl_update_timeline: an event pair around every launch.)
  python tools/timeline_probe.py [envs] [out.json]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepterrainrl_b200 as trl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
out = sys.argv[2] if len(sys.argv) > 2 else None
pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sc = trl.ScenarioPoliEval(pack, n)
sc.BenchUpdates(120, 1.0 / 30.0, flush_l2=False)
runs = []
for _ in range(5):
    sc.BenchUpdates(2, 1.0 / 30.0, flush_l2=False)
    runs.append(sc.UpdateTimeline())
tl = runs[-1]
names = {0: "T", 1: "S", 2: "D", 3: "C", 4: "F"}      # D = decision (conv stage in the batched build), F = its FC stage
dur = {k: [] for k in names}
for r in runs[1:]:
    for kind, idx, a, b in r:
        dur[kind].append(b - a)
total = [max(b for _, _, _, b in r) for r in runs[1:]]
summary = {"envs": n, "build": ("v1 one cluster per decision" if os.environ.get("TRL_DECIDE_V1") == "1" else "v2 batched conv + FC"),
           "update_ms_mean": float(np.mean(total)), "per_step_us": float(np.mean(total)) * 1e3 / 20,
           **{f"{names[k]}_us_mean": float(np.mean(v)) * 1e3 for k, v in dur.items() if v},
           **{f"{names[k]}_us_max": float(np.max(v)) * 1e3 for k, v in dur.items() if v}}
print(json.dumps(summary))
for kind, idx, a, b in tl[:22]:
    print(f"  {names[kind]}{idx:<3d} {a * 1e3:8.1f} -> {b * 1e3:8.1f} us  ({(b - a) * 1e3:6.1f})")
if out:
    json.dump({"summary": summary, "last": tl}, open(out, "w"))
