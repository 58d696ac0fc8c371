"""Target for ncu: rolls 4096 dog/slopes_mixed envs to a steady state (untimed, not graph-launched so every kernel is a
separate launch), then runs a few more outer updates for the profiler to sample.

  ncu --set full --clock-control none --import-source on -k regex:trl_step_kernel -s 2600 -c 2 -o gpurun_out/prof python tools/profile_target.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepterrainrl_b200 as trl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pack = sys.argv[2] if len(sys.argv) > 2 else "dog_slopes_mixed.trlpack"
sc = trl.ScenarioPoliEval(os.path.join(ROOT, "assets", pack), n)
for _ in range(125):            # ~4.2 s simulated: 125 x 42 launches = 5250 kernel launches
    sc.UpdateTimed(1.0 / 30.0)
sc.Sync()
print(sc._stats())
