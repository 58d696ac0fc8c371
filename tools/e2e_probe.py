import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepterrainrl_b200 as trl
np.set_printoptions(precision=0, suppress=True, linewidth=250)
n = 4096
sc = trl.ScenarioPoliEval(os.path.join(ROOT, "assets", "dog_slopes_mixed.trlpack"), n)
DT = 1 / 30
for chunk in range(6):
    ms = sc.BenchUpdates(50, DT, False) / 50
    st = sc._stats()
    ps, pd = sc.UpdateTimedDetail(DT)
    print("upd %d: %.3f ms/update episodes=%d" % ((chunk + 1) * 51, ms, st["episodes"]))
    print("  step us:", ps * 1e3)
    print("  decide us:", pd * 1e3)
