import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deepterrainrl_b200 as trl
sc = trl.ScenarioPoliEval(os.path.join(ROOT, "assets", "dog_slopes_mixed.trlpack"), 4096)
sc.BenchUpdates(30, 1 / 30, False)
for n in (1,):
    ms = C.c_double(0)
    sc._ck(sc.L.trl_debug_time_decide(sc.h, n, 5, C.byref(ms)))
    print("pending %5d: %.1f us" % (n, ms.value * 1e3))
