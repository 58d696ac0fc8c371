"""In-kernel phase stamps of the batched decision path's FC kernel (trl_debug_fc_phases) for 16 and 32 pending decisions."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepterrainrl_b200 as trl  # noqa: E402

pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sc = trl.ScenarioPoliEval(pack, 4096)
sc.BenchUpdates(60, 1.0 / 30.0, flush_l2=False)
names = ["entry", "staged", "tip0 tiles", "sync", "concat0", "sync", "ip0", "sync", "heads0", "sync", "heads1", "sync", "decisions", "sync"]
for m in (1, 16, 32, 64):
    out = np.zeros(16); cu = C.c_double(0); fu = C.c_double(0)
    rc = sc.L.trl_debug_fc_phases(sc.h, m, out.ctypes.data_as(C.c_void_p), C.byref(cu), C.byref(fu))
    assert rc == 0, sc.L.trl_last_error().decode()
    d = np.diff(np.concatenate([[0.0], out[:14]]))
    print(f"M={m}: conv launch {cu.value:.1f} us, fc launch {fc.value if False else fu.value:.1f} us; fc phases (us): " +
          ", ".join(f"{n} {x / 1e3:.1f}" for n, x in zip(names, d)))
