"""Launch timeline of one outer update in steady state (globaltimer stamps written by the kernels themselves; TRL_TRACE=1).
Usage: python tools/timeline_probe.py [envs] [presim_seconds]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ["TRL_TRACE"] = "1"
import deepterrainrl_b200 as trl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
presim = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sc = trl.ScenarioPoliEval(pack, n)
for _ in range(int(presim * 30)):
    sc.Update(1.0 / 30.0)
buf = np.zeros((128, 2), np.uint64)
sc.L.trl_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
sc._ck(sc.L.trl_debug_trace(sc.h, buf.ctypes.data_as(C.c_void_p), 128))     # re-arm
sc.Update(1.0 / 30.0)
sc._ck(sc.L.trl_debug_trace(sc.h, buf.ctypes.data_as(C.c_void_p), 128))
valid = buf[:, 1] > 0
t0 = int(buf[valid, 0].min())
names = {0: "T"}
for i in range(0, 22):
    names[1 + i] = f"S{i}"
    names[32 + i] = f"D{i}"
    names[64 + i] = f"C{i}"
rows = sorted((int(buf[k, 0]) - t0, int(buf[k, 1]) - t0, names.get(k, str(k))) for k in range(128) if valid[k])
for a, b, nm in rows:
    print(f"{nm:4s} start {a / 1e3:9.1f} us  end {b / 1e3:9.1f} us  dur {(b - a) / 1e3:8.1f} us")
print("update span %.1f us" % ((max(r[1] for r in rows)) / 1e3))
