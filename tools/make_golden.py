"""Generates the small golden fixtures under tests/golden/ from the CPU oracle (regression pins of the oracle itself).

These fixtures pin the restated oracle against drift -- they are not reference outputs (DeepTerrainRL ships no tests or golden
vectors).  Reference-side fixtures exist too: tools/make_ref_golden.py writes tests/golden/ref_*.npz from the reference's own
sources compiled into oracle/_ref.

    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyoracle import Oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    # 1. BASELINE config 1: dog, flat, fixed action, 300 env-steps
    o = Oracle(os.path.join(ROOT, "assets", "dog_flat.trlpack"), 1, 0)
    qs, qds, taus, cs = [], [], [], []
    for k in range(300):
        o.env_step(0, 1.0 / 600.0)
        q, qd, tau, c = o.get_state(0)
        if k % 10 == 9:
            qs.append(q); qds.append(qd); taus.append(tau); cs.append(c)
    np.savez_compressed(os.path.join(OUT, "dog_flat_300.npz"), q=np.array(qs), qd=np.array(qds), tau=np.array(taus),
                        contact=np.array(cs))
    # 2. terrain strips (libstdc++ RNG semantics) for three seeds, slopes_mixed
    o = Oracle(os.path.join(ROOT, "assets", "dog_slopes_mixed.trlpack"), 3, 0, terrain_seeds=[1, 2, 12345])
    t = {}
    for e in range(3):
        for s in range(2):
            d, mx, fl = o.terrain(e, s)
            t[f"e{e}s{s}"] = d
            t[f"e{e}s{s}_minx"] = np.array([mx])
    np.savez_compressed(os.path.join(OUT, "terrain_slopes_mixed.npz"), **t)
    # 3. first policy decision of env 0: policy state, net output, decoded action
    o.env_step(0, 1.0 / 600.0)
    np.savez_compressed(os.path.join(OUT, "first_decision.npz"), poli_state=o.poli_state(0), net_out=o.net_out(0),
                        ctrl=o.get_ctrl(0))
    # 4. 3 s of policy evaluation, 4 envs: root x and counters
    o = Oracle(os.path.join(ROOT, "assets", "dog_slopes_mixed.trlpack"), 4, 0)
    for k in range(90):
        o.update(1.0 / 30.0, 4)
    st = o.eval_stats()
    np.savez_compressed(os.path.join(OUT, "poli_eval_3s.npz"), x=np.array([o.get_state(e)[0][0] for e in range(4)]),
                        cycles=np.array([st["cycles"]]), episodes=np.array([st["episodes"]]))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
