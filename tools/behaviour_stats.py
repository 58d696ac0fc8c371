"""Behavioural acceptance statistic of the physics stand-in: over S terrain seeds, how far does the shipped policy carry the
character in T seconds, and how often does it fall?  Runs on the CPU oracle (the GPU path is checked against the oracle separately);
used by tests/test_behaviour_cpu.py and quoted in DESIGN §3.
  python tools/behaviour_stats.py [scene] [seeds] [seconds] [threads]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def run(scene, seeds=64, seconds=20.0, threads=8, seed0=1, fast=True):
    from pyoracle import Oracle
    pack = os.path.join(ROOT, "assets", scene + ".trlpack")
    sd = np.arange(seed0, seed0 + seeds, dtype=np.uint64)
    o = Oracle(pack, seeds, 0, terrain_seeds=sd, fast=fast)
    x0 = np.array([o.get_state(e)[0][0] for e in range(seeds)])
    first_fall_x = np.full(seeds, np.nan)       # distance of the first episode (until the first fall / reset), if any
    prev_x = x0.copy()
    for _ in range(int(round(seconds * 30))):
        o.update(1.0 / 30.0, threads)
        x = np.array([o.get_state(e)[0][0] for e in range(seeds)])
        reset = x < prev_x - 1.0                # a reset puts the character back to the start
        for e in np.nonzero(reset & np.isnan(first_fall_x))[0]:
            first_fall_x[e] = prev_x[e] - x0[e]
        prev_x = x
    no_fall = np.isnan(first_fall_x)
    dist = np.where(no_fall, prev_x - x0, first_fall_x)
    st = o.eval_stats()
    return {"scene": scene, "seeds": seeds, "seconds": seconds, "no_fall_frac": float(no_fall.mean()),
            "mean_first_episode_dist": float(dist.mean()), "median_first_episode_dist": float(np.median(dist)),
            "dist_no_fall_min": float(dist[no_fall].min()) if no_fall.any() else None,
            "dist_no_fall_mean": float(dist[no_fall].mean()) if no_fall.any() else None,
            "episodes": int(st["episodes"]), "avg_episode_dist": float(st["avg_dist"]), "dist": dist.tolist()}


if __name__ == "__main__":
    scene = sys.argv[1] if len(sys.argv) > 1 else "dog_slopes_mixed"
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 20.0
    threads = int(sys.argv[4]) if len(sys.argv) > 4 else 8
    r = run(scene, seeds, seconds, threads)
    d = np.array(r.pop("dist"))
    print(json.dumps(r))
    for thr in (43.0, 80.0):
        print(f"frac of seeds beyond {thr} m: {(d > thr).mean():.3f}")
