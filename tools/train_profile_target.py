"""ncu target: a few trainer iterations on a filled replay memory (no rollout), so the launch list shows the trainer kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import deepterrainrl_b200 as trl  # noqa: E402

pack = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets", "dog_slopes_mixed.trlpack")
sc = trl.ScenarioExpMACE(pack, 64)
tr = trl.MACETrainer(sc, replay_mem_size=4096, num_init_samples=256, freeze_target_iters=2, seed=1)
rng = np.random.default_rng(0)
S, A = tr.S, tr.A
off, scale = tr.get("in_off"), tr.get("in_scale")
rows = np.zeros((1024, tr.W))
rows[:, 0] = rng.uniform(0, 8, 1024)
for k in (1, 1 + S + A):
    rows[:, k:k + S] = rng.normal(size=(1024, S)) / np.where(scale == 0, 1.0, scale) - off
rows[:, 1 + S] = rng.integers(0, 3, 1024)
rows[:, 2 + S:1 + S + A] = rng.normal(size=(1024, A - 1)) * 0.2
flags = np.where(rng.uniform(size=1024) < 0.5, 4, 0).astype(np.uint32)
tr.AddTuples(rows, flags)
tr.Train(int(sys.argv[1]) if len(sys.argv) > 1 else 4)
print(tr.counters())
