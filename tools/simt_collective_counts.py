"""Dynamic warp-collective counts of the env-step path per build, measured on the SIMT emulator (tests/simt/): shuffled doubles,
32-bit shuffles, ballots and __syncwarp per warp and env-step in steady state.  Reproduces the table in
profiles/step_kernel_r02_experiment_queue.md.  Measured on the trl_env_step path (controller half and physics half in separate
launches), so TRL_REUSE_KIN, which needs the fused launch of trl_update, shows no effect here (there: -26 shuffled doubles).  No GPU needed:  python tools/simt_collective_counts.py [scene ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
from loader import simt_library  # noqa: E402
import deepterrainrl_b200 as trl  # noqa: E402
from deepterrainrl_b200.scenario import VARIANTS  # noqa: E402

NAMES = {2: "shfl64", 3: "shfl32", 4: "ballot", 5: "syncwarp"}
H = 1.0 / 600.0


def counts(defines, pack, n=4, steps=100):
    with simt_library(defines) as L:
        g = trl.ScenarioPoliEval(pack, n)
        for _ in range(60):                      # past the first decision
            g.EnvStep(H)
        c0 = {k: L.simt_counter(k) for k in NAMES}
        for _ in range(steps):
            g.EnvStep(H)
        c1 = {k: L.simt_counter(k) for k in NAMES}
        g.close()
    return {NAMES[k]: (c1[k] - c0[k]) / 32.0 / (n * steps) for k in NAMES}


if __name__ == "__main__":
    scenes = sys.argv[1:] or ["dog_slopes_mixed", "raptor_narrow_gaps"]
    for scene in scenes:
        pack = os.path.join(ROOT, "assets", scene + ".trlpack")
        print(scene)
        for name, defines in [("product", [])] + [(k, v) for k, v in VARIANTS.items() if "MIN_BLOCKS" not in " ".join(v) and "NOINLINE" not in " ".join(v)]:
            r = counts(defines, pack)
            print(f"  {name:24s} " + "  ".join(f"{k} {v:8.1f}" for k, v in r.items()))
