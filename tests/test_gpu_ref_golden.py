"""The CUDA path (through the C ABI) against golden vectors the REFERENCE's own compiled code produced (tests/golden/ref_scenario_*.npz,
tools/make_ref_golden.py): the torques the reference's compiled controller stack + cJoint clamp would hand to the physics, and its
gait-machine state, env-step by env-step, for a policy-evaluation run from the reference's arg file.  The fixtures were computed
from the CPU oracle's state sequence; the CUDA path runs its own physics from the same seed, so agreement here needs both the
controller parity and the state parity (north_star tolerance 1e-4; seen ~1e-11 against the oracle)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H = 1.0 / 600.0
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("scene", ["dog_slopes_mixed", "raptor_narrow_gaps"])
def test_cuda_path_reproduces_the_compiled_reference(assets, scene):
    import deepterrainrl_b200 as trl
    g = np.load(os.path.join(GOLD, "ref_scenario_%s.npz" % scene))
    assert int(g["stats"][1]) == 0          # no episode ends inside the fixture: EnvStep carries no end-of-update handling
    applied, fsm = g["applied"], g["fsm"]
    sc = trl.ScenarioPoliEval(os.path.join(assets, scene + ".trlpack"), 1, terrain_seeds=[int(g["seed"])])
    worst = 0.0
    n = min(600, len(applied))
    for k in range(n):
        sc.EnvStep(H)
        _, _, tau, _ = sc.GetState(0)
        err = np.max(np.abs(tau - applied[k]) / (1.0 + np.abs(applied[k])))
        worst = max(worst, err)
        assert err < 1e-4, (scene, k, err)
        assert int(sc.GetCtrl(0)[0]) == int(fsm[k, 0]), (scene, k, "gait state")
    print(f"{scene}: {n} env-steps, worst relative torque difference to the compiled reference {worst:.1e}")
