"""The CPU oracle against golden vectors that the REFERENCE's own compiled code produced (tests/golden/ref_*.npz, written by
tools/make_ref_golden.py from oracle/_ref where /root/reference is mounted).  Unlike tests/test_ref_pinning_cpu.py these tests
need neither the reference tree nor oracle/_ref, so they also run wherever only the repository travelled."""
import ctypes as C
import os

import numpy as np
import pytest

from pyoracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
TYPES = ["flat", "gaps", "steps", "walls", "bumps", "mixed", "narrow_gaps", "slopes", "slopes_gaps", "slopes_walls", "slopes_steps",
         "slopes_mixed", "slopes_narrow_gaps", "cliffs"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("scene", ["dog_slopes_mixed", "goat_cliffs", "raptor_narrow_gaps"])
def test_oracle_reproduces_the_compiled_reference_scenario(assets, scene):
    """ref_scenario_<scene>.npz holds what the reference's compiled cScenarioPoliEval + controller stack computed, env-step by
    env-step, from the oracle's state sequence (lock-step harness, see tools/make_ref_golden.py): controller torques, gait state /
    phase, root position after every outer update (the reference's own reset included), final statistics.  The oracle, run alone
    from the same seed, must reproduce all of it."""
    g = np.load(os.path.join(GOLD, "ref_scenario_%s.npz" % scene))
    o = Oracle(os.path.join(assets, scene + ".trlpack"), 1, 0, terrain_seeds=[int(g["seed"])])
    o.L.orc_end_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
    tau, fsm, root = g["tau"], g["fsm"], g["root"]
    g = {k: g[k] for k in g.files}
    worst = 0.0
    k = 0
    for u in range(int(g["n_updates"])):
        for i in range(20):
            o.env_step(0, 1.0 / 600.0)
            to = o.last_tau(0)
            err = np.max(np.abs(tau[k] - to)) / max(1.0, np.max(np.abs(to)))
            worst = max(worst, err)
            assert err < 1e-9, (scene, k, err)
            held = o.get_state(0)[2]                     # what the physics receives: the reference's cJoint clamp vs the oracle's
            assert np.max(np.abs(g["applied"][k] - held)) <= 1e-9 * max(1.0, np.max(np.abs(held))), (scene, k)
            if i == 19:                                  # the fixture's last sample of an update is taken after the reference's
                o.L.orc_end_update(o.h, 0, 1.0 / 30.0)   # end-of-update handling (a reset puts the gait machine back to its start)
            oc = o.get_ctrl(0)
            assert int(fsm[k, 0]) == int(oc[0]) and abs(fsm[k, 1] - oc[1]) < 1e-9, (scene, k)
            k += 1
        assert np.max(np.abs(root[u] - o.get_state(0)[0][:3])) < 1e-9, (scene, u)
    es = o.eval_stats()
    assert (es["cycles"], es["episodes"]) == (int(g["stats"][0]), int(g["stats"][1]))
    assert abs(es["avg_dist"] - g["stats"][2]) < 1e-9
    assert np.allclose(o.dist_log(0), g["dist_log"], rtol=0, atol=1e-9)
    if scene == "goat_cliffs":
        assert es["episodes"] >= 2                     # the fixture covers the reference's fall handling and reset
    print(f"{scene}: {k} env-steps of compiled-reference torques reproduced, worst relative difference {worst:.1e}")


def test_oracle_terrain_equals_the_compiled_reference_generators(assets):
    """cTerrainGen2D (compiled) strips for all 14 terrain types x 2 seeds at the reference's default parameters: bit for bit."""
    g = np.load(os.path.join(GOLD, "ref_terrain.npz"))
    L = Oracle(os.path.join(assets, "dog_flat.trlpack"), 1, 0).L
    L.orc_terrain_build.argtypes = [C.c_int, C.c_void_p, C.c_ulong, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
    params = np.ascontiguousarray(g["params"])
    n_checked = 0
    for t, name in enumerate(TYPES):
        for seed in (3, 4242):
            want = g["%s_%d" % (name, seed)]
            buf = np.zeros(4096, np.float32)
            tw = C.c_double()
            n = L.orc_terrain_build(t, _p(params), seed, 40.0, _p(buf), 4096, C.byref(tw))
            assert n == len(want) and np.array_equal(buf[:n], want), (name, seed)
            n_checked += 1
    assert n_checked == 28
