"""CPU tests of the oracle (no GPU): physical invariants of the restated dynamics, regression pins, terrain semantics."""
import os

import numpy as np
import pytest

from pyoracle import Oracle

H = 1.0 / 600.0


@pytest.fixture(scope="module")
def flat(assets):
    return os.path.join(assets, "dog_flat.trlpack")


@pytest.fixture(scope="module")
def mixed(assets):
    return os.path.join(assets, "dog_slopes_mixed.trlpack")


def test_mass_matrix_symmetric_total_mass(flat):
    o = Oracle(flat, 1, 0)
    M, C = o.rbd()
    assert np.abs(M - M.T).max() < 1e-12
    assert abs(M[0, 0] - 33.67) < 1e-9 and abs(M[1, 1] - 33.67) < 1e-9   # data/characters/dog.txt total mass
    assert np.all(np.linalg.eigvalsh(M) > 0)


def test_free_flight_com_follows_gravity(flat):
    """With no contacts, internal joint torques cannot move the COM: COM acceleration must equal g exactly."""
    o = Oracle(flat, 1, 0)
    q, qd, _, _ = o.get_state()
    q[1] += 2.0
    rng = np.random.default_rng(0)
    tau = np.zeros(23); tau[3:] = rng.normal(size=20) * 20
    o.set_state(q=q, qd=qd)
    qdd = o.forward_dynamics(tau)
    eps = 1e-5
    o.set_state(q=q + eps * qd + 0.5 * eps * eps * qdd, qd=qd + eps * qdd); _, v1 = o.com()
    o.set_state(q=q - eps * qd + 0.5 * eps * eps * qdd, qd=qd - eps * qdd); _, v2 = o.com()
    acc = (v1 - v2) / (2 * eps)
    assert abs(acc[0]) < 1e-3 and abs(acc[1] + 9.8) < 1e-3


def test_golden_flat_300(flat):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "dog_flat_300.npz"))
    o = Oracle(flat, 1, 0)
    i = 0
    for k in range(300):
        o.env_step(0, H)
        if k % 10 == 9:
            q, qd, tau, c = o.get_state(0)
            np.testing.assert_allclose(q, g["q"][i], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(qd, g["qd"][i], rtol=1e-8, atol=1e-8)
            np.testing.assert_array_equal(c, g["contact"][i])
            i += 1
    # the shipped bound gait must actually run in this physics: > 1.5 m in 0.5 s, still upright
    q, _, _, _ = o.get_state(0)
    assert q[0] > 1.5 and abs(q[2]) < 2.0


def test_flat_dog_bounds_for_five_seconds(flat):
    """Behavioural check of controller + physics: no fall, ~4 m/s (cDogController::GetTargetVel) on flat ground."""
    o = Oracle(flat, 1, 0)
    for _ in range(150):
        o.update(1.0 / 30.0)
    st = o.eval_stats()
    q, _, _, _ = o.get_state(0)
    assert st["episodes"] == 0
    assert 15.0 < q[0] < 26.0
    assert st["cycles"] >= 9


def test_terrain_golden_and_structure(mixed):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "terrain_slopes_mixed.npz"))
    o = Oracle(mixed, 3, 0, terrain_seeds=[1, 2, 12345])
    for e in range(3):
        d0, mx0, _ = o.terrain(e, 0)
        d1, mx1, _ = o.terrain(e, 1)
        np.testing.assert_array_equal(d0, g[f"e{e}s0"])
        np.testing.assert_array_equal(d1, g[f"e{e}s1"])
        assert mx0 == g[f"e{e}s0_minx"][0] and mx1 == g[f"e{e}s1_minx"][0]
        # segment 0 ends at x = -1 with height 0, segment 1 starts there with a 2 m flat pad (sim/GroundVar2D.cpp:239-355)
        sp = float(np.float32(0.1))
        assert abs(mx0 + (d0.size - 1) * sp - (-1.0)) < 1e-9 and mx1 == -1.0
        assert d0[-1] == 0.0 and np.all(d1[:21] == 0.0)
        assert d0.size >= 201 and d1.size >= 221
    assert not np.array_equal(g["e0s1"], g["e1s1"])


def test_terrain_streams_forward(mixed):
    o = Oracle(mixed, 1, 0)
    q, qd, _, _ = o.get_state(0)
    d1, mx1, fl = o.terrain(0, 1)
    end_x = mx1 + (d1.size - 1) * float(np.float32(0.1))
    q2 = q.copy(); q2[0] = end_x - 10.5     # view window [x-2, x+11] crosses the end of the max segment
    q2[1] = o.sample_height(q2[0]) + 3.0
    o.set_state(q=q2, qd=np.zeros(23))
    o.env_step(0, H)
    dn, mxn, fln = o.terrain(0, 0)
    assert fln != fl
    assert abs(mxn - end_x) < 1e-9            # new segment starts where the old one ended ...
    assert dn[0] == d1[-1]                    # ... at the same height


def test_first_decision_golden(mixed):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "first_decision.npz"))
    o = Oracle(mixed, 3, 0, terrain_seeds=[1, 2, 12345])
    o.env_step(0, H)
    np.testing.assert_allclose(o.poli_state(0), g["poli_state"], rtol=1e-10, atol=1e-10)
    y = o.net_out(0)
    np.testing.assert_allclose(y, g["net_out"], rtol=1e-9, atol=1e-9)
    # restated net + normalisation reproduces its own probe
    np.testing.assert_allclose(o.net_eval(o.poli_state(0)), y, rtol=1e-12)
    ctrl = o.get_ctrl(0)
    a = int(np.argmax(y[:3]))
    assert int(ctrl[11]) == a
    np.testing.assert_allclose(ctrl[12 + 1 + 1:12 + 30], y[3 + 29 * a + 1:3 + 29 * (a + 1)], rtol=1e-12)  # params[2:]
    assert ctrl[12 + 1] == abs(y[3 + 29 * a])                                                            # |Cv|


def test_poli_eval_golden(mixed):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "poli_eval_3s.npz"))
    o = Oracle(mixed, 4, 0)
    for _ in range(90):
        o.update(1.0 / 30.0, 2)
    st = o.eval_stats()
    assert st["cycles"] == int(g["cycles"][0]) and st["episodes"] == int(g["episodes"][0])
    x = np.array([o.get_state(e)[0][0] for e in range(4)])
    np.testing.assert_allclose(x, g["x"], rtol=1e-6, atol=1e-6)
    assert st["steps"] == 90 * 20 * 4


def test_exploration_tuples_structure(mixed):
    o = Oracle(mixed, 8, 1)
    o.set_explore(1, 0.5, 0.025, 0.05)
    for _ in range(75):
        o.update(1.0 / 30.0, 4)
    rows, flags, ids = o.tuples()
    assert rows.shape[0] > 8 and rows.shape[1] == 1 + 283 + 30 + 283
    assert np.all((rows[:, 0] >= 0) & (rows[:, 0] <= 1.0))              # reward in [0, 1]
    assert np.all(np.isin(rows[:, 1 + 283], [0, 1, 2]))                 # actor index
    assert np.any(flags & 0b110)                                        # some exploration flags set
    ok = (flags & 1) == 0
    # consecutive tuples of one env chain: s'(k) == s(k+1) unless a reset intervened
    for e in np.unique(ids):
        r = rows[ids == e]
        f = flags[ids == e]
        for k in range(len(r) - 1):
            if not (f[k] & 1):
                np.testing.assert_array_equal(r[k][1 + 283 + 30:], r[k + 1][1:1 + 283])
    assert ok.sum() > 0


def test_raptor_runs_and_flips_stance(assets):
    """Behavioural check of the raptor controller restatement: with the shipped narrow_gaps policy the raptor runs at
    ~4 m/s, alternating stance legs every cycle (sim/RaptorController.cpp:837-840)."""
    o = Oracle(os.path.join(assets, "raptor_narrow_gaps.trlpack"), 4, 0)
    stances = []
    for k in range(60):
        o.update(1.0 / 30.0, 4)
        stances.append(int(o.get_ctrl(0)[-1]))
    x = np.array([o.get_state(e)[0][0] for e in range(4)])
    st = o.eval_stats()
    assert np.sum(x > 6.0) + st["episodes"] >= 4 and np.max(x) > 7.0
    assert 0 in stances and 1 in stances
    assert st["cycles"] >= 4 * 6
