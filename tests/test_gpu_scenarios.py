"""More GPU parity cases: goat/cliffs (1 physics sub-step per env-step, gravity compensation off, init x, target 2 m/s),
set/get state round trip, weight upload, per-env reset, determinism, size-independent properties at the full batch."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

H = 1.0 / 600.0


def _relerr(a, b):
    return np.max(np.abs(a - b) / (1.0 + np.abs(b)))


def test_goat_cliffs_parity(assets):
    from pyoracle import Oracle
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "goat_cliffs.trlpack")
    n = 8
    seeds = (1 + 7919 * np.arange(n)).astype(np.uint64)       # SURVEY §8d config 5 seeding
    g = trl.ScenarioPoliEval(pack, n, terrain_seeds=seeds)
    o = Oracle(pack, n, 0, terrain_seeds=seeds)
    for env in range(n):
        for seg in (0, 1):
            gd, gmx, gfl = g.GetTerrain(env, seg)
            od, omx, ofl = o.terrain(env, seg)
            np.testing.assert_array_equal(gd, od)
            assert gmx == omx and gfl == ofl
    for k in range(30):
        g.Update(1.0 / 30.0)
        o.update(1.0 / 30.0, 4)
        if k % 5 == 4:
            gq, gqd = g.GetStateAll()
            for e in range(n):
                oq, oqd, _, _ = o.get_state(e)
                assert _relerr(gq[:, e], oq) < 1e-4 and _relerr(gqd[:, e], oqd) < 1e-4, (k, e)
    assert g._stats()["cycles"] == o.eval_stats()["cycles"]
    assert g._stats()["episodes"] == o.eval_stats()["episodes"]


def test_raptor_narrow_gaps_parity(assets):
    """BASELINE config 3 (raptor + narrow_gaps, MACE policy): per-step state / controller parity over 600 env-steps,
    then 3 s of the update loop incl. falls (stance flips, leg-swapped policy state, gated stance-hip PD)."""
    from pyoracle import Oracle
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "raptor_narrow_gaps.trlpack")
    n = 8
    g = trl.ScenarioPoliEval(pack, n)
    o = Oracle(pack, n, 0)
    assert g.state_size == 275 and g.num_dof == 21 and g.num_joints == 19
    worst = 0.0
    for k in range(600):
        g.EnvStep(H)
        for e in range(n):
            o.env_step(e, H)
        if k % 25 == 24 or k < 3:
            for e in range(n):
                gq, gqd, gt, gc = g.GetState(e)
                oq, oqd, ot, oc = o.get_state(e)
                worst = max(worst, _relerr(gq, oq), _relerr(gqd, oqd))
                assert _relerr(gq, oq) < 1e-4 and _relerr(gqd, oqd) < 1e-4, (k, e)
                assert _relerr(gt, ot) < 1e-4, (k, e)
                np.testing.assert_array_equal(gc, oc)
                gctl, octl = g.GetCtrl(e), o.get_ctrl(e)
                assert gctl.size == octl.size
                assert gctl[0] == octl[0] and gctl[-1] == octl[-1], (k, e, "fsm state / stance")
    np.testing.assert_allclose(g.GetPoliState(0), o.poli_state(0), rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(g.GetNetOut(0, 87), o.net_out(0, 87), rtol=1e-7, atol=1e-7)
    print("raptor 600 steps worst rel err", worst)
    g = trl.ScenarioPoliEval(pack, 32)
    o = Oracle(pack, 32, 0)
    for _ in range(90):
        g.Update(1.0 / 30.0)
        o.update(1.0 / 30.0, 8)
    assert g._stats()["cycles"] == o.eval_stats()["cycles"]
    assert g._stats()["episodes"] == o.eval_stats()["episodes"]


def test_set_state_roundtrip_and_single_env_reset(assets):
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    g = trl.ScenarioPoliEval(pack, 16)
    q0, qd0, _, _ = g.GetState(3)
    for _ in range(5):
        g.Update(1.0 / 30.0)
    q1, qd1, tau1, c1 = g.GetState(3)
    assert not np.allclose(q0, q1)
    other = g.GetState(4)[0].copy()
    g.Reset([3])
    q2, qd2, tau2, _ = g.GetState(3)
    np.testing.assert_array_equal(q2[2:], q0[2:])            # joint angles back to the state file
    np.testing.assert_array_equal(qd2, qd0)
    assert np.all(tau2 == 0)
    np.testing.assert_array_equal(g.GetState(4)[0], other)   # neighbours untouched
    g.SetState(5, q=q1, qd=qd1, tau=tau1, contact=c1)
    q3, qd3, tau3, c3 = g.GetState(5)
    np.testing.assert_array_equal(q3, q1); np.testing.assert_array_equal(qd3, qd1)
    np.testing.assert_array_equal(tau3, tau1); np.testing.assert_array_equal(c3, c1)


def test_determinism_and_weight_upload(assets):
    from pack_scene import read_pack, NET_LAYERS
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    a = trl.ScenarioPoliEval(pack, 64)
    b = trl.ScenarioPoliEval(pack, 64)
    p = read_pack(pack)
    blobs = []
    for name in NET_LAYERS:
        blobs += [p["net_" + name + "_w"], p["net_" + name + "_b"]]
    b.SetWeights(blobs, p["net_in_offset"], p["net_in_scale"], p["net_out_offset"], p["net_out_scale"])  # CopyModel
    for _ in range(15):
        a.Update(1.0 / 30.0); b.Update(1.0 / 30.0)
    qa, qda = a.GetStateAll(); qb, qdb = b.GetStateAll()
    np.testing.assert_array_equal(qa, qb)                    # bit-identical: deterministic kernels, same weights
    np.testing.assert_array_equal(qda, qdb)
    # zeroed actor heads must change behaviour (the upload is really used)
    blobs2 = [x.copy() for x in blobs]
    for i in range(14, 26):
        blobs2[i][:] = 0
    b.SetWeights(blobs2, p["net_in_offset"], p["net_in_scale"], p["net_out_offset"], p["net_out_scale"])
    for _ in range(30):
        a.Update(1.0 / 30.0); b.Update(1.0 / 30.0)
    assert not np.array_equal(a.GetStateAll()[0], b.GetStateAll()[0])


def test_full_batch_properties(assets):
    """4096 envs (BASELINE configs[1] size): counters add up, no NaNs, env 0..7 identical to a small batch with the
    same seeds (batch-size independence), dogs make forward progress."""
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 4096
    big = trl.ScenarioPoliEval(pack, n)
    small = trl.ScenarioPoliEval(pack, 8)
    for _ in range(30):
        big.Update(1.0 / 30.0); small.Update(1.0 / 30.0)
    qb, qdb = big.GetStateAll(); qs, qds = small.GetStateAll()
    assert np.all(np.isfinite(qb)) and np.all(np.isfinite(qdb))
    np.testing.assert_array_equal(qb[:, :8], qs)
    st = big._stats()
    assert st["steps"] == 30 * 20 * n
    assert st["cycles"] >= 2 * n
    d, e = big.GetDistLog()
    assert d.size == st["episodes"]
    assert np.median(qb[0]) > 2.5                            # ~4 m/s for 1 s


def test_explore_gather_block_views(assets):
    import torch
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    g = trl.ScenarioExpMACE(pack, 256)
    g.EnableExplore(1, 0.2, 0.025, 0.002)
    for _ in range(45):
        g.Update(1.0 / 30.0)
    g.Sync()
    rows, flags, env, count = g.DeviceTupleBlock()
    n = int(count.item())
    hr, hf, he = g.GetTuples(f64=True)
    assert n == hr.shape[0] > 0
    np.testing.assert_array_equal(rows[:n].cpu().numpy(), hr)
    np.testing.assert_array_equal(flags[:n].cpu().numpy().astype(np.uint32), hf)
    np.testing.assert_array_equal(env[:n].cpu().numpy(), he)
    g.ResetTupleBuffer(); g.Sync()
    assert g.GetNumTuples() == 0


def test_overlap_matches_serial(assets, monkeypatch):
    """The overlapped schedule (decisions + catch-up launches on a side stream, concurrent with the main step launch)
    must give the same per-env results as the serial schedule S_0 D_0 S_1 D_1 ...: same states, counters, tuples."""
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 512
    monkeypatch.setenv("TRL_SERIAL_SCHEDULE", "1")
    ser = trl.ScenarioExpMACE(pack, n, rng_seed=7)
    monkeypatch.setenv("TRL_SERIAL_SCHEDULE", "0")
    ovl = trl.ScenarioExpMACE(pack, n, rng_seed=7)
    for sc in (ser, ovl):
        sc.EnableExplore(True, 0.2, 0.025, 0.01)
    l0s, l0o = ser.KernelLaunches(), ovl.KernelLaunches()
    for _ in range(45):
        ser.Update(1.0 / 30.0); ovl.Update(1.0 / 30.0)
    assert ser.KernelLaunches() - l0s == 45 * 62 and ovl.KernelLaunches() - l0o == 45 * 100   # terrain + 21 steps + 20 x (conv + FC) decision launches; overlapped: 2 env groups x 20 step launches + S_end + 40 decision launches + 19 catch-ups
    # The catch-up launches are a second instantiation of the step kernel's source (csrc/trl_step_cg.cu: L1-bypassing loads, the
    # loop over the env-steps a pending env trails by): same operations in the same order, but the compiler's multiply-add
    # contraction may differ, so the two schedules agree to rounding amplified over 45 updates (observed 2.5e-12), not bit for bit.
    # On the emulator, where both builds are compiled without contraction, they are bit-identical (tests/test_simt_cpu.py).
    qa, qda = ser.GetStateAll(); qb, qdb = ovl.GetStateAll()
    np.testing.assert_allclose(qa, qb, rtol=0, atol=1e-8)
    np.testing.assert_allclose(qda, qdb, rtol=0, atol=1e-6)
    sa, sb = ser._stats(), ovl._stats()
    assert (sa["steps"], sa["cycles"], sa["episodes"]) == (sb["steps"], sb["cycles"], sb["episodes"])
    ra, fa, ea = ser.GetTuples(f64=True)
    rb, fb, eb = ovl.GetTuples(f64=True)
    assert ra.shape == rb.shape and ra.shape[0] > n
    # slot order depends on the atomic cursor; compare as sets of (env, flags, row)
    ka = np.lexsort(np.column_stack([ea, fa, ra]).T[::-1]); kb = np.lexsort(np.column_stack([eb, fb, rb]).T[::-1])
    np.testing.assert_array_equal(ea[ka], eb[kb])
    np.testing.assert_array_equal(fa[ka], fb[kb])
    np.testing.assert_allclose(ra[ka], rb[kb], rtol=0, atol=1e-7)


def test_env_groups_match_monolithic_launch(assets, monkeypatch):
    """TRL_GROUPS=G: every main step launch is split into G launches over contiguous env ranges on G streams that only meet at the
    decision / catch-up launches.  Same kernels, same per-env arithmetic: states, counters and tuples are bit-identical."""
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 1000
    scs = []
    for G in ("1", "3", "8"):
        monkeypatch.setenv("TRL_GROUPS", G)
        scs.append(trl.ScenarioExpMACE(pack, n, rng_seed=7))
    monkeypatch.delenv("TRL_GROUPS")
    for sc in scs:
        sc.EnableExplore(True, 0.2, 0.025, 0.01)
    l0 = [sc.KernelLaunches() for sc in scs]
    for _ in range(45):
        for sc in scs:
            sc.Update(1.0 / 30.0)
    assert [sc.KernelLaunches() - a for sc, a in zip(scs, l0)] == [45 * 80, 45 * 120, 45 * 220]
    ref = scs[0]
    ra, fa, ea = ref.GetTuples(f64=True)
    ka = np.lexsort(np.column_stack([ea, fa, ra]).T[::-1])
    assert ra.shape[0] > n
    for sc in scs[1:]:
        for a, b in zip(ref.GetStateAll(), sc.GetStateAll()):
            np.testing.assert_array_equal(a, b)
        assert ref._stats() == sc._stats()
        rb, fb, eb = sc.GetTuples(f64=True)
        kb = np.lexsort(np.column_stack([eb, fb, rb]).T[::-1])
        np.testing.assert_array_equal(ea[ka], eb[kb])
        np.testing.assert_array_equal(fa[ka], fb[kb])
        np.testing.assert_array_equal(ra[ka], rb[kb])


def test_two_scenes_interleaved(assets):
    """Two handles with different scenes alive in one process: the scene constants are re-uploaded on every switch, so
    interleaved updates give the same states as running each scene alone."""
    import deepterrainrl_b200 as trl
    dog = os.path.join(assets, "dog_slopes_mixed.trlpack")
    rap = os.path.join(assets, "raptor_narrow_gaps.trlpack")
    a = trl.ScenarioPoliEval(dog, 32)
    for _ in range(6):
        a.Update(1.0 / 30.0)
    qa = a.GetStateAll()[0].copy()
    b = trl.ScenarioPoliEval(rap, 32)
    for _ in range(6):
        b.Update(1.0 / 30.0)
    qb = b.GetStateAll()[0].copy()
    a2 = trl.ScenarioPoliEval(dog, 32)
    b2 = trl.ScenarioPoliEval(rap, 32)
    for _ in range(6):
        a2.Update(1.0 / 30.0); b2.Update(1.0 / 30.0)
    np.testing.assert_array_equal(a2.GetStateAll()[0], qa)
    np.testing.assert_array_equal(b2.GetStateAll()[0], qb)


def test_baseline_config_sizes(assets):
    """BASELINE.json configs[2] (raptor + narrow_gaps, 8192 envs on one GPU) and the per-GPU share of configs[4] (goat + cliffs,
    16384 envs over 8 GPUs = 2048 per GPU, mixed terrain seeds): size-independent properties at the full sizes."""
    import deepterrainrl_b200 as trl
    rap = trl.ScenarioPoliEval(os.path.join(assets, "raptor_narrow_gaps.trlpack"), 8192)
    small = trl.ScenarioPoliEval(os.path.join(assets, "raptor_narrow_gaps.trlpack"), 16)
    for _ in range(20):
        rap.Update(1.0 / 30.0); small.Update(1.0 / 30.0)
    q, qd = rap.GetStateAll()
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(qd))
    np.testing.assert_array_equal(q[:, :16], small.GetStateAll()[0])          # batch-size independence
    st = rap._stats()
    assert st["steps"] == 20 * 20 * 8192 and st["cycles"] >= 8192
    assert np.median(q[0]) > 1.5                                               # the raptor runs forward
    # ... and against the CPU oracle on a sample of the 8192 environments (same terrain seeds 1 + env)
    from pyoracle import Oracle
    sample = np.array([0, 1, 777, 4095, 4096, 6001, 8191])
    o = Oracle(os.path.join(assets, "raptor_narrow_gaps.trlpack"), len(sample), 0, terrain_seeds=(1 + sample).astype(np.uint64))
    for _ in range(20):
        o.update(1.0 / 30.0, 4)
    oq = np.stack([o.get_state(e)[0] for e in range(len(sample))], axis=1)
    assert np.max(np.abs(q[:, sample] - oq) / (1.0 + np.abs(oq))) < 1e-8
    del rap, small
    n = 2048
    seeds = (1 + 7919 * (3 * n + np.arange(n))).astype(np.uint64)              # rank 3's shard of the mixed-seed config
    goat = trl.ScenarioPoliEval(os.path.join(assets, "goat_cliffs.trlpack"), n, terrain_seeds=seeds)
    for _ in range(30):
        goat.Update(1.0 / 30.0)
    q, qd = goat.GetStateAll()
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(qd))
    st = goat._stats()
    assert st["steps"] == 30 * 20 * n and st["cycles"] >= n
    d, e = goat.GetDistLog()
    assert d.size == st["episodes"]
    sample = np.array([0, 5, 1023, 2047])
    o = Oracle(os.path.join(assets, "goat_cliffs.trlpack"), len(sample), 0, terrain_seeds=seeds[sample])
    for _ in range(30):
        o.update(1.0 / 30.0, 4)
    oq = np.stack([o.get_state(e)[0] for e in range(len(sample))], axis=1)
    assert np.max(np.abs(q[:, sample] - oq) / (1.0 + np.abs(oq))) < 1e-8          # incl. the episode resets of these 30 updates
