"""The native multi-GPU exchange (csrc/trl_comm.cu) on a real device: NCCL opened by the library itself, pack kernel + all-gather
+ trainer hand-over at world size 1 (a single-rank communicator runs the same code path; the N-rank run is tools/train_multi.py,
measured by bench.py --gpus N), and the tuple-block overflow accounting of the C ABI."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _uid(L):
    buf = (C.c_ubyte * 128)()
    assert L.trl_comm_unique_id(buf) == 0, L.trl_last_error().decode()
    return bytes(buf)


def test_native_gather_world1_equals_local_hand_over(assets):
    """trl_gather_tuples + trl_trainer_add_gathered over NCCL == trl_trainer_add_from_scene (the single-GPU hand-over),
    bit for bit: same replay rows, same buffers, same weights after training."""
    import deepterrainrl_b200 as trl
    from deepterrainrl_b200 import parallel
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 256
    kw = dict(replay_mem_size=4096, num_init_samples=64, freeze_target_iters=3, seed=9)
    a = trl.ScenarioExpMACE(pack, n, rng_seed=77); ta = trl.MACETrainer(a, **kw)
    b = trl.ScenarioExpMACE(pack, n, rng_seed=77); tb = trl.MACETrainer(b, **kw)
    comm = parallel.Comm(b, 0, 1, backend="nccl", unique_id=_uid(b.L))
    comm.SetEnvOffset(1000)
    for sc in (a, b):
        sc.EnableExplore(True, 0.3, 0.025, 0.02)
    seen = 0
    for k in range(40):
        a.Update(); b.Update()
        if k % 4 == 3:
            loc, lf, le = b.GetTuples()
            ta.AddTuplesFromScene()
            comm.GatherTuples(block_rows=64)              # blocks of 64: bursts of more than 64 tuples drain over several calls
            counts, rows, flags, env = comm.Fetch()
            m = int(counts[0])
            assert m == min(len(loc), 64) and b.GetNumTuples() == len(loc) - m
            np.testing.assert_array_equal(rows, loc[:m]); np.testing.assert_array_equal(env, le[:m] + 1000)
            np.testing.assert_array_equal(flags & 0x7fffffff, lf[:m])
            comm.AddGathered(tb)
            while b.GetNumTuples() > 0:
                comm.GatherTuples(block_rows=64); comm.AddGathered(tb)
            seen += len(loc)
            assert comm.LastGatherMs() >= 0.0
        if k >= 30:
            ta.Train(1); tb.Train(1)
    ca, cb = ta.counters(), tb.counters()
    assert seen > 100 and ca["total"] == cb["total"] == seen and ca["iter"] == cb["iter"] == 10
    # identical arrival order needs the burst to fit one block; with a queue the order differs only in WHEN a tuple arrives, so
    # compare as multisets of rows, then the exact state for the run where every burst fitted
    ra, fa = ta.rows(np.arange(ca["num"])); rb, fb = tb.rows(np.arange(cb["num"]))
    key = lambda r: r[np.lexsort(r.T[::-1])]
    np.testing.assert_array_equal(key(ra), key(rb))
    assert comm.ReplicaSpread(tb) == 0.0
    st = comm.EvalStats(); loc = b._stats()
    assert st["steps"] == loc["steps"] == n * 40 * 20 and st["cycles"] == loc["cycles"]
    comm.BroadcastTrainer(tb, 0)
    assert comm.TuplesDropped() == 0
    comm.close()


def test_native_gather_exact_when_bursts_fit(assets):
    import deepterrainrl_b200 as trl
    from deepterrainrl_b200 import parallel
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 128
    kw = dict(replay_mem_size=2048, num_init_samples=64, freeze_target_iters=3, seed=9)
    a = trl.ScenarioExpMACE(pack, n, rng_seed=5); ta = trl.MACETrainer(a, **kw)
    b = trl.ScenarioExpMACE(pack, n, rng_seed=5); tb = trl.MACETrainer(b, **kw)
    comm = parallel.Comm(b, 0, 1, backend="nccl", unique_id=_uid(b.L))
    for sc in (a, b):
        sc.EnableExplore(True, 0.3, 0.025, 0.02)
    for k in range(45):
        a.Update(); b.Update()
        ta.AddTuplesFromScene()
        comm.GatherTuples(block_rows=n); comm.AddGathered(tb)       # an env finishes at most one cycle per update: always fits
        if k >= 35:
            ta.Train(1); tb.Train(1)
    ca, cb = ta.counters(), tb.counters()
    assert ca == cb and ca["total"] > 100 and ca["iter"] == 10
    for which in ("critic", "actor", "actor_batch"):
        np.testing.assert_array_equal(ta.lists(which), tb.lists(which))
    ra, fa = ta.rows(np.arange(ca["num"])); rb, fb = tb.rows(np.arange(cb["num"]))
    np.testing.assert_array_equal(ra, rb); np.testing.assert_array_equal(fa, fb)
    np.testing.assert_array_equal(ta.get("theta"), tb.get("theta"))
    comm.close()


def test_tuple_block_overflow_is_reported(assets):
    """a full tuple block refuses rows: the readers say so (TRL_E_TUPLE_OVERFLOW) and trl_tuples_dropped counts them"""
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 4096                                   # tuple_cap = max(4096, 2 n): every env contributes ~1 tuple per 0.45 s cycle
    sc = trl.ScenarioExpMACE(pack, n)
    sc.EnableExplore(True, 0.2, 0.025, 0.002)
    for _ in range(120):                       # 4 s without a hand-over: ~8 cycles per env > capacity
        sc.Update()
    L = sc.L
    cnt = C.c_int(0)
    rc = L.trl_num_tuples(sc.h, C.byref(cnt))
    assert rc == 2 and cnt.value == 8192 and "overflow" in L.trl_last_error().decode()
    with pytest.raises(RuntimeError, match="overflow"):
        sc.GetTuples()
    sc.ResetTupleBuffer()
    d = C.c_int64(0)
    assert L.trl_tuples_dropped(sc.h, C.byref(d)) == 0 and d.value > 0
    assert sc.GetNumTuples() == 0
    sc.Update()
    rows, flags, env = sc.GetTuples()          # back to normal
    assert len(rows) < 8192


def test_asynchronous_trainer_is_reproducible(assets):
    """trl_trainer_set_async: hand-over + training overlap the next update on their own stream; two runs give bit-identical weights
    (every dependency between the streams is an event, nothing is timing dependent), all tuples arrive, iterations are counted"""
    import deepterrainrl_b200 as trl
    from deepterrainrl_b200 import parallel
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    sp = np.array([0.5, 0.2, 20.0, 0.025, 0.3, 0.002, 200.0, 200.0, 0.0])
    out = []
    for run in range(2):
        sc = trl.ScenarioExpMACE(pack, 512, rng_seed=31)
        tr = trl.MACETrainer(sc, replay_mem_size=8192, num_init_samples=256, freeze_target_iters=5, seed=9)
        comm = parallel.Comm(sc, 0, 1, backend="nccl", unique_id=_uid(sc.L))
        L = sc.L
        L.trl_train_run_timed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        assert L.trl_trainer_set_async(tr.h, 1) == 0, L.trl_last_error().decode()
        state, ms = C.c_int64(0), C.c_double(0)
        assert L.trl_train_run_timed(tr.h, sp.ctypes.data_as(C.c_void_p), 90, 2, 256, C.c_double(1.0 / 30.0), 0, C.byref(state), C.byref(ms)) == 0, \
            L.trl_last_error().decode()
        c = tr.counters()
        assert state.value == 180 and 100 < c["iter"] <= 180 and c["total"] > 2000 and sc.GetNumTuples() == 0
        out.append((tr.get("theta"), c))
        comm.close(); tr.close(); sc.close()
    assert out[0][1] == out[1][1]
    np.testing.assert_array_equal(out[0][0], out[1][0])
