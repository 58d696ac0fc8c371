"""GPU MACE trainer (csrc/trl_train.cu) against the CPU restatement (oracle/trainer.h): same tuple stream, same sampling RNG ->
same replay buffers (exact) and the same weights after several critic + actor solver steps (f64, different summation order)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _synthetic_tuples(n, S, A, in_off, in_scale, seed, actor_frac=0.45):
    rng = np.random.default_rng(seed)
    W = 1 + 2 * S + A
    rows = np.zeros((n, W))
    rows[:, 0] = rng.uniform(0, 8, n)                                 # large rewards: many positive temporal differences
    for k in (1, 1 + S + A):
        rows[:, k:k + S] = rng.normal(size=(n, S)) / np.where(in_scale == 0, 1.0, in_scale) - in_off
    rows[:, 1 + S] = rng.integers(0, 3, n)
    rows[:, 2 + S:1 + S + A] = rng.normal(size=(n, A - 1)) * 0.2
    flags = np.where(rng.uniform(size=n) < actor_frac, 4, 0).astype(np.uint32)
    flags |= (rng.uniform(size=n) < 0.1).astype(np.uint32)            # failures
    return rows, flags


@pytest.mark.parametrize("scene", ["dog_slopes_mixed", "raptor_narrow_gaps"])
def test_trainer_matches_oracle(assets, scene):
    from pyoracle import OracleTrainer
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, scene + ".trlpack")
    kw = dict(num_init_samples=96, num_steps_per_iter=1, freeze_target_iters=3, init_input_offset_scale=1, seed=21)
    sc = trl.ScenarioExpMACE(pack, 64)
    g = trl.MACETrainer(sc, replay_mem_size=160, **kw)
    o = OracleTrainer(pack, replay_cap=160, **kw)
    assert g.num_params == o.num_params == (570474 if scene.startswith("dog") else 568039)
    np.testing.assert_array_equal(g.get("theta"), o.get("theta"))
    in_off, in_scale = o.get("in_off"), o.get("in_scale")
    rows, flags = _synthetic_tuples(128, g.S, g.A, in_off, in_scale, 3)
    rows[17, 40] = np.inf                                              # CheckTuple must drop it on both sides
    g.AddTuples(rows, flags); o.add_tuples(rows, flags)
    cg, co = g.counters(), o.counters()
    for k in ("num", "head", "total", "critic", "actor", "stage"):
        assert cg[k] == co[k], k
    np.testing.assert_array_equal(g.lists("critic"), o.lists("critic"))
    np.testing.assert_array_equal(g.lists("actor"), o.lists("actor"))
    for it in range(8):
        g.Train(1); o.train()
        if it == 2:                                                    # wrap the ring: overwritten slots change buffers
            r2, f2 = _synthetic_tuples(64, g.S, g.A, in_off, in_scale, 9)
            g.AddTuples(r2, f2); o.add_tuples(r2, f2)
        cg, co = g.counters(), o.counters()
        for k in ("iter", "actor_iter", "stage", "num", "head", "total", "critic", "actor", "actor_batch"):
            assert cg[k] == co[k], (it, k, cg, co)
        np.testing.assert_array_equal(g.lists("critic"), o.lists("critic"))
        np.testing.assert_array_equal(g.lists("actor"), o.lists("actor"))
        np.testing.assert_array_equal(g.lists("actor_batch"), o.lists("actor_batch"))
        lo = o.losses()
        assert abs(cg["critic_loss"] - lo[0]) <= 1e-10 * max(1.0, lo[0])
        tg, to = g.get("theta"), o.get("theta")
        assert np.max(np.abs(tg - to)) <= 1e-10 * max(1.0, np.max(np.abs(to))), it
    assert co["iter"] == 8 and co["actor_iter"] >= 1                   # both kinds of solver step happened
    np.testing.assert_allclose(g.get("in_off"), o.get("in_off"), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g.get("in_scale"), o.get("in_scale"), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g.get("target"), o.get("target"), rtol=0, atol=1e-10)
    np.testing.assert_allclose(g.get("history"), o.get("history"), rtol=0, atol=1e-10)
    assert not np.array_equal(g.get("target"), g.get("theta"))


def test_rollout_uses_trainer_weights_and_device_tuples(assets):
    """The scenario evaluates the trainer's net in place: zeroing the actor heads through the trainer changes the rollout;
    tuples flow from the scenario's device block into the replay memory without the host, and training runs on them."""
    import deepterrainrl_b200 as trl
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n = 512
    a = trl.ScenarioExpMACE(pack, n, rng_seed=5)
    b = trl.ScenarioExpMACE(pack, n, rng_seed=5)
    for sc in (a, b):
        sc.EnableExplore(True, 0.2, 0.025, 0.01)
    tr = trl.MACETrainer(b, replay_mem_size=20000, num_init_samples=1500, freeze_target_iters=4, seed=2)
    for _ in range(40):
        a.Update(1.0 / 30.0); b.Update(1.0 / 30.0)
    np.testing.assert_array_equal(a.GetStateAll()[0], b.GetStateAll()[0])      # binding alone changes nothing
    nt = b.GetNumTuples()
    assert nt > n
    src_rows, src_flags, src_env = b.GetTuples(f64=True)                       # does not reset the scenario's buffer
    canon = np.argsort(src_env, kind="stable")                                 # the hand-over ranks the block by env id (reproducible)
    src_rows, src_flags = src_rows[canon], src_flags[canon]
    tr.AddTuplesFromScene()
    c = tr.counters()
    assert c["num"] == nt and c["critic"] + c["actor"] == nt and c["actor"] > 0
    got_rows, got_flags = tr.rows(np.arange(nt))                               # replay slots 0..nt-1 in canonical arrival order
    np.testing.assert_array_equal(got_rows, src_rows.astype(np.float32))       # SetTuple stores floats
    np.testing.assert_array_equal(got_flags, src_flags.astype(np.int32))
    assert b.GetNumTuples() == 0                                               # ResetTupleBuffer happened on the device
    theta0 = tr.get("theta")
    for _ in range(30):
        b.Update(1.0 / 30.0)
        tr.AddTuplesFromScene()
        tr.Train(2)
    c = tr.counters()
    assert c["stage"] == 1 and c["iter"] > 4 and np.isfinite(c["critic_loss"])
    theta1 = tr.get("theta")
    assert np.all(np.isfinite(theta1)) and not np.array_equal(theta0, theta1)
    assert np.all(np.isfinite(b.GetStateAll()[0]))
    # weights owned by the trainer: trl_set_weights is refused, set_theta reaches the rollout
    with pytest.raises(RuntimeError):
        b.SetWeights([np.zeros(1)] * 26, np.zeros(1), np.zeros(1), np.zeros(1), np.zeros(1))
    tr.close()
    for _ in range(3):
        b.Update(1.0 / 30.0)                                                   # scenario stays usable with the trained weights
    assert np.all(np.isfinite(b.GetStateAll()[0]))


def test_training_driver_end_to_end(assets, tmp_path):
    """ScenarioTrainMACE: rollout -> device tuples -> trainer iterations -> annealed exploration, model written in the Caffe
    HDF5 layout equals the trainer's weights and loads back into a fresh scenario."""
    import deepterrainrl_b200 as trl
    from deepterrainrl_b200.model_io import read_model
    from deepterrainrl_b200.train import ScenarioTrainMACE, TrainSchedule
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    sched = TrainSchedule(init_exp_rate=0.9, exp_rate=0.2, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.9, exp_base_rate=0.002,
                          trainer_num_anneal_iters=40, exp_base_anneal_iters=40)
    st = ScenarioTrainMACE(pack, 1024, schedule=sched, rng_seed=3,
                           trainer_params=dict(replay_mem_size=50000, num_init_samples=800, freeze_target_iters=10, seed=4))
    st.Run(90)
    c = st.trainer.counters()
    assert c["stage"] == 1 and c["iter"] >= 5 and c["num"] > 800 and np.isfinite(c["critic_loss"])
    assert c["actor"] > 0 and c["critic"] > 0
    s = sched(st.trainer.GetIter())
    assert s["exp_rate"] < 0.9                                       # annealing follows the iteration count
    out = str(tmp_path / "model.h5")
    st.OutputModel(out)
    layers, scale = read_model(out)
    blobs = st.trainer.blobs()
    for name, (w, b) in blobs.items():
        np.testing.assert_array_equal(layers[name][0], w)
        np.testing.assert_array_equal(layers[name][1], b)
    np.testing.assert_array_equal(scale["InputOffset"], st.trainer.get("in_off"))
    # native write-back of the same policy is byte-identical, and trl_load_model reads it into a fresh scenario
    out2 = str(tmp_path / "native" / "model.h5")
    os.makedirs(os.path.dirname(out2))
    st.exp.OutputModel(out2, mtime=7)
    from deepterrainrl_b200.model_io import write_model
    write_model(str(tmp_path / "py7.h5"), blobs, st.trainer.get("in_off"), st.trainer.get("in_scale"), st.trainer.get("out_off"),
                st.trainer.get("out_scale"), mtime=7)
    assert open(out2, "rb").read() == open(tmp_path / "py7.h5", "rb").read()
    ev2 = trl.ScenarioPoliEval(pack, 64)
    ev2.LoadModel(out2)
    # the written model drives a fresh evaluation scenario (cNeuralNet::LoadModel path = SetWeights)
    ev = trl.ScenarioPoliEval(pack, 64)
    flat = []
    from deepterrainrl_b200.trainer import NET_LAYERS
    for name in NET_LAYERS:
        flat += [layers[name][0].ravel(), layers[name][1]]
    ev.SetWeights(flat, scale["InputOffset"], scale["InputScale"], scale["OutputOffset"], scale["OutputScale"])
    for _ in range(10):
        ev.Update(1.0 / 30.0); ev2.Update(1.0 / 30.0)
    assert np.all(np.isfinite(ev.GetStateAll()[0]))
    np.testing.assert_array_equal(ev.GetStateAll()[0], ev2.GetStateAll()[0])       # both loading paths give the same policy


def test_training_from_scratch(assets):
    """trl_trainer_init_fresh: xavier weights inside their bounds, zero biases, controller-derived output offset / scale; the
    rollout runs on the fresh net (exploring from the start like the reference) and the trainer leaves the init stage."""
    import deepterrainrl_b200 as trl
    from deepterrainrl_b200.train import ScenarioTrainMACE, TrainSchedule
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    # (with the reference's initial rates 0.9 / 0.9 only 1 % of the tuples reach the critic buffer; it then relies on 50000 initial samples)
    sched = TrainSchedule(init_exp_rate=0.5, exp_rate=0.2, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.3, exp_base_rate=0.002,
                          trainer_num_anneal_iters=50000, exp_base_anneal_iters=50000)
    st = ScenarioTrainMACE(pack, 1024, schedule=sched, rng_seed=8, iters_per_update=2,
                           trainer_params=dict(replay_mem_size=50000, num_init_samples=600, freeze_target_iters=20, seed=6))
    shipped = st.trainer.get("theta")
    st.trainer.InitFresh(seed=3)
    blobs = st.trainer.blobs()
    for name, (w, b) in blobs.items():
        fan_in = w.size / w.shape[0]
        s = np.sqrt(3.0 / fan_in)
        assert np.all(np.abs(w) <= s) and np.abs(w).max() > 0.9 * s and abs(w.mean()) < 0.05 * s
        assert np.all(b == 0)
    assert not np.array_equal(st.trainer.get("theta"), shipped)
    np.testing.assert_array_equal(st.trainer.get("target"), st.trainer.get("theta"))
    oo, os_ = st.trainer.get("out_off"), st.trainer.get("out_scale")
    go, gs = st.exp.GetOutputOffsetScale()
    np.testing.assert_array_equal(oo, go); np.testing.assert_array_equal(os_, gs)
    assert np.all(oo[:3] == -0.5) and np.all(os_[:3] == 2.0)
    assert np.all(np.isfinite(os_)) and np.all(os_[3:] > 0)
    np.testing.assert_array_equal(os_[3:32], os_[32:61])               # every actor shares the action-library scale
    assert np.all(st.trainer.get("in_off") == 0) and np.all(st.trainer.get("in_scale") == 1)
    st.Run(70)
    c = st.trainer.counters()
    assert c["stage"] == 1 and c["iter"] > 10 and np.isfinite(c["critic_loss"])
    assert np.all(np.isfinite(st.trainer.get("theta"))) and np.all(np.isfinite(st.exp.GetStateAll()[0]))
    assert not np.all(st.trainer.get("in_off") == 0)                   # refitted from the replay memory at the stage switch


def test_native_training_loop_equals_python_loop(assets):
    """trl_train_run (cScenarioTrain::Run behind the C ABI) and the Python ScenarioTrainMACE.Run drive the same sequence of calls:
    identical weights, counters and env states."""
    from deepterrainrl_b200.train import ScenarioTrainMACE, TrainSchedule
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    mk = lambda: ScenarioTrainMACE(pack, 512, rng_seed=11, iters_per_update=2,
                                   schedule=TrainSchedule(init_exp_rate=0.5, exp_rate=0.2, init_exp_temp=5, exp_temp=0.025,
                                                          init_exp_base_rate=0.2, exp_base_rate=0.002, trainer_num_anneal_iters=60,
                                                          exp_base_anneal_iters=60),
                                   trainer_params=dict(replay_mem_size=20000, num_init_samples=300, freeze_target_iters=7, seed=2))
    a, b = mk(), mk()
    a.Run(50)
    b.RunNative(50)
    ca, cb = a.trainer.counters(), b.trainer.counters()
    assert ca == cb and ca["iter"] > 5
    np.testing.assert_array_equal(a.trainer.get("theta"), b.trainer.get("theta"))
    np.testing.assert_array_equal(a.exp.GetStateAll()[0], b.exp.GetStateAll()[0])
