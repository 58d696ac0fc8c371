"""The drop-in seam, compiled and run: the reference's OWN training scenario (cScenarioTrain + cScenarioTrainMACE, with the
compiled cNeuralNetLearner / cMACETrainer / cNeuralNetTrainer behind it; oracle/_ref/libref_ctrl.so) drives a batch of environments
through include/terrainrl_b200_adapter.h -- the C++ binding of INTEGRATION.md, compiled against the reference's headers --
and the C ABI.  Only cScenarioTrain::BuildExpScene is overridden (it returns the adapter); BuildScenePool, SetupLearner,
UpdateExpScene, the annealing schedule and the trainer run as compiled.

There is no GPU here, so the C ABI underneath is the SIMT-emulator build of the very same CUDA sources (tests/simt/); on a GPU
box the adapter links against libterrainrl_b200.so instead.  The compiled trainer's network is the stand-in of the pinning
harness (Caffe is absent): its operations are answered by an oracle network object, and on every cNeuralNetLearner::SyncNet
the test hands that object's weights to the adapter's handle (PushWeights' job in a deployment).

Checked against a second handle driven directly through the Python mirror + the oracle trainer + the product's schedule
function, i.e. the product's own training loop, fed the same way: the compiled reference loop and the product loop stay
bit-identical -- trainer weights after every call, tuple counts, iteration counts, exploration rates, and the state of every
environment at the end (which also proves that rates, curriculum phase and weights reached the batch through the adapter)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "simt"))
REF_CTRL = os.path.join(HERE, "..", "oracle", "_ref", "libref_ctrl.so")
ARG_FILE = "args/opt_args_train_mace.txt"

pytestmark = pytest.mark.skipif(not (os.path.exists(REF_CTRL) and os.path.isdir("/root/reference/args"))
                                or __import__("platform").machine() != "x86_64",
                                reason="oracle/_ref/libref_ctrl.so or the reference arg / data files absent (or not x86-64: emulator)")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _blob_sizes(n_char, n_frags, frag):
    sizes = [16 * 8, 16, 32 * 16 * 4, 32, 32 * 32 * 4, 32, 64 * 5984, 64, 256 * (64 + n_char), 256]
    for out in [n_frags] + [frag] * n_frags:
        sizes += [128 * 256, 128, out * 128, out]
    return sizes


def _load_abi_and_reference():
    """The C ABI the adapter binds (the emulator build, its symbols visible to what is loaded next) and a private image of the
    reference library: its weak references to the C ABI are bound when it is loaded, and another test of this process may have
    loaded the shared one before any exporter of trl_* existed."""
    import shutil
    import tempfile
    import build as simt_build
    from loader import open_simt
    C.CDLL(simt_build.build(), mode=C.RTLD_GLOBAL)
    L = open_simt()
    tmpdir = tempfile.mkdtemp(prefix="ref_adapter_")
    ref = C.CDLL(shutil.copy(REF_CTRL, os.path.join(tmpdir, "libref_ctrl_adapter.so")))
    shutil.rmtree(tmpdir, ignore_errors=True)
    return L, ref


def test_policy_evaluation_calls_of_the_reference_through_the_adapter(assets):
    """cScenarioPoliEvalBatchedT under a cScenarioPoliEval pointer: the calls cOptScenarioPoliEval makes on a pooled scene
    (ParseArgs, Init, SetRandSeed, Reset, Update, GetNumCycles / GetNumEpisodes / GetAvgDist / ResetAvgDist / GetDistLog)
    against the same batch driven through the Python mirror."""
    from deepterrainrl_b200 import scenario
    import deepterrainrl_b200 as trl
    L, ref = _load_abi_and_reference()
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n, seed = 6, 500
    ref.ref_beval_create.restype = C.c_void_p
    ref.ref_beval_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_ulonglong, C.c_ulong]
    for f in ("ref_beval_destroy", "ref_beval_reset_avg_dist"):
        getattr(ref, f).argtypes = [C.c_void_p]
    ref.ref_beval_update.argtypes = [C.c_void_p, C.c_double]
    ref.ref_beval_handle.restype = C.c_void_p
    ref.ref_beval_handle.argtypes = [C.c_void_p]
    ref.ref_beval_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    ref.ref_beval_dist_log.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    saved = scenario._LIB
    scenario._LIB = L
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        h = ref.ref_beval_create(b"args/dog_slopes_mixed_args.txt", pack.encode(), n, 1234, seed)
    finally:
        os.chdir(cwd)
    assert h
    h = C.c_void_p(h)
    try:
        g = trl.ScenarioPoliEval(pack, n, terrain_seeds=np.arange(seed, seed + n, dtype=np.uint64))
        g.Reset()

        def ref_stats():
            c, e, a = C.c_long(0), C.c_long(0), C.c_double(0)
            ref.ref_beval_stats(h, C.byref(c), C.byref(e), C.byref(a))
            return c.value, e.value, a.value

        class View(trl.ScenarioPoliEval):             # the adapter's handle through the Python mirror (no ownership)
            def __init__(self, hh):
                self.L, self.h, self.num_dof, self.num_joints, self.num_envs = L, C.c_void_p(hh), g.num_dof, g.num_joints, n

            def close(self):
                self.h = None
        v = View(ref.ref_beval_handle(h))
        for k in range(30):
            ref.ref_beval_update(h, 1.0 / 30.0)
            g.Update(1.0 / 30.0)
            if k == 24:        # tip env 2 over on both sides: the next update ends its episode (fall -> distance record -> reset)
                for sc in (v, g):
                    q, qd, _, _ = sc.GetState(2)
                    q[2] = 3.0
                    sc.SetState(2, q=q, qd=qd)
        s = g._stats()
        assert ref_stats() == (s["cycles"], s["episodes"], s["avg_dist"])
        assert s["episodes"] >= 1 and s["cycles"] > 0 and s["avg_dist"] != 0.0
        log = np.zeros(64)
        nl = ref.ref_beval_dist_log(h, _p(log), 64)
        gl = g.GetDistLog()
        glog = gl[0] if isinstance(gl, tuple) else gl
        assert nl == len(glog) >= 1 and np.array_equal(log[:nl], np.asarray(glog))
        ref.ref_beval_reset_avg_dist(h)                 # cScenarioPoliEval::ResetAvgDist: episodes and mean back to 0, cycles kept
        g.ResetAvgDist()
        s2 = g._stats()
        assert ref_stats() == (s["cycles"], 0, 0.0) == (s2["cycles"], s2["episodes"], s2["avg_dist"])
        q1, _ = v.GetStateAll()
        q2, _ = g.GetStateAll()
        assert np.array_equal(q1, q2)
        v.close()
        g.close()
    finally:
        ref.ref_beval_destroy(h)
        scenario._LIB = saved


def test_reference_evaluation_driver_runs_over_the_adapter(assets, tmp_path):
    """optimizer/scenarios/OptScenarioPoliEval.cpp compiled as it is -- Run, its worker thread, EvalHelper (Update until the
    episode / cycle budget is spent, UpdateRecord, ResetAvgDist), OutputResults -- over the batched adapter; BuildScenePool is the
    reference's own lines with the scene class swapped.  Against the same protocol written out over the Python mirror."""
    from deepterrainrl_b200 import scenario
    import deepterrainrl_b200 as trl
    L, ref = _load_abi_and_reference()
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n, max_ep = 6, 1
    out_file = str(tmp_path / "poli_eval.txt")
    extra = ["-poli_eval_max_episodes=", str(max_ep), "-poli_eval_max_cycles=", "100000", "-poli_eval_rand_seed=", "31",
             "-output_path=", out_file]
    extra = [e.encode() for e in extra]
    arr = (C.c_char_p * len(extra))(*extra)
    ref.ref_opteval_create.restype = C.c_void_p
    ref.ref_opteval_create.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_ulonglong]
    ref.ref_opteval_handle.restype = C.c_void_p
    for f in ("ref_opteval_run", "ref_opteval_destroy", "ref_opteval_handle"):
        getattr(ref, f).argtypes = [C.c_void_p]
    ref.ref_opteval_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    saved = scenario._LIB
    scenario._LIB = L
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        h = ref.ref_opteval_create(b"args/dog_slopes_mixed_args.txt", arr, len(extra), pack.encode(), n, 1234)
    finally:
        os.chdir(cwd)
    assert h
    h = C.c_void_p(h)
    try:
        counts, avg, seed0 = (C.c_long * 2)(), C.c_double(0), C.c_ulong(0)
        ref.ref_opteval_results(h, counts, C.byref(avg), C.byref(seed0))
        # the product-side batch, seeded as the compiled BuildScenePool seeded the adapter's (cRand(31) -> first RandInt)
        g = trl.ScenarioPoliEval(pack, n, terrain_seeds=np.arange(seed0.value, seed0.value + n, dtype=np.uint64))
        g.Reset()

        class View(trl.ScenarioPoliEval):
            def __init__(self, hh):
                self.L, self.h, self.num_dof, self.num_joints, self.num_envs = L, C.c_void_p(hh), g.num_dof, g.num_joints, n

            def close(self):
                self.h = None
        v = View(ref.ref_opteval_handle(h))
        q1, _ = v.GetStateAll()
        q2, _ = g.GetStateAll()
        assert np.array_equal(q1, q2)
        for sc in (v, g):                               # run both batches past their first cycles, then tip one environment over
            for k in range(24):
                sc.Update(1.0 / 30.0)
            q, qd, _, _ = sc.GetState(1)
            q[2] = 3.0
            sc.SetState(1, q=q, qd=qd)
        ref.ref_opteval_run(h)                          # cOptScenarioPoliEval::Run: thread -> EvalHelper -> OutputResults
        # EvalHelper over the Python mirror (optimizer/scenarios/OptScenarioPoliEval.cpp:170-198)
        rec = dict(avg=0.0, ep=0, cyc=0)
        num_episodes, num_cycles, prev = 0, 0, 0
        while num_episodes < max_ep and num_cycles < 100000:
            g.Update(1.0 / 30.0)
            num_cycles = g.GetNumCycles()
            cur = g.GetNumEpisodes()
            if cur >= 10 or cur + num_episodes >= max_ep:
                a = g.GetAvgDist()
                rec["avg"] = (rec["ep"] * rec["avg"] + cur * a) / (rec["ep"] + cur)        # cMathUtil::AddAverage
                rec["ep"] += cur
                rec["cyc"] += num_cycles - prev
                g.ResetAvgDist()
                num_episodes += cur
                prev = num_cycles
        ref.ref_opteval_results(h, counts, C.byref(avg), C.byref(seed0))
        assert (counts[0], counts[1]) == (rec["ep"], rec["cyc"]) and rec["ep"] >= 1 and rec["cyc"] > 0
        assert avg.value == rec["avg"] != 0.0
        dist = g.GetDistLog()[0]
        assert open(out_file).read() == ", ".join("%f" % d for d in dist) + "\n"            # OutputResults: std::to_string per entry
        q1, _ = v.GetStateAll()
        q2, _ = g.GetStateAll()
        assert np.array_equal(q1, q2)
        v.close()
        g.close()
    finally:
        ref.ref_opteval_destroy(h)
        scenario._LIB = saved


def test_reference_training_scenario_drives_the_batch_through_the_adapter(assets):
    from pyoracle import Oracle, OracleTrainer
    from deepterrainrl_b200 import scenario
    from deepterrainrl_b200.train import TrainSchedule
    import deepterrainrl_b200 as trl

    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    n_envs, B, TB, rseed, rng_seed = 8, 8, 4, 4321, 77
    L, ref = _load_abi_and_reference()

    kw = dict(replay_cap=400, num_init_samples=8, num_steps_per_iter=1, freeze_target_iters=3, init_input_offset_scale=1, seed=1)
    eng = OracleTrainer(pack, **kw)         # the network under the compiled reference trainer
    orc = OracleTrainer(pack, **kw)         # the trainer of the product-side loop
    dummy = Oracle(pack, 1, 1)              # never stepped: carries the restated cRand the oracle trainer samples from
    OL = eng.L
    for f, a in (("orc_use_ref_rand", [C.c_void_p, C.c_ulong]), ("orc_trainer_share_rand", [C.c_void_p, C.c_void_p]),
                 ("orc_trainer_set_batch", [C.c_void_p, C.c_int])):
        getattr(OL, f).argtypes = a
    OL.orc_trainer_solver_step.restype = C.c_double
    OL.orc_trainer_set_batch(eng.h, B)
    OL.orc_trainer_set_batch(orc.h, B)
    OL.orc_use_ref_rand(dummy.h, rseed)
    OL.orc_trainer_share_rand(orc.h, dummy.h)
    S, no = orc.n_in, orc.n_out
    SCHED = dict(init_exp_rate=0.5, exp_rate=0.2, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.3, exp_base_rate=0.002,
                 trainer_num_anneal_iters=12, exp_base_anneal_iters=8, trainer_curriculum_iters=10)
    sched = TrainSchedule(**SCHED)

    saved = scenario._LIB
    scenario._LIB = L
    st = dict(err=None, adapter=None, h=None)
    log = dict(copies=[], trains=0, evals=[0, 0, 0])
    try:
        # ---- the product-side loop's environment batch (same pack, seeds, call sequence as BuildScenePool makes through the adapter)
        g2 = trl.ScenarioExpMACE(pack, n_envs, rng_seed=rng_seed)
        sizes = _blob_sizes(S - 200, g2.num_frags, g2.frag_size)
        assert sum(sizes) == orc.num_params

        def weights_of(tr):
            theta = tr.get("theta")
            blobs, o = [], 0
            for n in sizes:
                blobs.append(theta[o:o + n]); o += n
            return blobs, tr.get("in_off"), tr.get("in_scale"), tr.get("out_off"), tr.get("out_scale")

        class Adapter(trl.ScenarioExpMACE):          # a Python view of the handle the C++ adapter owns (no ownership)
            def __init__(self, h):
                self.L = L
                self.h = C.c_void_p(h)
                v = [C.c_int(0) for _ in range(7)]
                self._ck(L.trl_sizes(self.h, *[C.byref(x) for x in v]))
                self.num_envs, self.state_size, self.action_size, self.num_frags, self.frag_size, self.num_dof, self.num_joints = \
                    [x.value for x in v]

            def close(self):
                self.h = None

        DP = C.POINTER(C.c_double)
        EV = C.CFUNCTYPE(None, C.c_int, DP, C.c_int, DP, C.c_void_p)
        TR = C.CFUNCTYPE(None, C.c_int, DP, DP, C.c_int, C.c_void_p)
        CP = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_void_p)
        CO = C.CFUNCTYPE(None, DP, C.c_int, DP, DP, C.c_void_p)
        SO = C.CFUNCTYPE(None, C.c_int, DP, DP, C.c_void_p)

        def guarded(fn):
            def w(*a):
                if st["err"] is not None:
                    return
                try:
                    fn(*a)
                except BaseException as e:       # cannot cross the C frames
                    st["err"] = e
            return w

        # cNeuralNet instances in construction order: 0 the controller's (= the learner's) network, 1 the trainer's, 2 its target
        def ev(net, X, Bn, Y, u):
            assert net in (1, 2), "the batch evaluates its own network: the controller's net is never asked"
            x = np.ctypeslib.as_array(X, (Bn, S)).copy()
            y = np.zeros((Bn, no))
            OL.orc_trainer_eval_batch(eng.h, 1 if net == 2 else 0, _p(x), Bn, _p(y))
            np.ctypeslib.as_array(Y, (Bn, no))[:] = y
            log["evals"][net] += 1

        def tr(net, X, Y, Bn, u):
            assert net == 1 and Bn == B
            x = np.ctypeslib.as_array(X, (Bn, S)).copy()
            y = np.ctypeslib.as_array(Y, (Bn, no)).copy()
            OL.orc_trainer_solver_step(eng.h, _p(x), _p(y))
            log["trains"] += 1

        pending_sync = [False]

        def push_weights(trainer):               # cScenarioExpBatched::PushWeights through the compiled adapter
            blobs, in_off, in_scale, out_off, out_scale = weights_of(trainer)
            blobs = [np.ascontiguousarray(b) for b in blobs]
            ptrs = (C.c_void_p * len(blobs))(*[b.ctypes.data for b in blobs])
            counts = np.array([b.size for b in blobs], np.int64)
            ref.ref_btrain_push_weights(st["h"], ptrs, _p(counts), len(blobs), _p(in_off), _p(in_scale), _p(out_off), _p(out_scale))

        def cp(dst, src, u):
            log["copies"].append((dst, src))
            if (dst, src) == (2, 1):
                OL.orc_trainer_copy_to_target(eng.h)
            elif (dst, src) == (0, 1):           # cNeuralNetLearner::SyncNet: hand the trained weights to the batch
                if st["adapter"] is not None:
                    push_weights(eng)
                else:
                    pending_sync[0] = True       # the first SyncNet happens inside Init, before the handle is reachable from here
            else:
                raise AssertionError((dst, src))

        def co(X, n, off, sc, u):
            x = np.ctypeslib.as_array(X, (n, S)).copy()
            of, s = np.zeros(S), np.zeros(S)
            OL.orc_calc_offset_scale(_p(x), n, S, _p(of), _p(s))
            np.ctypeslib.as_array(off, (S,))[:] = of
            np.ctypeslib.as_array(sc, (S,))[:] = s

        def so(net, off, sc, u):
            of = np.ctypeslib.as_array(off, (S,)).copy()
            s = np.ctypeslib.as_array(sc, (S,)).copy()
            OL.orc_trainer_set_input_offset_scale(eng.h, 1 if net == 2 else 0, _p(of), _p(s))

        cbs = (EV(guarded(ev)), TR(guarded(tr)), CP(guarded(cp)), CO(guarded(co)), SO(guarded(so)))
        ref.ref_ctrl_set_net_output(S, _p(np.zeros(no)), _p(np.ascontiguousarray(orc.get("out_scale"))), no)
        extra = ["-init_exp_rate=", "0.5", "-init_exp_base_rate=", "0.3", "-terrain_file=", "data/terrain/slopes_mixed.txt",
                 "-tuple_buffer_size=", str(TB), "-trainer_replay_mem_size=", "400", "-trainer_num_init_samples=", "8",
                 "-trainer_freeze_target_iters=", "3", "-trainer_num_anneal_iters=", "12", "-exp_base_anneal_iters=", "8",
                 "-trainer_curriculum_iters=", "10", "-trainer_int_iter=", "0", "-trainer_iters_per_output=", "100000",
                 "-output_path=", "/tmp/ref_btrain_model.h5"]
        extra = [e.encode() for e in extra]
        arr = (C.c_char_p * len(extra))(*extra)
        dims = np.array([S, no, B], np.int32)
        ref.ref_btrain_create.restype = C.c_void_p
        ref.ref_btrain_create.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_ulonglong, C.c_ulong, C.c_void_p,
                                          EV, TR, CP, CO, SO, C.c_void_p]
        ref.ref_btrain_update.argtypes = [C.c_void_p, C.c_double]
        ref.ref_btrain_handle.restype = C.c_void_p
        ref.ref_btrain_handle.argtypes = [C.c_void_p]
        ref.ref_btrain_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        ref.ref_btrain_destroy.argtypes = [C.c_void_p]
        ref.ref_btrain_push_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4
        cwd = os.getcwd()
        os.chdir("/root/reference")
        try:
            h = ref.ref_btrain_create(ARG_FILE.encode(), arr, len(extra), pack.encode(), n_envs, rng_seed, rseed, _p(dims), *cbs, None)
        finally:
            os.chdir(cwd)
        assert h and st["err"] is None, st["err"]
        h = C.c_void_p(h)
        st["h"] = h
        # Init of the (never stepped) reference scene behind the adapter drew from cMathUtil's engine (CommandRandAction); from here
        # on only the compiled trainer's minibatch sampling does -- start it where the oracle trainer's restated cRand starts
        ref.ref_btrain_reseed.argtypes = [C.c_ulong]
        ref.ref_btrain_reseed(rseed)
        try:
            assert log["copies"] == [(2, 1), (0, 1)]       # cMACETrainer::BuildNetPool (target), cNeuralNetLearner::Init (SyncNet)
            ad = Adapter(ref.ref_btrain_handle(h))
            st["adapter"] = ad
            assert (ad.num_envs, ad.state_size, ad.action_size) == (n_envs, g2.state_size, g2.action_size)
            if pending_sync[0]:
                push_weights(eng)
            # the product-side loop, brought to the state BuildScenePool leaves the adapter's batch in
            g2.SetWeights(*weights_of(orc))
            g2.EnableExplore(1, SCHED["init_exp_rate"], SCHED["init_exp_temp"], SCHED["init_exp_base_rate"])
            g2.SetTerrainParamsLerp(1.0)                   # gInitCurriculumPhase (scenarios/ScenarioTrain.cpp:6,213)
            g2.Reset()
            q1, qd1 = ad.GetStateAll()
            q2, qd2 = g2.GetStateAll()
            assert np.array_equal(q1, q2) and np.array_equal(qd1, qd2)

            DT = 1.0 / 30.0
            n_calls, flag_bits = 0, 0
            for k in range(400):
                ref.ref_btrain_update(h, DT)               # cScenarioTrain::Update -> UpdateExpScene -> adapter -> C ABI
                if st["err"] is not None:
                    raise st["err"]
                g2.Update(DT)
                if g2.GetNumTuples() >= TB:                # the product's side of cScenarioTrain::UpdateExpScene
                    rows, flags, _ = g2.GetTuples(f64=True)
                    flag_bits |= int(np.bitwise_or.reduce(flags))
                    orc.add_tuples(rows, flags)
                    orc.train()
                    g2.SetWeights(*weights_of(orc))
                    it = orc.counters()["iter"]
                    s = sched(it)
                    g2.EnableExplore(1, s["exp_rate"], s["exp_temp"], s["exp_base_rate"])
                    g2.SetTerrainParamsLerp(s["curriculum_phase"])
                    g2.ResetTupleBuffer()
                    n_calls += 1
                    cnt, rates = (C.c_long * 2)(), (C.c_double * 3)()
                    ref.ref_btrain_status(h, cnt, rates)
                    assert (cnt[0], cnt[1]) == (it, orc.counters()["total"]), (k, list(cnt), it)
                    assert np.allclose(list(rates), [s["exp_rate"], s["exp_temp"], s["exp_base_rate"]], rtol=1e-15, atol=0)
                    assert np.array_equal(eng.get("theta"), orc.get("theta")), (k, it)
                    assert ad.GetNumTuples() == 0           # the compiled loop has emptied the adapter's buffer too
                    if it >= 3:
                        break
            assert n_calls >= 3 and orc.counters()["iter"] >= 3 and log["trains"] >= 3, (n_calls, orc.counters(), log)
            assert flag_bits & 6, "no exploration flag ever set: the rates did not reach the batch"
            q1, qd1 = ad.GetStateAll()
            q2, qd2 = g2.GetStateAll()
            assert np.array_equal(q1, q2) and np.array_equal(qd1, qd2)
            assert not np.array_equal(eng.get("theta"), OracleTrainer(pack, **kw).get("theta"))    # training did move the weights
            print(f"compiled cScenarioTrainMACE over the adapter: {k + 1} updates x {n_envs} envs, {n_calls} trainer calls, "
                  f"{orc.counters()['iter']} iterations, weights and all environment states bit-identical to the product loop")
        finally:
            st["adapter"] = None
            ref.ref_btrain_destroy(h)
            g2.close()
    finally:
        scenario._LIB = saved
