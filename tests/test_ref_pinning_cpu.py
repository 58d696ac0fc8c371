"""Pins the oracle against the REFERENCE's own code where that code can be compiled.  `make -C oracle ref` builds, unmodified from
/root/reference and against stand-ins for the absent Eigen / jsoncpp headers (oracle/ref_shim):
  oracle/_ref/libref_terrain.so  sim/TerrainGen2D.cpp, util/Rand.cpp, util/ArgParser.cpp, util/FileUtil.cpp
  oracle/_ref/libref_rbd.so      anim/KinTree.cpp, sim/SpAlg.cpp, sim/RBDModel.cpp, sim/RBDUtil.cpp, util/MathUtil.cpp, util/JsonUtil.cpp
Checked: the cRand streams, all 14 terrain generators, the default parameter table, the type names and the arg-file semantics
(bit for bit); and, at random poses of the dog, goat and raptor read by the reference's own loaders from the character files,
the mass matrix, the bias force (with the reference's BuildCjPlanar), the gravity force, the Jacobian, the centre of mass and
the joint positions of the controller's rigid-body model (oracle/rbd.h).  Skipped where the libraries have not been built."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "..", "oracle", "_ref", "libref_terrain.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (reference sources absent)")

TYPES = ["flat", "gaps", "steps", "walls", "bumps", "mixed", "narrow_gaps", "slopes", "slopes_gaps", "slopes_walls", "slopes_steps",
         "slopes_mixed", "slopes_narrow_gaps", "cliffs"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def libs():
    from pyoracle import lib
    ref = C.CDLL(REF_LIB)
    orc = lib()
    for L, pre in ((ref, "ref"), (orc, "orc")):
        f = getattr(L, pre + "_terrain_build")
        f.argtypes = [C.c_int, C.c_void_p, C.c_ulong, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
        g = getattr(L, pre + "_terrain_build_after_flat")
        g.argtypes = [C.c_int, C.c_void_p, C.c_ulong, C.c_double, C.c_double, C.c_void_p, C.c_int]
        r = getattr(L, pre + "_rand_stream")
        r.argtypes = [C.c_ulong, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p]
        r.restype = None
    return ref, orc


def _params(ref, rng=None):
    p = np.zeros(40)
    ref.ref_terrain_default_params(_p(p))
    if rng is not None:                                   # perturbed but ordered (min <= max) parameter sets
        q = p * rng.uniform(0.6, 1.4, 40)
        for i in range(0, 36, 2):
            lo, hi = sorted((q[i], q[i + 1]))
            q[i], q[i + 1] = lo, hi
        q[28:30] = np.round(np.clip(q[28:30], 1, 4))      # narrow-gap counts
        q[36] = np.round(q[36])                           # cliff mini count
        q[38:40] = sorted(q[38:40])
        p = q
    return p


def test_rand_streams_match_reference(libs):
    ref, orc = libs
    for seed in (1, 7, 123456789, 2 ** 31 + 5):
        for kind, a, b in ((0, 0, 0), (1, -2.5, 7.25), (1, 3.0, 3.0), (2, 0, 0), (3, 0, 10), (3, -4, 9), (3, 5, 5), (4, 0, 0), (5, 0, 0)):
            x = np.zeros(257); y = np.zeros(257)
            ref.ref_rand_stream(seed, kind, a, b, 257, _p(x))
            orc.orc_rand_stream(seed, kind, a, b, 257, _p(y))
            np.testing.assert_array_equal(x, y)


def test_terrain_generators_match_reference_bit_for_bit(libs):
    ref, orc = libs
    rng = np.random.default_rng(0)
    cap = 4096
    checked = 0
    for t, name in enumerate(TYPES):
        assert ref.ref_terrain_parse_type(name.encode()) == t
        for trial in range(6):
            p = _params(ref, rng if trial else None)
            seed = int(rng.integers(1, 2 ** 31))
            width = float(rng.uniform(3.0, 40.0))
            a = np.zeros(cap, np.float32); b = np.zeros(cap, np.float32)
            wa = C.c_double(0); wb = C.c_double(0)
            na = ref.ref_terrain_build(t, _p(p), seed, width, _p(a), cap, C.byref(wa))
            nb = orc.orc_terrain_build(t, _p(p), seed, width, _p(b), cap, C.byref(wb))
            assert na == nb and 0 < na <= cap, (name, trial)
            assert wa.value == wb.value
            np.testing.assert_array_equal(a[:na].view(np.uint32), b[:nb].view(np.uint32))
            flat_w = float(rng.uniform(0.5, 3.0))
            na = ref.ref_terrain_build_after_flat(t, _p(p), seed, flat_w, width, _p(a), cap)
            nb = orc.orc_terrain_build_after_flat(t, _p(p), seed, flat_w, width, _p(b), cap)
            assert na == nb
            np.testing.assert_array_equal(a[:na].view(np.uint32), b[:nb].view(np.uint32))
            checked += 2
    assert checked == 14 * 12


def test_default_params_match_the_packs(libs, assets):
    from pack_scene import read_pack
    ref, _ = libs
    p = _params(ref)
    for pack in glob.glob(os.path.join(assets, "*.trlpack")):
        np.testing.assert_array_equal(read_pack(pack)["terrain_default_params"], p)


@pytest.mark.skipif(not os.path.isdir("/root/reference/args"), reason="reference arg files absent")
def test_arg_file_semantics_match_reference_parser(libs):
    """cArgParser compiled from the reference vs the product's readers (train.parse_arg_file; tools/pack_scene tokenizer) on every
    shipped arg file: same token count, same value for every key the product reads."""
    from deepterrainrl_b200.train import parse_arg_file
    from pack_scene import tokenize_arg_file
    ref, _ = libs
    ref.ref_args_string.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    files = sorted(glob.glob("/root/reference/args/*.txt"))
    assert len(files) >= 10
    n_keys = 0
    for f in files:
        assert ref.ref_args_count(f.encode()) == len(tokenize_arg_file(f))
        ours = parse_arg_file(f)
        for key, val in ours.items():
            buf = C.create_string_buffer(1024)
            ok = ref.ref_args_string(f.encode(), key.encode(), buf, 1024)
            if val == "":
                assert not ok                                   # a key followed by another key parses as "absent"
                continue
            assert ok and buf.value.decode() == val.split()[0], (f, key)
            n_keys += 1
            try:
                x = float(val.split()[0])
            except ValueError:
                continue
            d = C.c_double(0)
            assert ref.ref_args_double(f.encode(), key.encode(), C.byref(d)) and d.value == x
    assert n_keys > 150


REF_RBD = os.path.join(HERE, "..", "oracle", "_ref", "libref_rbd.so")
CHARS = {"dog_slopes_mixed": "dog.txt", "goat_cliffs": "goat.txt", "raptor_narrow_gaps": "raptor.txt"}


@pytest.mark.skipif(not (os.path.exists(REF_RBD) and os.path.isdir("/root/reference/data/characters")),
                    reason="oracle/_ref/libref_rbd.so or the reference character files absent")
@pytest.mark.parametrize("scene", sorted(CHARS))
def test_rigid_body_model_matches_reference_code(assets, scene):
    """cRBDModel::Update + cRBDUtil::{BuildMassMat, BuildBiasForce, CalcGravityForce, BuildJacobian, CalcCoM} compiled from the
    reference vs oracle/rbd.h, the character read by cKinTree::Load / LoadBodyDefs vs the .trlpack the product and the oracle load."""
    from pyoracle import Oracle
    ref = C.CDLL(REF_RBD)
    ref.ref_rbd_create.restype = C.c_void_p
    ref.ref_rbd_create.argtypes = [C.c_char_p, C.c_double, C.c_double]
    h = ref.ref_rbd_create(("/root/reference/data/characters/" + CHARS[scene]).encode(), 0.0, -9.8)
    assert h
    h = C.c_void_p(h)
    o = Oracle(os.path.join(assets, scene + ".trlpack"), 1, 0)
    nd, nj = ref.ref_rbd_num_dof(h), ref.ref_rbd_num_joints(h)
    assert (nd, nj) == (o.ndof, o.nj)
    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(12):
        q = rng.uniform(-1.2, 1.2, nd); q[0] = rng.uniform(-5, 50); q[1] = rng.uniform(0, 2); q[2] = rng.uniform(-3.1, 3.1)
        qd = rng.normal(size=nd) * (0 if trial == 0 else 3.0)
        ref.ref_rbd_update(h, _p(q), _p(qd))
        o.set_state(0, q, qd)
        M = np.zeros((nd, nd)); Cb = np.zeros(nd); G = np.zeros(nd); J = np.zeros((6, nd)); com = np.zeros(3); cv = np.zeros(3)
        ref.ref_rbd_mass_bias(h, _p(M), _p(Cb)); ref.ref_rbd_gravity_force(h, _p(G)); ref.ref_rbd_jacobian(h, _p(J))
        ref.ref_rbd_com(h, _p(com), _p(cv))
        Mo, Co = o.rbd(0)
        pairs = [(M, Mo), (Cb, Co), (G, o.rbd_extra("gravity")), (J, o.rbd_extra("jacobian"))]
        oc, ov = o.com(0)
        pairs += [(com[:2], oc), (cv[:2], ov)]
        jp = np.zeros((nj, 3))
        for j in range(nj):
            ref.ref_rbd_joint_world_pos(h, j, _p(jp[j]))
        pairs.append((jp, o.rbd_extra("joint_pos")))
        for a, b in pairs:
            err = np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(a)))
            worst = max(worst, err)
            assert err < 1e-12, (scene, trial, err)
        assert np.allclose(M, M.T, atol=1e-12) and np.all(np.linalg.eigvalsh(M) > 0)
    ref.ref_rbd_destroy(h)
    print(f"{scene}: worst relative difference vs the compiled reference {worst:.2e}")


REF_CTRL = os.path.join(HERE, "..", "oracle", "_ref", "libref_ctrl.so")
# scene -> (character file, controller kind: 0 cDogControllerQ (fixed gait), 1 cDogControllerMACE, 2 cGoatControllerMACE,
#           4 cRaptorControllerMACE, env-steps)
CTRL_CASES = {"dog_flat": ("dog.txt", 0, 1200), "dog_slopes_mixed": ("dog.txt", 1, 1500), "goat_cliffs": ("goat.txt", 2, 1800),
              "raptor_narrow_gaps": ("raptor.txt", 4, 1500)}


@pytest.mark.skipif(not (os.path.exists(REF_CTRL) and os.path.isdir("/root/reference/data/characters")),
                    reason="oracle/_ref/libref_ctrl.so or the reference data files absent")
@pytest.mark.parametrize("scene", sorted(CTRL_CASES))
def test_controller_matches_reference_code(assets, scene):
    """The reference's OWN controller stack (DogController / DogControllerQ / DogControllerMACE / BaseControllerMACE /
    TerrainRLCharController / NNController / ImpPDController / PDController + KinTree / RBDUtil, compiled unmodified into
    oracle/_ref/libref_ctrl.so) is driven, env-step by env-step, with the oracle's state (pose, velocity, contacts, COM, terrain
    heights, and -- at policy decisions -- the oracle's network output), on a character back end that answers cSimCharacter's
    virtual calls with the reference's own cKinTree kinematics; the character class is the reference's own cSimDog / cSimRaptor,
    so fall and stumble detection (cSimCharSoftFall, cSimDog::HasStumbled / CheckFallContact / FailFallMisc) run as compiled.
    Compared every step: HasFallen / HasStumbled, the joint torques handed to
    ApplyControlForces vs the oracle's controller torques, the clamped torques the reference's own cJoint (sim/Joint.cpp, compiled)
    would pass to the physics vs the oracle's held torques, the gait-machine state / phase / action id, the action parameters,
    at every decision the policy state vector (terrain samples + character features) the reference built, and at the end of every
    cycle the reward c{Dog,Raptor}Controller::CalcReward returns."""
    from pyoracle import Oracle
    char_file, kind, n_steps = CTRL_CASES[scene]
    ref = C.CDLL(REF_CTRL)
    HFN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    o = Oracle(os.path.join(assets, scene + ".trlpack"), 1, 0)
    o.L.orc_sample_height.restype = C.c_double
    o.L.orc_sample_height.argtypes = [C.c_void_p, C.c_int, C.c_double]
    cb = HFN(lambda x, u: o.L.orc_sample_height(o.h, 0, x))
    ref.ref_ctrl_create.restype = C.c_void_p
    ref.ref_ctrl_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, HFN, C.c_void_p]
    cwd = os.getcwd()
    os.chdir("/root/reference")                       # the controller files are named relative to the reference's root
    try:
        h = ref.ref_ctrl_create(("data/characters/" + char_file).encode(), kind, 0.0, -9.8, cb, None)
    finally:
        os.chdir(cwd)
    assert h
    h = C.c_void_p(h)
    assert ref.ref_ctrl_valid(h) == 1
    if kind >= 1:
        n_out = 3 * (1 + (o.A - 1))
        ref.ref_ctrl_set_net_output(o.S, _p(np.zeros(n_out)), _p(np.ones(n_out)), n_out)
        assert ref.ref_ctrl_load_net(h) == 1           # the reference's own size checks: 283 inputs, 90 outputs for the dog
    nd = ref.ref_ctrl_num_dof(h)
    assert nd == o.ndof
    H = 1.0 / 600.0
    # cSimCharSoftFall::Reset with the initial state installed (the fall-distance check starts from the root position)
    q, qd, _, contact = o.get_state(0)
    com, cv = o.com(0)
    ref.ref_ctrl_set_state(h, _p(q), _p(qd), _p(contact.astype(np.uint8)), _p(com), _p(cv))
    ref.ref_char_reset(h)
    ref.ref_char_update.argtypes = [C.c_void_p, C.c_double]
    worst_tau = worst_state = 0.0
    n_stumbled = n_clamped = 0
    decisions = 0
    rewards = []
    ref.ref_ctrl_calc_reward.restype = C.c_double
    last_cycles = o.flags(0)[2]
    out_scale = np.ones(o.n_out if hasattr(o, "n_out") else 90)
    for k in range(n_steps):
        o.env_step(0, H)
        q, qd, _, contact = o.get_state(0)
        com, cv = o.com(0)
        fallen, stumbled, cycles = o.flags(0)
        if kind >= 1:
            y = o.net_out(0, 90)
            ref.ref_ctrl_set_net_output(o.S, _p(y), _p(out_scale), 90)
        ref.ref_ctrl_set_state(h, _p(q), _p(qd), _p(contact.astype(np.uint8)), _p(com), _p(cv))
        ref.ref_ctrl_update(h, C.c_double(H))
        ref.ref_char_update(h, H)                     # cSimCharSoftFall::Update: controller first, then the fall checks
        assert (ref.ref_char_has_fallen(h), ref.ref_char_has_stumbled(h)) == (fallen, stumbled), (scene, k)
        n_stumbled += stumbled
        tr = np.zeros(nd)
        ref.ref_ctrl_get_tau(h, _p(tr))
        to = o.last_tau(0)
        err = np.max(np.abs(tr - to)) / max(1.0, np.max(np.abs(to)))
        worst_tau = max(worst_tau, err)
        assert err < 1e-9, (scene, k, err)
        # the torque the physics receives: cJoint's accumulator after cJoint::ClampTotalTorque (limit from the PD parameters)
        ta = np.zeros(nd)
        ref.ref_ctrl_get_applied_tau(h, _p(ta))
        held = o.get_state(0)[2]
        assert np.max(np.abs(ta - held)) <= 1e-9 * max(1.0, np.max(np.abs(held))), (scene, k)
        n_clamped += int(np.any(np.abs(ta - tr) > 1e-6))
        f = np.zeros(64)
        n = ref.ref_ctrl_get_fsm(h, _p(f), 64)
        oc = o.get_ctrl(0)
        assert int(f[0]) == int(oc[0]) and abs(f[1] - oc[1]) < 1e-12, (scene, k, f[:3], oc[:3])     # gait state, phase
        if cycles != last_cycles:
            decisions += 1
            last_cycles = cycles
            if decisions >= 2:                                    # the start-up "cycle" has no previous cycle to rate (its tuple is dropped)
                r_ref = ref.ref_ctrl_calc_reward(h)               # reward of the cycle that just ended (cScenarioExp::CalcReward)
                assert abs(r_ref - o.calc_reward(0)) < 1e-12, (scene, k, r_ref, o.calc_reward(0))
                rewards.append(r_ref)
            if kind >= 1:
                s_ref = np.zeros(o.S)
                assert ref.ref_ctrl_poli_state(h, _p(s_ref), o.S) == o.S
                es = np.max(np.abs(s_ref - o.poli_state(0)))
                worst_state = max(worst_state, es)
                assert es < 1e-9, (scene, k, es)
    ref.ref_ctrl_destroy(h)
    assert decisions >= 3 and max(rewards) > 0.2
    print(f"{scene}: {n_steps} env-steps ({n_clamped} with a joint at its torque limit), {decisions} cycles; worst torque difference {worst_tau:.2e} (relative), "
          f"worst policy-state difference {worst_state:.2e}")


@pytest.mark.skipif(not (os.path.exists(REF_CTRL) and os.path.isdir("/root/reference/data/characters")),
                    reason="oracle/_ref/libref_ctrl.so or the reference data files absent")
@pytest.mark.parametrize("scene,char_file,kind", [("dog_slopes_mixed", "dog.txt", 1), ("raptor_narrow_gaps", "raptor.txt", 4)])
def test_fall_and_stumble_logic_matches_reference_code(assets, scene, char_file, kind):
    """cSimCharSoftFall + cSimDog / cSimRaptor (HasFallen, HasStumbled, CheckFallContact, FailFallMisc, the 5 s progress check and
    the discounted body-contact sum) compiled from the reference vs the oracle, on scripted sequences that trigger every branch:
    random part contacts, long body contact, a flipped root, and standing still for more than 5 s."""
    from pyoracle import Oracle
    ref = C.CDLL(REF_CTRL)
    HFN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    o = Oracle(os.path.join(assets, scene + ".trlpack"), 1, 0)
    cb = HFN(lambda x, u: 0.0)
    ref.ref_ctrl_create.restype = C.c_void_p
    ref.ref_ctrl_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, HFN, C.c_void_p]
    ref.ref_char_update.argtypes = [C.c_void_p, C.c_double]
    o.L.orc_fall_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        h = C.c_void_p(ref.ref_ctrl_create(("data/characters/" + char_file).encode(), kind, 0.0, -9.8, cb, None))
    finally:
        os.chdir(cwd)
    nd, nj = o.ndof, o.nj
    H = 1.0 / 600.0
    rng = np.random.default_rng(3)
    q0, qd0, _, _ = o.get_state(0)
    seen = {"fallen": 0, "stumbled": 0, "both_clear": 0}
    for episode, script in enumerate(("contacts", "body", "flip", "still")):
        q = q0.copy(); qd = qd0.copy()
        contact = np.zeros(nj, np.uint8)
        o.set_state(0, q, qd, None, contact)
        ref.ref_ctrl_set_state(h, _p(q), _p(qd), _p(contact), _p(np.zeros(2)), _p(np.zeros(2)))
        o.L.orc_fall_reset(o.h, 0)
        ref.ref_char_reset(h)
        steps = 3400 if script == "still" else 700
        for k in range(steps):
            if script != "still":
                q[0] += 4.0 * H                                            # keeps the progress check quiet
            if script == "contacts" and k % 7 == 0:
                contact = (rng.uniform(size=nj) < 0.15).astype(np.uint8)
            if script == "body":
                contact[:] = 0
                contact[rng.integers(0, 6)] = 1 if (k // 60) % 3 != 2 else 0   # spine / torso on the ground most of the time
            if script == "flip":
                q[2] = 0.005 * k * (1 if episode % 2 else -1)              # root pitch walks past 0.8 pi
            o.set_state(0, q, qd, None, contact)
            ref.ref_ctrl_set_state(h, _p(q), _p(qd), _p(contact), _p(np.zeros(2)), _p(np.zeros(2)))
            o.L.orc_fall_update(o.h, 0, H)
            ref.ref_char_update(h, H)
            fo, so, _ = o.flags(0)
            fr, sr = ref.ref_char_has_fallen(h), ref.ref_char_has_stumbled(h)
            assert (fr, sr) == (fo, so), (scene, script, k, (fr, sr), (fo, so))
            seen["fallen"] += fr; seen["stumbled"] += sr; seen["both_clear"] += (not fr and not sr)
        if script in ("body", "flip", "still"):
            assert ref.ref_char_has_fallen(h) == 1, (scene, script)       # every scripted failure mode ends fallen
    ref.ref_ctrl_destroy(h)
    assert min(seen.values()) > 100


@pytest.mark.skipif(not os.path.exists(REF_CTRL), reason="oracle/_ref/libref_ctrl.so not built")
def test_streaming_ground_matches_reference_code(libs):
    """cGroundVar2D (sim/GroundVar2D.cpp compiled from the reference: InitSegments, Update, BuildSegment, AddPadding, the segment
    ping-pong, SampleHeight) vs oracle/terrain.h's Ground: the view window [x - 2, x + 11] of a character running forward (and,
    for a while, backward) over every terrain type.  Vertex data and the segment bookkeeping must be identical; segment origins
    and sampled heights agree to 1e-6 relative -- the reference keeps each segment's origin in single precision at world scale
    (cWorld::SetPos), the oracle and the product keep it in double.  The stand-in for Bullet's body->getAabb is the ideal box of
    the height grid: real Bullet adds the shape's collision margin in float arithmetic, which is the one part of the reference's
    ground that cannot be reproduced without Bullet (it shifts where successive segments start by about a centimetre)."""
    terr, orc = libs
    ref = C.CDLL(REF_CTRL)
    for L, pre in ((ref, "ref"), (orc, "orc")):
        getattr(L, pre + "_ground_create").restype = C.c_void_p
        getattr(L, pre + "_ground_create").argtypes = [C.c_int, C.c_void_p, C.c_ulong, C.c_double, C.c_double]
        getattr(L, pre + "_ground_update").argtypes = [C.c_void_p, C.c_double, C.c_double]
        getattr(L, pre + "_ground_segment").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        getattr(L, pre + "_ground_sample").restype = C.c_double
        getattr(L, pre + "_ground_sample").argtypes = [C.c_void_p, C.c_double]
        getattr(L, pre + "_ground_flipped").argtypes = [C.c_void_p]
        getattr(L, pre + "_ground_destroy").argtypes = [C.c_void_p]
    rng = np.random.default_rng(11)
    cap = 2048
    rebuilds = 0
    for t in range(14):
        p = _params(terr, rng if t % 2 else None)
        seed = int(rng.integers(1, 2 ** 31))
        x = 0.0
        r = C.c_void_p(ref.ref_ground_create(t, _p(p), seed, x - 2.0, x + 11.0))
        o = C.c_void_p(orc.orc_ground_create(t, _p(p), seed, x - 2.0, x + 11.0))
        last_flip = ref.ref_ground_flipped(r)
        for step in range(260):
            x += rng.uniform(0.0, 0.9) if step < 200 else -rng.uniform(0.0, 0.9)      # forward 90 m, then back
            ref.ref_ground_update(r, x - 2.0, x + 11.0)
            orc.orc_ground_update(o, x - 2.0, x + 11.0)
            fl = ref.ref_ground_flipped(r)
            assert fl == orc.orc_ground_flipped(o), (t, step)
            if fl != last_flip:
                rebuilds += 1
                last_flip = fl
            for s in (0, 1):
                a = np.zeros(cap, np.float32); b = np.zeros(cap, np.float32)
                ma = C.c_double(0); mb = C.c_double(0)
                na = ref.ref_ground_segment(r, s, _p(a), cap, C.byref(ma))
                nb = orc.orc_ground_segment(o, s, _p(b), cap, C.byref(mb))
                assert na == nb and abs(ma.value - mb.value) < 2e-6 * max(1.0, abs(mb.value)), (t, step, s, na, nb, ma.value, mb.value)
                np.testing.assert_array_equal((a[:na] / np.float32(4.0)).view(np.uint32), b[:nb].view(np.uint32))   # stored x world scale 4
            for xs in rng.uniform(x - 2.0, x + 11.0, 6):
                # the reference's segment origin is a float: its sample equals the oracle's within that shift of the abscissa
                d = 2e-6 * max(1.0, abs(xs))
                band = [orc.orc_ground_sample(o, xs + e) for e in (-d, 0.0, d)]
                hr = ref.ref_ground_sample(r, xs)
                slack = 1e-7 + (max(band) - min(band))           # a vertex (kink) may lie inside the +-d window
                assert min(band) - slack <= hr <= max(band) + slack, (t, step, xs, hr, band)
        ref.ref_ground_destroy(r); orc.orc_ground_destroy(o)
    assert rebuilds > 14 * 4


@pytest.mark.skipif(not (os.path.exists(REF_CTRL) and os.path.isdir("/root/reference/data/characters")),
                    reason="oracle/_ref/libref_ctrl.so or the reference data files absent")
@pytest.mark.parametrize("scene,char_file,kind", [("dog_slopes_mixed", "dog.txt", 1), ("goat_cliffs", "goat.txt", 2),
                                                  ("raptor_narrow_gaps", "raptor.txt", 4)])
def test_output_offset_scale_matches_reference_code(assets, scene, char_file, kind):
    """cBaseControllerMACE::BuildNNOutputOffsetScale as compiled from the reference vs the PRODUCT's host function behind
    trl_get_output_offset_scale / trl_trainer_init_fresh (evaluated from the scene pack, no device needed)."""
    import deepterrainrl_b200 as trl
    L = trl.load_library()
    ref = C.CDLL(REF_CTRL)
    HFN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    cb = HFN(lambda x, u: 0.0)
    ref.ref_ctrl_create.restype = C.c_void_p
    ref.ref_ctrl_create.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_double, HFN, C.c_void_p]
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        h = C.c_void_p(ref.ref_ctrl_create(("data/characters/" + char_file).encode(), kind, 0.0, -9.8, cb, None))
    finally:
        os.chdir(cwd)
    frag = 28 if kind == 4 else 29
    n = 3 * (1 + frag)
    ref.ref_ctrl_set_net_output(275 if kind == 4 else 283, _p(np.zeros(n)), _p(np.ones(n)), n)
    assert ref.ref_ctrl_load_net(h) == 1
    ro = np.zeros(n); rs = np.zeros(n)
    assert ref.ref_ctrl_output_offset_scale(h, _p(ro), _p(rs), n) == n
    po = np.zeros(n); ps = np.zeros(n)
    rc = L.trl_pack_output_offset_scale(os.path.join(assets, scene + ".trlpack").encode(), _p(po), _p(ps), n)
    assert rc == 0, L.trl_last_error().decode()
    np.testing.assert_allclose(po, ro, rtol=0, atol=1e-15)
    np.testing.assert_allclose(ps, rs, rtol=1e-15, atol=0)
    assert np.all(ro[:3] == -0.5) and np.all(rs[:3] == 2.0)
    ref.ref_ctrl_destroy(h)


# ---------------------------------------------------------------------------------------------------- whole scenarios
SCN_ARGS = {"dog_slopes_mixed": "args/dog_slopes_mixed_args.txt", "goat_cliffs": "args/goat_cliffs_args.txt",
            "raptor_narrow_gaps": "args/raptor_narrow_gaps_args.txt"}
# (scene, mode 0 cScenarioPoliEval / 1 cScenarioExpMACE / 2 cScenarioExpMACE with exploration on, outer updates, exact segment origin)
SCN_CASES = [("goat_cliffs", 0, 400, True), ("raptor_narrow_gaps", 0, 400, True), ("dog_slopes_mixed", 0, int(os.environ.get("PIN_UPDATES", 200)), True),
             ("dog_slopes_mixed", 0, 200, False), ("dog_slopes_mixed", 1, 250, True), ("goat_cliffs", 1, 250, True),
             ("raptor_narrow_gaps", 1, 250, True), ("dog_slopes_mixed", 2, 400, True), ("goat_cliffs", 2, 400, True),
             ("raptor_narrow_gaps", 2, 400, True)]
# every other scenario arg file the reference ships (args/<name>_args.txt): the pack is made on the fly by the PRODUCT's native
# loader (trl.pack_from_args: arg file, character / controller / terrain JSON, Caffe HDF5 weights), so these cases also pin that
# loader against the compiled reference reading the same files itself; the sim_* files have no policy (fixed gait)
SCN_CASES += [(name, 0, 250, True) for name in ("dog_mixed", "dog_narrow_gaps", "dog_tight_gaps", "raptor_mixed", "raptor_slopes_mixed",
                                                 "sim_dog", "sim_goat", "sim_raptor")]


@pytest.mark.skipif(not (os.path.exists(REF_CTRL) and os.path.isdir("/root/reference/args")),
                    reason="oracle/_ref/libref_ctrl.so or the reference arg / data files absent")
@pytest.mark.parametrize("scene,mode,n_updates,exact_origin", SCN_CASES)
def test_scenario_matches_reference_code(assets, tmp_path, scene, mode, n_updates, exact_origin):
    """The reference's OWN scenario classes -- cScenario, cScenarioSimChar, cScenarioPoliEval, cScenarioExp, cScenarioExpMACE,
    compiled unmodified into oracle/_ref/libref_ctrl.so together with the controller stack, cGroundVar2D, cTerrainGen2D,
    cArgParser and tExpTuple -- run from the reference's own arg file (args/*_args.txt: character, controller, terrain, step
    counts, world scale).  Two factory functions are overridden (oracle/ref_ctrl_api.cpp): BuildWorld makes a world whose
    Update(h) calls back into this test, which advances the ORACLE by one env-step and installs the oracle's pose, velocity and
    contact bits as the simulation state; CreateCharacter makes the reference's cSimDog / cSimRaptor on the kinematic back end.
    cNeuralNet::Eval calls back with the input vector the reference built and receives the oracle's network output for it.
    Everything else is the reference's compiled code: the 20-step loop and its order (world, ground, character, post-substep),
    terrain streaming and sampling on its own cGroundVar2D, the controller, cycle counting, tuple recording (states, action,
    reward, flags, warm-up rule, ring buffer), the end-of-update fall test, distance bookkeeping, and Reset (pose0, controller
    reset, ground rebuild from the seeded generator, spawn height).

    Compared after every env-step: torques and gait state / phase; after every outer update: pose and velocity (i.e. the reset
    state when the episode ended), scenario time, cycle / episode / tuple counts, average distance; at the end the distance log or
    every recorded tuple.

    exact_origin: Bullet keeps a terrain segment's position in single precision, so the reference samples heights ~1e-7 m off;
    the stand-in world can keep the origin in double instead.  With it everything agrees to rounding (1e-10 bound, ~2e-12
    seen); without it (the reference as it is) the single-precision origin shows up as ~1e-6 relative in the torques.

    Mode 1 runs the exploration scenario with exp_rate = exp_base_rate = 0 and a Boltzmann temperature of 1e-6 (selection =
    argmax), cScenarioExp::CommandRandAction overridden to command the oracle's draw: the deterministic path, with the oracle's
    counter-based random streams as shipped.
    Mode 2 turns exploration ON (exp_rate 0.3, temperature 0.1, base-action rate 0.05): the reference draws from its
    process-global cMathUtil engine (seeded after Init), the oracle -- for this test -- from the restated cRand seeded alike, so
    random commands, random base actions (incl. the fragment assignment draws), Boltzmann actor selection, exploration noise and
    the ExpCritic / ExpActor flags are compared draw for draw (the product replaces the engine by counter-based streams per
    environment -- the reference's is shared by all threads -- but keeps this order and arithmetic)."""
    from pyoracle import Oracle, OracleTrainer
    seed = 77
    explore = mode == 2
    mode = min(mode, 1)
    ref = C.CDLL(REF_CTRL)
    ref.ref_world_exact_origin(1 if exact_origin else 0)
    tol = 1e-10 if exact_origin else 2e-5
    pack = os.path.join(assets, scene + ".trlpack")
    arg_file = SCN_ARGS.get(scene, "args/%s_args.txt" % scene)
    if not os.path.exists(pack):
        import deepterrainrl_b200 as trl
        pack = str(tmp_path / (scene + ".trlpack"))
        trl.pack_from_args(["-arg_file=", arg_file], "/root/reference", pack)
    o = Oracle(pack, 1, mode, terrain_seeds=[seed])
    L = o.L
    if explore:
        o.set_explore(1, 0.3, 0.1, 0.05)
        L.orc_use_ref_rand.argtypes = [C.c_void_p, C.c_ulong]
        L.orc_reseed_reset.argtypes = [C.c_void_p, C.c_int, C.c_ulong]
        L.orc_use_ref_rand(o.h, 999)
        L.orc_reseed_reset(o.h, 0, seed)
    elif mode == 1:
        o.set_explore(1, 0.0, 1e-6, 0.0)
    L.orc_end_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.orc_pending_command.argtypes = [C.c_void_p, C.c_int]
    WFN = C.CFUNCTYPE(None, C.c_double, C.c_void_p)
    NFN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p)
    CFN = C.CFUNCTYPE(C.c_int, C.c_void_p)
    nd = o.ndof
    DT = 1.0 / 30.0
    NN_LAYER = b"a1_ip0"              # the blob cScenarioPoliEval::RecordNNActivation dumps in the recorded run
    st = dict(h=None, steps=0, worst_tau=0.0, worst_pose=0.0, ended=False, in_update=False, cmp=False, err=None, evals=0)

    def ref_state():
        pose = np.zeros(nd); vel = np.zeros(nd); tau = np.zeros(nd)
        ref.ref_scn_get_state(st["h"], _p(pose), _p(vel), _p(tau))
        return pose, vel, tau

    def compare_step():
        _, _, tau = ref_state()
        to = o.last_tau(0)
        err = np.max(np.abs(tau - to)) / max(1.0, np.max(np.abs(to)))
        st["worst_tau"] = max(st["worst_tau"], err)
        f = np.zeros(3)
        ref.ref_scn_get_fsm(st["h"], _p(f))
        oc = o.get_ctrl(0)
        assert int(f[0]) == int(oc[0]) and abs(f[1] - oc[1]) < 1e-9, (st["steps"], f, oc[:3])
        if err >= tol and os.environ.get("PIN_DEBUG"):
            print("DBG step", st["steps"], "tau ref", np.round(tau, 6).tolist(), "tau orc", np.round(to, 6).tolist(), "contact used", st.get("prev_contact"), "now", st.get("cur_contact"), "fsm", f.tolist(), "diff", np.round(tau - to, 6).tolist())
        assert err < tol, (st["steps"], err)

    def world(hh, user):
        # cWorld::Update(h) of the compiled scenario: the previous env-step's controller output is compared first
        if st["err"] is not None:
            return
        try:
            if st["cmp"]:
                compare_step()
            o.env_step(0, hh)
            q, qd, _, contact = o.get_state(0)
            st["prev_contact"] = st.get("cur_contact"); st["cur_contact"] = contact.tolist()
            ref.ref_scn_set_state(st["h"], _p(q), _p(qd), _p(contact.astype(np.uint8)))
            st["steps"] += 1
            st["cmp"] = True
            cyc = o.flags(0)[2]
            if record and cyc != st["cyc"]:
                # a new gait cycle: what one tuple of the stream carries for it (state at the decision, action id + optimised
                # parameters; for the dog every parameter but the first is optimised)
                oc = o.get_ctrl(0)
                row = np.zeros(1 + 2 * o.S + o.A)
                row[1:1 + o.S] = o.poli_state(0)
                row[1 + o.S] = oc[11]
                row[2 + o.S:1 + o.S + o.A] = oc[13:12 + o.num_params]
                if st["cyc"] >= 1:           # cScenarioPoliEval::IsValidCycle: the first cycle is warm-up
                    blob = np.zeros(8192)
                    nb = L.orc_net_layer(o.h, _p(np.ascontiguousarray(row[1:1 + o.S])), NN_LAYER, _p(blob), 8192)
                    rec.record_nn_activation(oc[11], blob[:nb])          # RecordNNActivation comes first in NewCycleUpdate
                    rec.consume(row[None, :], np.zeros(1, np.uint32), np.zeros(1, np.int32))
                st["cyc"] = cyc
        except BaseException as e:          # an exception cannot cross the C frames: keep it for the main loop
            st["err"] = e

    def net(x, n_in, y, n_out, user):
        xi = np.ctypeslib.as_array(x, (n_in,)).copy()
        yo = o.net_eval(xi, n_out)
        for i in range(n_out):
            y[i] = yo[i]
        st["evals"] += 1

    def cmd(user):
        # cScenarioExp::Reset at the end of an update: the oracle's own end-of-update (and reset) has to come first
        if st["in_update"] and not st["ended"]:
            L.orc_end_update(o.h, 0, DT)
            st["ended"] = True
        return L.orc_pending_command(o.h, 0)

    wcb, ncb, ccb = WFN(world), NFN(net), CFN(cmd)
    n_out = 3 * (1 + (o.A - 1))
    has_net = not scene.startswith("sim_")
    if has_net:
        out_scale = np.ascontiguousarray(OracleTrainer(pack).get("out_scale"))
        ref.ref_ctrl_set_net_output(o.S, _p(np.zeros(n_out)), _p(out_scale), n_out)   # sizes for cNNController::LoadNet's checks; noise scale
    else:
        ref.ref_ctrl_set_net_output(0, _p(np.zeros(1)), _p(np.ones(1)), 0)            # cNeuralNet::HasNet() == false: fixed gait
    ref.ref_scn_create.restype = C.c_void_p
    ref.ref_scn_create.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_ulong, WFN, NFN, CFN, C.c_void_p, C.c_ulong]
    ref.ref_scn_update.argtypes = [C.c_void_p, C.c_double]
    ref.ref_scn_time.restype = C.c_double
    extra = [b"-exp_rate=", b"0", b"-exp_base_rate=", b"0", b"-exp_temp=", b"0.000001", b"-tuple_buffer_size=", b"4096"]
    if explore:
        extra = [b"-exp_rate=", b"0.3", b"-exp_base_rate=", b"0.05", b"-exp_temp=", b"0.1", b"-tuple_buffer_size=", b"4096"]
    # the analysis dumps of cScenarioPoliEval (RecordAction, RecordActionIDState; scenarios/ScenarioPoliEval.cpp:262-404), written by
    # the compiled scenario, against the product's deepterrainrl_b200.records.CycleRecorder fed with the same cycles
    record = scene == "dog_slopes_mixed" and mode == 0 and exact_origin
    st["cyc"] = 0
    if record:
        from deepterrainrl_b200.records import CycleRecorder
        rec = CycleRecorder(0, o.S, o.A, str(tmp_path / "actions.txt"), str(tmp_path / "ids.txt"), vel_file=str(tmp_path / "vel.txt"),
                            pack=pack, nn_activation_file=str(tmp_path / "nn.txt"), nn_activation_layer=NN_LAYER.decode())
        # cNeuralNet::GetLayerState of the compiled scenario's (stand-in) network: the oracle network's blob for the last Eval input
        LFN = C.CFUNCTYPE(C.c_int, C.c_char_p, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p)

        def layer(name, x, n_in, out, cap, user):
            xi = np.ctypeslib.as_array(x, (n_in,)).copy()
            buf = np.zeros(cap)
            n = L.orc_net_layer(o.h, _p(xi), name, _p(buf), cap)
            for i in range(min(n, cap)):
                out[i] = buf[i]
            return n
        lcb = LFN(layer)
        st["lcb"] = lcb
        ref.ref_ctrl_set_layer_cb(lcb)
        extra += [b"-record_nn_activation=", b"true", b"-nn_activation_layer=", NN_LAYER, b"-nn_activation_output_file=",
                  str(tmp_path / "ref_nn.txt").encode()]
        extra += [b"-record_vel=", b"true", b"-vel_output_file=", str(tmp_path / "ref_vel.txt").encode()]
        extra += [b"-record_actions=", b"true", b"-action_output_file=", str(tmp_path / "ref_actions.txt").encode(),
                  b"-record_action_id_state=", b"true", b"-action_id_state_output_file=", str(tmp_path / "ref_ids.txt").encode()]
    arr = (C.c_char_p * len(extra))(*extra)
    cwd = os.getcwd()
    os.chdir("/root/reference")                       # the arg file names its data files relative to the reference's root
    try:
        h = ref.ref_scn_create(arg_file.encode(), mode, arr, len(extra), seed, wcb, ncb,
                               C.cast(None, CFN) if explore else ccb, None, 999 if explore else 0)
    finally:
        os.chdir(cwd)
    assert h
    h = C.c_void_p(h)
    st["h"] = h
    try:
        assert ref.ref_scn_num_dof(h) == nd
        pose, vel, _ = ref_state()
        q, qd, _, _ = o.get_state(0)
        assert np.max(np.abs(pose - q)) < tol and np.max(np.abs(vel - qd)) < tol      # spawn state on the first terrain
        resets = 0
        for k in range(n_updates):
            st["in_update"], st["ended"] = True, False
            ref.ref_scn_update(h, DT)
            if st["err"] is not None:
                raise st["err"]
            if not st["ended"]:
                L.orc_end_update(o.h, 0, DT)
            st["in_update"] = False
            if record:
                rec.poll(o.get_ctrl(0), DT)               # RecordVel
            compare_step()
            st["cmp"] = False
            pose, vel, _ = ref_state()
            q, qd, _, _ = o.get_state(0)
            d = max(np.max(np.abs(pose - q)), np.max(np.abs(vel - qd)))
            st["worst_pose"] = max(st["worst_pose"], d)
            assert d < tol, (k, d)
            t_ref = ref.ref_scn_time(h)
            resets += t_ref == 0.0
            if mode == 0:
                cy, ep, ad = C.c_long(), C.c_long(), C.c_double()
                ref.ref_scn_eval_stats(h, C.byref(cy), C.byref(ep), C.byref(ad))
                es = o.eval_stats()
                assert (cy.value, ep.value) == (es["cycles"], es["episodes"]), (k, cy.value, ep.value, es)
                assert abs(ad.value - es["avg_dist"]) <= tol * max(1.0, abs(ad.value)), (k, ad.value, es)
            else:
                tc, cy = C.c_long(), C.c_long()
                ref.ref_scn_exp_counts(h, C.byref(tc), C.byref(cy))
                assert tc.value == L.orc_num_tuples(o.h), (k, tc.value, L.orc_num_tuples(o.h))
                assert cy.value == int(o.get_ctrl(0)[-2]), (k, cy.value)
        assert st["steps"] == 20 * n_updates and st["evals"] >= (5 if has_net else 0)
        if mode == 0:
            log = np.zeros(4096)
            n = ref.ref_scn_dist_log(h, _p(log), 4096)
            olog = o.dist_log(0)
            assert n == len(olog)
            assert np.allclose(log[:n], olog, rtol=0, atol=tol)
            if scene in ("goat_cliffs", "raptor_narrow_gaps"):
                assert n >= 1 and resets >= 1             # episodes ended by the reference's own fall test
            summary = f"{es['cycles']} cycles, {n} episodes"
            if record:
                ref_a = open(tmp_path / "ref_actions.txt").read().splitlines()
                n_act = len(ref_a) - rec.cycles           # InitActionRecord first lists the base actions ("%i, %.5f")
                assert n_act >= 1 and rec.cycles >= 10 and all(", " in l and ",\t" not in l for l in ref_a[:n_act])
                assert ref_a == open(tmp_path / "actions.txt").read().splitlines()          # table + records, byte for byte
                assert open(tmp_path / "ref_ids.txt").read() == open(tmp_path / "ids.txt").read()
                ref_nn = open(tmp_path / "ref_nn.txt").read()
                assert ref_nn == open(tmp_path / "nn.txt").read() and ref_nn.count("\n") == rec.cycles     # RecordNNActivation, byte for byte
                assert all(len(l.split(",\t")) == 1 + 128 for l in ref_nn.splitlines())                      # action id + the 128 units of a1_ip0
                ref_v = open(tmp_path / "ref_vel.txt").read().splitlines()
                got_v = open(tmp_path / "vel.txt").read().splitlines()
                assert len(ref_v) == len(got_v) == rec.cycles
                assert np.max(np.abs(np.array(ref_v, float) - np.array(got_v, float))) <= 2e-6     # "%f" of values equal to ~1e-12
                summary += f", {rec.cycles} action / action-id-state records byte-identical to the reference's files"
        else:
            rows, flags, _ = o.tuples()
            assert len(rows) >= 5
            W = o.S + o.A + o.S
            worst = 0.0
            for i in range(len(rows)):
                r, fl, row = C.c_double(), C.c_uint(), np.zeros(W)
                assert ref.ref_scn_get_tuple(h, i, C.byref(r), C.byref(fl), _p(row), W) == W
                assert fl.value == int(flags[i]), (i, fl.value, flags[i])
                worst = max(worst, abs(r.value - rows[i, 0]), np.max(np.abs(row - rows[i, 1:])))
            assert worst < tol, worst
            if explore:                                   # every flag combination the trainer distinguishes was produced
                assert np.any(flags & 2) and np.any(flags & 4) and np.any((flags & 6) == 6), flags
            summary = (f"{len(rows)} tuples ({int(np.sum((flags & 1) != 0))} ending in a fall, {int(np.sum((flags & 2) != 0))} ExpCritic, "
                       f"{int(np.sum((flags & 4) != 0))} ExpActor), worst tuple difference {worst:.1e}")
        print(f"{scene} mode {mode}{' exploring' if explore else ''} {'exact' if exact_origin else 'float'} origin: {st['steps']} env-steps, {resets} resets, {summary}; "
              f"worst torque difference {st['worst_tau']:.1e} (relative), worst pose difference {st['worst_pose']:.1e}")
    finally:
        ref.ref_scn_destroy(h)
        ref.ref_world_exact_origin(0)


# ---------------------------------------------------------------------------------------------------- the MACE trainer
REF_TRAIN = os.path.join(HERE, "..", "oracle", "_ref", "libref_train.so")


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="oracle/_ref/libref_train.so not built")
def test_mace_trainer_matches_reference_code(assets):
    """The reference's OWN trainer -- cTrainerInterface, cNeuralNetTrainer, cMACETrainer, cNeuralNetLearner, tExpTuple, with
    cMathUtil / cRand behind the sampling, compiled unmodified into oracle/_ref/libref_train.so -- is driven the way
    cScenarioTrain drives it (cNeuralNetLearner::Train(tuples) per full tuple buffer).  Its cNeuralNet (Caffe in the reference) is
    a stand-in whose EvalBatch / Train / CopyModel / CalcOffsetScale / SetInputOffsetScale call back into this test, which answers
    with the network-level operations of a first oracle trainer object (forward pass, one solver step on a problem, copy to the
    target).  A second oracle trainer object runs oracle/trainer.h's restatement of the whole algorithm independently on the same
    tuples, drawing its sample indices from the restated cRand (seeded like cMathUtil::SeedRand).

    Everything the reference trainer does itself is therefore compared as compiled: replay memory (float rows, ring buffer with
    wrap-around, flag buffer), critic / actor index buffers incl. swap-removal on overwrite, minibatch sampling (order and number
    of draws), target values (reward normalisation, fail flag, discounted max over the target net's critic outputs), the problem
    matrices handed to the solver, the positive-temporal-difference filter and FIFO of the actor batch buffer, the initial
    stage (input offset / scale from the first samples), iteration counters, target refresh every `freeze_target_iters`.
    Since both sides use the same network arithmetic, the weights must come out bit-identical -- any difference in a sampled id,
    a label or the order of solver steps would show."""
    from pyoracle import OracleTrainer
    pack = os.path.join(assets, "dog_slopes_mixed.trlpack")
    ref = C.CDLL(REF_TRAIN)
    kw = dict(replay_cap=400, num_init_samples=96, num_steps_per_iter=2, freeze_target_iters=3, init_input_offset_scale=1, seed=11)
    eng = OracleTrainer(pack, **kw)      # the network under the compiled reference trainer
    orc = OracleTrainer(pack, **kw)      # the restated trainer, on its own
    L = orc.L
    L.orc_trainer_use_ref_rand.argtypes = [C.c_void_p, C.c_ulong]
    L.orc_trainer_use_ref_rand(orc.h, 4242)
    L.orc_trainer_solver_step.restype = C.c_double
    S, no, W = orc.n_in, orc.n_out, orc.W
    A = W - 1 - 2 * S
    B, n_frags, frag = 32, 3, A - 1
    DP = C.POINTER(C.c_double)
    EV = C.CFUNCTYPE(None, C.c_int, DP, C.c_int, DP, C.c_void_p)
    TR = C.CFUNCTYPE(None, C.c_int, DP, DP, C.c_int, C.c_void_p)
    CP = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_void_p)
    CO = C.CFUNCTYPE(None, DP, C.c_int, DP, DP, C.c_void_p)
    SO = C.CFUNCTYPE(None, C.c_int, DP, DP, C.c_void_p)
    log = dict(evals=0, trains=0, copies=[], calc=0, set=[], err=None)

    def guarded(fn):
        def w(*a):
            try:
                fn(*a)
            except BaseException as e:       # cannot cross the C frames
                log["err"] = e
        return w

    def ev(net, X, Bn, Y, u):
        x = np.ctypeslib.as_array(X, (Bn, S)).copy()
        y = np.zeros((Bn, no))
        L.orc_trainer_eval_batch(eng.h, 1 if net == 1 else 0, _p(x), Bn, _p(y))
        np.ctypeslib.as_array(Y, (Bn, no))[:] = y
        log["evals"] += 1

    def tr(net, X, Y, Bn, u):
        assert net == 0 and Bn == B
        x = np.ctypeslib.as_array(X, (Bn, S)).copy()
        y = np.ctypeslib.as_array(Y, (Bn, no)).copy()
        L.orc_trainer_solver_step(eng.h, _p(x), _p(y))
        log["trains"] += 1

    def cp(dst, src, u):
        log["copies"].append((dst, src))
        if (dst, src) == (1, 0):
            L.orc_trainer_copy_to_target(eng.h)

    def co(X, n, off, sc, u):
        x = np.ctypeslib.as_array(X, (n, S)).copy()
        o, s = np.zeros(S), np.zeros(S)
        L.orc_calc_offset_scale(_p(x), n, S, _p(o), _p(s))
        np.ctypeslib.as_array(off, (S,))[:] = o
        np.ctypeslib.as_array(sc, (S,))[:] = s
        log["calc"] += 1

    def so(net, off, sc, u):
        o = np.ctypeslib.as_array(off, (S,)).copy()
        s = np.ctypeslib.as_array(sc, (S,)).copy()
        if net in (0, 1):
            L.orc_trainer_set_input_offset_scale(eng.h, net, _p(o), _p(s))
        log["set"].append(net)

    cbs = (EV(guarded(ev)), TR(guarded(tr)), CP(guarded(cp)), CO(guarded(co)), SO(guarded(so)))
    p = np.array([S, no, B, n_frags, frag, kw["replay_cap"], kw["num_init_samples"], kw["num_steps_per_iter"],
                  kw["freeze_target_iters"], 0.9, kw["init_input_offset_scale"], 4242], float)
    ref.ref_trainer_create.restype = C.c_void_p
    ref.ref_trainer_create.argtypes = [C.c_void_p, EV, TR, CP, CO, SO, C.c_void_p]
    ref.ref_trainer_learn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    h = C.c_void_p(ref.ref_trainer_create(_p(p), *cbs, None))
    assert log["copies"] == [(1, 0)]                  # cMACETrainer::BuildNetPool: the target starts as a copy
    assert ref.ref_trainer_width(h) == W
    rng = np.random.default_rng(0)
    in_off, in_scale = orc.get("in_off"), orc.get("in_scale")

    def make(n):
        rows = np.zeros((n, W))
        rows[:, 0] = rng.uniform(0, 1, n)
        for k in (1, 1 + S + A):
            rows[:, k:k + S] = rng.normal(size=(n, S)) / np.where(in_scale == 0, 1.0, in_scale) - in_off
        rows[:, 1 + S] = rng.integers(0, n_frags, n)
        rows[:, 2 + S:1 + S + A] = rng.normal(size=(n, frag))
        fl = np.zeros(n, np.uint32)
        u = rng.uniform(size=n)
        fl[u < 0.1] |= 1                    # cMACETrainer::eFlagFail
        fl[(u > 0.3) & (u < 0.5)] |= 2      # eFlagExpCritic
        fl[(u > 0.4) & (u < 0.8)] |= 4      # eFlagExpActor
        return rows, fl

    try:
        n_calls = 22
        for k in range(n_calls):
            rows, fl = make(32)
            ref.ref_trainer_learn(h, _p(rows), _p(fl), 32)
            if log["err"] is not None:
                raise log["err"]
            orc.add_tuples(rows, fl)
            orc.train()
            c = (C.c_long * 8)()
            ref.ref_trainer_counters(h, c)
            oc = orc.counters()
            assert (c[0], c[1], c[2], c[3], c[4], c[5]) == (oc["iter"], oc["actor_iter"], oc["stage"], oc["num"], oc["head"], oc["total"]), (k, list(c), oc)
            assert (c[6], c[7]) == (c[0], c[5])                       # the learner mirrors the trainer's iteration / tuple counts
            for which, name in enumerate(("critic", "actor", "actor_batch")):
                buf = np.zeros(4096, np.int32)
                n = ref.ref_trainer_list(h, which, _p(buf), 4096)
                ol = orc.lists(name)
                assert n == len(ol) and np.array_equal(buf[:n], ol), (k, name)
            assert np.array_equal(eng.get("theta"), orc.get("theta")), (k, np.max(np.abs(eng.get("theta") - orc.get("theta"))))
        assert np.array_equal(eng.get("target"), orc.get("target"))
        assert np.array_equal(eng.get("in_off"), orc.get("in_off")) and np.array_equal(eng.get("in_scale"), orc.get("in_scale"))
        ids = np.arange(kw["replay_cap"], dtype=np.int32)
        rr = np.zeros((len(ids), W), np.float32)
        rf = np.zeros(len(ids), np.int32)
        ref.ref_trainer_rows(h, _p(ids), len(ids), _p(rr), _p(rf))
        orows, oflags = orc.rows(ids)
        assert np.array_equal(rr, orows) and np.array_equal(rf, oflags)
        oc = orc.counters()
        assert oc["total"] == 32 * n_calls > kw["replay_cap"] and oc["iter"] >= 15 and oc["actor_iter"] >= 5      # wrapped, trained
        assert log["calc"] == 1 and log["set"] == [0, 1] and log["copies"].count((1, 0)) >= 4
        print(f"compiled cMACETrainer vs oracle/trainer.h: {oc['iter']} iterations ({log['trains']} solver steps, {oc['actor_iter']} of "
              f"them actor steps), {oc['total']} tuples through a {kw['replay_cap']}-row replay memory, weights bit-identical")
    finally:
        ref.ref_trainer_destroy(h)


# ---------------------------------------------------------------------------------------------------- the training scenario
@pytest.mark.skipif(not (os.path.exists(REF_CTRL) and os.path.isdir("/root/reference/args")),
                    reason="oracle/_ref/libref_ctrl.so or the reference arg / data files absent")
@pytest.mark.parametrize("scene,arg_file,terrain_file,exp_rate,exp_base_rate",
                         [("dog_slopes_mixed", "args/opt_args_train_mace.txt", "data/terrain/slopes_mixed.txt", 0.2, 0.002),
                          ("raptor_narrow_gaps", "args/opt_args_train_raptor_mace.txt", "data/terrain/narrow_gaps.txt", 0.1, 0.001)])
def test_training_scenario_matches_reference_code(assets, scene, arg_file, terrain_file, exp_rate, exp_base_rate):
    """The reference's OWN training scenario -- cScenarioTrain + cScenarioTrainMACE, with the compiled cMACETrainer /
    cNeuralNetTrainer / cNeuralNetLearner behind it and one compiled cScenarioExpMACE (fake-backed as in
    test_scenario_matches_reference_code) in its pool -- runs from args/opt_args_train_mace.txt / opt_args_train_raptor_mace.txt
    (overrides: the terrain of the pack, small replay memory / initial sample count / tuple buffer / anneal horizons so that everything happens within
    ~1200 updates; the net's batch size, which Caffe reads from the net file, is 8 here).  Only BuildExpScene is overridden.
    As compiled: BuildScenePool (initial exploration rates, curriculum phase, "rebuild ground" reset), InitTrainer / InitLearners
    (learner network = the controller's network, first SyncNet), UpdateExpScene (update, IsTupleBufferFull, UpdateTrainer ->
    cNeuralNetLearner::Train -> AddTuples / Train / SyncNet, CalcExpRate / CalcExpTemp / CalcExpBaseRate / CalcCurriculumPhase,
    SetExp*, UpdateSceneCurriculum, ResetTupleBuffer) -- with exploration ON and annealed, every random draw (exploration AND
    minibatch sampling) coming from the reference's single cMathUtil engine.

    Against it, in lock-step: the oracle's exploration environment + the oracle's trainer (sharing one restated cRand, as the
    reference shares its engine) + the PRODUCT's schedule function (trl_train_schedule through deepterrainrl_b200.train.TrainSchedule)
    + weight hand-over to the environment after every trainer call.  The compiled trainer's network operations are answered by a
    second oracle network object; the reference controller's evaluations by that object's weights after each SyncNet.

    Compared: torques / gait state at every env-step, pose after every update, the exploration rates the compiled scenario holds
    vs the product's schedule, the schedule function itself, trainer iteration / tuple counts, and after every trainer call the
    weights -- bit-identical, which they can only be if every exploration draw, tuple, sampled index, label and solver step
    agreed."""
    from pyoracle import Oracle, OracleTrainer
    from deepterrainrl_b200.train import TrainSchedule
    pack = os.path.join(assets, scene + ".trlpack")
    ref = C.CDLL(REF_CTRL)
    ref.ref_world_exact_origin(1)
    tol = 1e-9
    gseed, rseed, B, TB = 77, 999, 8, 8
    kw = dict(replay_cap=400, num_init_samples=24, num_steps_per_iter=1, freeze_target_iters=3, init_input_offset_scale=1, seed=1)
    o = Oracle(pack, 1, 1, terrain_seeds=[gseed])           # the oracle's exploration environment
    oe = Oracle(pack, 1, 1, terrain_seeds=[gseed])          # holds the weights the reference controller evaluates (after SyncNet)
    orc = OracleTrainer(pack, **kw)                          # the oracle's trainer
    eng = OracleTrainer(pack, **kw)                          # the network under the compiled reference trainer
    L = o.L
    for f, a in (("orc_use_ref_rand", [C.c_void_p, C.c_ulong]), ("orc_reseed_reset", [C.c_void_p, C.c_int, C.c_ulong]),
                 ("orc_end_update", [C.c_void_p, C.c_int, C.c_double]), ("orc_trainer_share_rand", [C.c_void_p, C.c_void_p]),
                 ("orc_set_net_from_trainer", [C.c_void_p, C.c_void_p]), ("orc_set_terrain_lerp", [C.c_void_p, C.c_double]),
                 ("orc_trainer_set_batch", [C.c_void_p, C.c_int])):
        getattr(L, f).argtypes = a
    L.orc_trainer_solver_step.restype = C.c_double
    L.orc_trainer_set_batch(orc.h, B)
    L.orc_trainer_set_batch(eng.h, B)
    SCHED = dict(init_exp_rate=0.5, exp_rate=exp_rate, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.3, exp_base_rate=exp_base_rate,
                 trainer_num_anneal_iters=12, exp_base_anneal_iters=8, trainer_curriculum_iters=10)
    sched = TrainSchedule(**SCHED)                           # the product's trl_train_schedule
    L.orc_use_ref_rand(o.h, rseed)
    o.set_explore(1, SCHED["init_exp_rate"], SCHED["init_exp_temp"], SCHED["init_exp_base_rate"])
    L.orc_set_terrain_lerp(o.h, 1.0)                         # gInitCurriculumPhase (scenarios/ScenarioTrain.cpp:6,213)
    L.orc_reseed_reset(o.h, 0, gseed)
    L.orc_trainer_share_rand(orc.h, o.h)
    S, no = orc.n_in, orc.n_out
    nd = o.ndof
    DP = C.POINTER(C.c_double)
    WFN = C.CFUNCTYPE(None, C.c_double, C.c_void_p)
    EV = C.CFUNCTYPE(None, C.c_int, DP, C.c_int, DP, C.c_void_p)
    TR = C.CFUNCTYPE(None, C.c_int, DP, DP, C.c_int, C.c_void_p)
    CP = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_void_p)
    CO = C.CFUNCTYPE(None, DP, C.c_int, DP, DP, C.c_void_p)
    SO = C.CFUNCTYPE(None, C.c_int, DP, DP, C.c_void_p)
    st = dict(h=None, steps=0, worst_tau=0.0, cmp=False, err=None)
    log = dict(evals=[0, 0, 0], trains=0, copies=[], calc=0, set=[])

    def ref_state():
        pose = np.zeros(nd); vel = np.zeros(nd); tau = np.zeros(nd)
        ref.ref_strain_get_state(st["h"], _p(pose), _p(vel), _p(tau))
        return pose, vel, tau

    def compare_step():
        _, _, tau = ref_state()
        to = o.last_tau(0)
        err = np.max(np.abs(tau - to)) / max(1.0, np.max(np.abs(to)))
        st["worst_tau"] = max(st["worst_tau"], err)
        f = np.zeros(64)
        ref.ref_strain_get_fsm(st["h"], _p(f), 64)
        oc = o.get_ctrl(0)
        assert int(f[0]) == int(oc[0]) and abs(f[1] - oc[1]) < 1e-9, (st["steps"], f[:3], oc[:3])
        assert err < tol, (st["steps"], err)

    def guarded(fn):
        def w(*a):
            if st["err"] is not None:
                return
            try:
                fn(*a)
            except BaseException as e:       # cannot cross the C frames
                st["err"] = e
        return w

    def world(hh, u):
        if st["cmp"]:
            compare_step()
        o.env_step(0, hh)
        q, qd, _, contact = o.get_state(0)
        ref.ref_strain_set_state(st["h"], _p(q), _p(qd), _p(contact.astype(np.uint8)))
        st["steps"] += 1
        st["cmp"] = True

    # cNeuralNet instances in construction order: 0 the controller's (= the learner's) network, 1 the trainer's, 2 its target
    def ev(net, X, Bn, Y, u):
        x = np.ctypeslib.as_array(X, (Bn, S)).copy()
        log["evals"][net] += 1
        if net == 0:
            assert Bn == 1
            np.ctypeslib.as_array(Y, (1, no))[0] = oe.net_eval(x[0], no)
        else:
            y = np.zeros((Bn, no))
            L.orc_trainer_eval_batch(eng.h, 1 if net == 2 else 0, _p(x), Bn, _p(y))
            np.ctypeslib.as_array(Y, (Bn, no))[:] = y

    def tr(net, X, Y, Bn, u):
        assert net == 1 and Bn == B
        x = np.ctypeslib.as_array(X, (Bn, S)).copy()
        y = np.ctypeslib.as_array(Y, (Bn, no)).copy()
        L.orc_trainer_solver_step(eng.h, _p(x), _p(y))
        log["trains"] += 1

    def cp(dst, src, u):
        log["copies"].append((dst, src))
        if (dst, src) == (2, 1):
            L.orc_trainer_copy_to_target(eng.h)
        elif (dst, src) == (0, 1):
            L.orc_set_net_from_trainer(oe.h, eng.h)          # cNeuralNetLearner::SyncNet: the controller gets the trained weights
        else:
            raise AssertionError((dst, src))

    def co(X, n, off, sc, u):
        x = np.ctypeslib.as_array(X, (n, S)).copy()
        of, s = np.zeros(S), np.zeros(S)
        L.orc_calc_offset_scale(_p(x), n, S, _p(of), _p(s))
        np.ctypeslib.as_array(off, (S,))[:] = of
        np.ctypeslib.as_array(sc, (S,))[:] = s
        log["calc"] += 1

    def so(net, off, sc, u):
        assert net in (1, 2)
        of = np.ctypeslib.as_array(off, (S,)).copy()
        s = np.ctypeslib.as_array(sc, (S,)).copy()
        L.orc_trainer_set_input_offset_scale(eng.h, 1 if net == 2 else 0, _p(of), _p(s))
        log["set"].append(net)

    cbs = (WFN(guarded(world)), EV(guarded(ev)), TR(guarded(tr)), CP(guarded(cp)), CO(guarded(co)), SO(guarded(so)))
    out_scale = np.ascontiguousarray(orc.get("out_scale"))
    ref.ref_ctrl_set_net_output(S, _p(np.zeros(no)), _p(out_scale), no)           # the exploration-noise scale the controller reads
    extra = ["-init_exp_rate=", "0.5", "-init_exp_base_rate=", "0.3", "-terrain_file=", terrain_file,
             "-tuple_buffer_size=", str(TB), "-trainer_replay_mem_size=", "400", "-trainer_num_init_samples=", "24",
             "-trainer_freeze_target_iters=", "3", "-trainer_num_anneal_iters=", "12", "-exp_base_anneal_iters=", "8",
             "-trainer_curriculum_iters=", "10", "-trainer_int_iter=", "0", "-trainer_iters_per_output=", "100000",
             "-output_path=", "/tmp/ref_strain_model.h5"]
    extra = [e.encode() for e in extra]
    arr = (C.c_char_p * len(extra))(*extra)
    dims = np.array([S, no, B], np.int32)
    ref.ref_strain_create.restype = C.c_void_p
    ref.ref_strain_create.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_ulong, C.c_ulong, C.c_void_p, WFN, EV, TR, CP, CO, SO, C.c_void_p]
    ref.ref_strain_update.argtypes = [C.c_void_p, C.c_double]
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        h = ref.ref_strain_create(arg_file.encode(), arr, len(extra), gseed, rseed, _p(dims), *cbs, None)
    finally:
        os.chdir(cwd)
    assert h and st["err"] is None, st["err"]
    h = C.c_void_p(h)
    st["h"] = h
    try:
        assert log["copies"] == [(2, 1), (0, 1)]       # cMACETrainer::BuildNetPool (target), cNeuralNetLearner::Init (SyncNet)
        pose, _, _ = ref_state()
        assert np.max(np.abs(pose - o.get_state(0)[0])) < tol
        DT = 1.0 / 30.0
        n_calls = 0
        rates_seen = []
        for k in range(1250):
            ref.ref_strain_update(h, DT)
            if st["err"] is not None:
                raise st["err"]
            L.orc_end_update(o.h, 0, DT)
            compare_step()
            st["cmp"] = False
            if L.orc_num_tuples(o.h) >= TB:             # the oracle's side of cScenarioTrain::UpdateExpScene
                rows, flags, _ = o.tuples()
                assert len(rows) == TB
                orc.add_tuples(rows, flags)
                orc.train()
                L.orc_set_net_from_trainer(o.h, orc.h)
                it = orc.counters()["iter"]
                s = sched(it)
                o.set_explore(1, s["exp_rate"], s["exp_temp"], s["exp_base_rate"])
                L.orc_set_terrain_lerp(o.h, s["curriculum_phase"])
                o.reset_tuples()
                n_calls += 1
                cnt, rates, r4 = (C.c_long * 3)(), (C.c_double * 3)(), (C.c_double * 4)()
                ref.ref_strain_status(h, cnt, rates)
                ref.ref_strain_schedule(h, it, r4)
                assert (cnt[0], cnt[1], cnt[2]) == (it, TB * n_calls, 0), (k, list(cnt), it)
                want = [s["exp_rate"], s["exp_temp"], s["exp_base_rate"]]
                assert np.allclose(list(rates), want, rtol=1e-15, atol=0), (list(rates), want)        # what the compiled scenario now holds
                assert np.allclose(list(r4), want + [s["curriculum_phase"]], rtol=1e-15, atol=0)      # cScenarioTrain::Calc* vs trl_train_schedule
                assert np.array_equal(eng.get("theta"), orc.get("theta")), (k, it)
                rates_seen.append(want)
            pose, vel, _ = ref_state()
            q, qd, _, _ = o.get_state(0)
            assert max(np.max(np.abs(pose - q)), np.max(np.abs(vel - qd))) < tol, k
        oc = orc.counters()
        assert oc["iter"] >= 6 and log["trains"] >= 6 and log["evals"][0] >= 30 and log["calc"] == 1
        assert np.array_equal(eng.get("target"), orc.get("target"))
        assert rates_seen[-1][0] < rates_seen[0][0] and rates_seen[-1][1] < rates_seen[0][1] and rates_seen[-1][2] < rates_seen[0][2]
        print(f"{scene}: compiled cScenarioTrainMACE vs oracle loop: {st['steps']} env-steps, {n_calls} trainer calls, {oc['iter']} iterations "
              f"({log['trains']} solver steps), rates annealed to {rates_seen[-1]}, weights bit-identical; worst torque difference "
              f"{st['worst_tau']:.1e}")
    finally:
        ref.ref_strain_destroy(h)
        ref.ref_world_exact_origin(0)
