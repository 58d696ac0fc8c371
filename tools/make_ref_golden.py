"""Golden vectors produced by the REFERENCE's own compiled code (oracle/_ref/libref_ctrl.so, libref_terrain.so; see oracle/README.md),
committed under tests/golden/ref_*.npz so that the oracle can be checked against reference outputs on a machine that has neither
/root/reference nor oracle/_ref (tests/test_ref_golden_cpu.py).  Run here, where the reference is mounted:

    make -C oracle ref && python tools/make_ref_golden.py

ref_scenario_<scene>.npz: the reference's compiled cScenarioPoliEval run from its own arg file, with the world update handed to the
oracle's physics (the lock-step harness of tests/test_ref_pinning_cpu.py, double-precision segment origin).  Stored per env-step:
the joint torques the compiled controller handed to cSimCharacter::ApplyControlForces, the clamped torques its cJoint would
apply, gait state and phase; per outer update: root position after the reference's own end-of-update handling (reset included);
at the end: cycles, episodes, average distance, distance log.  Because the oracle is deterministic, re-running it alone
reproduces the state sequence these outputs were computed from.
ref_terrain.npz: cTerrainGen2D strips for all 14 terrain types and two seeds (default parameters of the reference)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyoracle import Oracle, OracleTrainer  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SEED = 77
CASES = {"dog_slopes_mixed": 60, "goat_cliffs": 120, "raptor_narrow_gaps": 60}
TYPES = ["flat", "gaps", "steps", "walls", "bumps", "mixed", "narrow_gaps", "slopes", "slopes_gaps", "slopes_walls", "slopes_steps",
         "slopes_mixed", "slopes_narrow_gaps", "cliffs"]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def scenario(scene, n_updates):
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_ctrl.so"))
    ref.ref_world_exact_origin(1)
    pack = os.path.join(ROOT, "assets", scene + ".trlpack")
    o = Oracle(pack, 1, 0, terrain_seeds=[SEED])
    L = o.L
    L.orc_end_update.argtypes = [C.c_void_p, C.c_int, C.c_double]
    WFN = C.CFUNCTYPE(None, C.c_double, C.c_void_p)
    NFN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.c_int, C.c_void_p)
    CFN = C.CFUNCTYPE(C.c_int, C.c_void_p)
    nd = o.ndof
    st = dict(h=None, first=True)
    rec = dict(tau=[], applied=[], fsm=[], root=[])

    def grab():
        pose = np.zeros(nd); vel = np.zeros(nd); tau = np.zeros(nd)
        ref.ref_scn_get_state(st["h"], _p(pose), _p(vel), _p(tau))
        f = np.zeros(3)
        ref.ref_scn_get_fsm(st["h"], _p(f))
        ta = np.zeros(nd)
        ref.ref_scn_get_applied_tau(st["h"], _p(ta))
        rec["tau"].append(tau)
        rec["applied"].append(ta)
        rec["fsm"].append(f[:2].copy())

    def world(hh, user):
        if not st["first"]:
            grab()                      # the controller output of the previous env-step
        st["first"] = False
        o.env_step(0, hh)
        q, qd, _, contact = o.get_state(0)
        ref.ref_scn_set_state(st["h"], _p(q), _p(qd), _p(contact.astype(np.uint8)))

    def net(x, n_in, y, n_out, user):
        xi = np.ctypeslib.as_array(x, (n_in,)).copy()
        yo = o.net_eval(xi, n_out)
        for i in range(n_out):
            y[i] = yo[i]

    cbs = (WFN(world), NFN(net))
    n_out = 3 * (1 + (o.A - 1))
    out_scale = np.ascontiguousarray(OracleTrainer(pack).get("out_scale"))
    ref.ref_ctrl_set_net_output(o.S, _p(np.zeros(n_out)), _p(out_scale), n_out)
    ref.ref_scn_create.restype = C.c_void_p
    ref.ref_scn_create.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int, C.c_ulong, WFN, NFN, CFN, C.c_void_p, C.c_ulong]
    ref.ref_scn_update.argtypes = [C.c_void_p, C.c_double]
    cwd = os.getcwd()
    os.chdir("/root/reference")
    try:
        h = ref.ref_scn_create(("args/%s_args.txt" % scene).encode(), 0, None, 0, SEED, cbs[0], cbs[1], C.cast(None, CFN), None, 0)
    finally:
        os.chdir(cwd)
    assert h
    h = C.c_void_p(h)
    st["h"] = h
    for k in range(n_updates):
        ref.ref_scn_update(h, 1.0 / 30.0)
        grab()                          # the last env-step of the update
        st["first"] = True
        L.orc_end_update(o.h, 0, 1.0 / 30.0)
        pose = np.zeros(nd); vel = np.zeros(nd); tau = np.zeros(nd)
        ref.ref_scn_get_state(h, _p(pose), _p(vel), _p(tau))
        rec["root"].append(pose[:3].copy())
        assert np.max(np.abs(pose - o.get_state(0)[0])) < 1e-9       # the fixture is only written from a run that agrees
    cy, ep, ad = C.c_long(), C.c_long(), C.c_double()
    ref.ref_scn_eval_stats(h, C.byref(cy), C.byref(ep), C.byref(ad))
    log = np.zeros(4096)
    n = ref.ref_scn_dist_log(h, _p(log), 4096)
    ref.ref_scn_destroy(h)
    tau = np.array(rec["tau"])
    assert tau.shape == (20 * n_updates, nd)
    np.savez_compressed(os.path.join(OUT, "ref_scenario_%s.npz" % scene), tau=tau, applied=np.array(rec["applied"]), fsm=np.array(rec["fsm"]),
                        root=np.array(rec["root"]),
                        stats=np.array([cy.value, ep.value, ad.value]), dist_log=log[:n], seed=SEED, n_updates=n_updates)
    print(scene, tau.shape, "cycles", cy.value, "episodes", ep.value)


def terrain():
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_terrain.so"))
    ref.ref_terrain_default_params.argtypes = [C.c_void_p]
    params = np.zeros(40)
    ref.ref_terrain_default_params(_p(params))
    ref.ref_terrain_build.argtypes = [C.c_int, C.c_void_p, C.c_ulong, C.c_double, C.c_void_p, C.c_int, C.c_void_p]
    out = {"params": params}
    for t, name in enumerate(TYPES):
        for seed in (3, 4242):
            buf = np.zeros(4096, np.float32)
            tw = C.c_double()
            n = ref.ref_terrain_build(t, _p(params), seed, 40.0, _p(buf), 4096, C.byref(tw))
            out["%s_%d" % (name, seed)] = buf[:n].copy()
    np.savez_compressed(os.path.join(OUT, "ref_terrain.npz"), **out)
    print("terrain", len(out) - 1, "strips")


if __name__ == "__main__":
    for scene, n in CASES.items():
        scenario(scene, n)
    terrain()
