"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the header declares, fails loudly without
a GPU; the asset packs carry the reference's cross-file identities."""
import os
import re

import numpy as np
import pytest

import deepterrainrl_b200 as trl
from pack_scene import read_pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    trl.build_library()
    L = trl.load_library()
    hdr = open(os.path.join(ROOT, "include", "terrainrl_b200.h")).read()
    declared = set(re.findall(r"\b(trl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/terrainrl_b200.h but not exported"
    assert declared == set(trl.scenario.EXPORTS)


def test_public_header_is_self_contained_c99(tmp_path):
    """include/terrainrl_b200.h is what a maintainer's cgo / ctypes / C++ binding includes first: it must compile on its own as C99 and as C++"""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "terrainrl_b200.h"\nint main(void) { return 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", "-x", "c++", str(src)], check=True)


def test_null_handles_are_refused_with_a_message():
    """The reference asserts on misuse (scenarios/ScenarioTrain.cpp:282); across the ABI that is a non-zero return + trl_last_error."""
    import ctypes as C
    L = trl.load_library()
    out = C.c_int(0)
    for call in (lambda: L.trl_update(None, C.c_double(1.0 / 30.0)), lambda: L.trl_reset(None, None, 0), lambda: L.trl_reset_tuples(None),
                 lambda: L.trl_num_tuples(None, C.byref(out)), lambda: L.trl_set_explore(None, 1, C.c_double(0.1), C.c_double(0.1), C.c_double(0.1)),
                 lambda: L.trl_reset_avg_dist(None), lambda: L.trl_set_terrain_lerp(None, C.c_double(0.5))):
        assert call() != 0
        assert b"null handle" in L.trl_last_error()
    assert L.trl_destroy(None) == 0 and L.trl_trainer_destroy(None) == 0


def test_sm100a_cubin_and_no_cpu_fallback(assets):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", trl.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    if not _cuda_available():
        with pytest.raises(RuntimeError, match="GPU|CUDA"):
            trl.ScenarioPoliEval(os.path.join(assets, "dog_flat.trlpack"), 4)


def test_pack_cross_file_identities(assets):
    """the scale file's critic offset/scale are (-0.5, 2) (sim/BaseControllerMACE.cpp:108-109) and the
    actor offsets are minus the optimised parameters of the three gait files (sim/DogControllerMACE.cpp:93-99)."""
    p = read_pack(os.path.join(assets, "dog_slopes_mixed.trlpack"))
    off, sc = p["net_out_offset"], p["net_out_scale"]
    assert np.all(off[:3] == -0.5) and np.all(sc[:3] == 2.0)
    ctrl = p["ctrl_params"].reshape(-1, 30)
    for a in range(3):
        np.testing.assert_allclose(off[3 + 29 * a:3 + 29 * (a + 1)], -ctrl[a % 3][1:], atol=1e-6)
    dims = p["net_dims"]
    assert list(dims) == [283, 83, 90, 3, 29]
    n_params = sum(p[k].size for k in p if k.startswith("net_") and (k.endswith("_w") or k.endswith("_b")))
    assert n_params == 570474
    mi = p["meta_i32"]
    assert mi[2] == 20 and mi[3] == 5 and mi[8] == 21 and mi[9] == 23   # args/dog_slopes_mixed_args.txt:12-13
    bodies = p["bodies"].reshape(21, 9)
    assert abs(bodies[:, 1].sum() - 33.67) < 1e-9


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference tree not present on this box")
def test_packs_regenerate_from_reference(assets, tmp_path):
    from pack_scene import build_pack
    for arg, name in (("args/sim_dog_args.txt", "dog_flat.trlpack"), ("args/dog_slopes_mixed_args.txt", "dog_slopes_mixed.trlpack"),
                      ("args/goat_cliffs_args.txt", "goat_cliffs.trlpack"), ("args/raptor_narrow_gaps_args.txt", "raptor_narrow_gaps.trlpack")):
        rec = build_pack(os.path.join("/root/reference", arg), "/root/reference")
        old = read_pack(os.path.join(assets, name))
        assert set(rec) == set(old)
        for k in rec:
            np.testing.assert_array_equal(np.asarray(rec[k]).ravel(), old[k])


@pytest.mark.skipif(not os.path.isdir("/root/reference/data"), reason="reference tree not present on this box")
def test_native_loaders_match_python_packer(assets, tmp_path):
    """The C++ readers behind trl_create (arg file, JSON assets, Caffe HDF5 weights; csrc/ref_loader.*) produce the
    same named arrays, bit for bit, as tools/pack_scene.py -- including the 570,474 f64 policy weights."""
    for arg, name in (("args/sim_dog_args.txt", "dog_flat.trlpack"), ("args/dog_slopes_mixed_args.txt", "dog_slopes_mixed.trlpack"),
                      ("args/goat_cliffs_args.txt", "goat_cliffs.trlpack"), ("args/raptor_narrow_gaps_args.txt", "raptor_narrow_gaps.trlpack")):
        out = tmp_path / name
        trl.pack_from_args(["-arg_file=", arg], "/root/reference", out)
        a, b = read_pack(out), read_pack(os.path.join(assets, name))
        assert list(a) == list(b)
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    # CLI tokens win over the arg file (optimizer/Main.cpp:19-32)
    out = tmp_path / "override.trlpack"
    trl.pack_from_args(["-num_sim_substeps=", "2", "-arg_file=", "args/sim_dog_args.txt"], "/root/reference", out)
    assert read_pack(out)["meta_i32"][3] == 2
    with pytest.raises(RuntimeError):
        trl.pack_from_args(["-arg_file=", "args/does_not_exist.txt"], "/root/reference", tmp_path / "x.trlpack")


def test_raptor_pack_identities(assets):
    p = read_pack(os.path.join(assets, "raptor_narrow_gaps.trlpack"))
    assert list(p["net_dims"]) == [275, 75, 87, 3, 28]          # sim/NNController.cpp:49-78 size contract for the raptor
    assert p["meta_i32"][8] == 19 and p["meta_i32"][9] == 21 and p["meta_i32"][13] == 0 and p["meta_i32"][14] == 0
    ctrl = p["ctrl_params"].reshape(3, 37)
    mask = np.array([0, 1, 1, 0, 0] + [1, 0, 1, 1, 1, 1, 1, 1] * 2 + [0, 0, 1, 1, 1, 1, 1, 1] * 2, bool)
    assert mask.sum() == 28
    off = p["net_out_offset"]
    for a in range(3):
        np.testing.assert_allclose(off[3 + 28 * a:3 + 28 * (a + 1)], -ctrl[a][mask], atol=1e-6)   # BuildActorBias


def test_plain_c_client_builds_and_fails_loudly_without_gpu(tmp_path):
    """examples/eval_and_train.c: the header is valid C99, the library links from C, and without a GPU the first create call
    fails with a message instead of falling back to anything."""
    import subprocess
    import deepterrainrl_b200 as trl
    lib = trl.library_path()
    exe = str(tmp_path / "client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "eval_and_train.c"),
                    "-L", os.path.dirname(lib), "-lterrainrl_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe], check=True)
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the client would run the full example")
    r = subprocess.run([exe, os.path.join(ROOT, "assets", "dog_flat.trlpack"), "8"], capture_output=True, text=True)
    assert r.returncode == 2 and "no GPU visible" in r.stderr
