"""The drop-in boundary on hardware: a plain-C client of include/terrainrl_b200.h (examples/eval_and_train.c -- what a maintainer's
adapter does, without Python) compiled with gcc, linked against the CUDA library and RUN on the GPU: evaluation of the shipped
policy, then the exploration + on-device trainer loop, then cNeuralNet::OutputModel.  (The C++ adapter over the reference's own
compiled scenario classes needs the reference's data tree, which exists only in the build container: tests/test_ref_adapter_cpu.py
runs it there over the emulator build of the same sources.)"""
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_client_runs_on_the_gpu(assets, tmp_path):
    import deepterrainrl_b200 as trl
    lib = trl.library_path()
    exe = str(tmp_path / "eval_and_train")
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "eval_and_train.c"),
                    "-L", os.path.dirname(lib), "-lterrainrl_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe], check=True)
    r = subprocess.run([exe, os.path.join(assets, "dog_slopes_mixed.trlpack"), "512"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"eval: (\d+) env-steps, (\d+) cycles, (\d+) episodes, avg distance ([-\d.]+) m", r.stdout)
    assert m, r.stdout
    assert int(m.group(1)) == 512 * 300 * 20 and int(m.group(2)) > 512 * 10
    t = re.search(r"train: iter (\d+), (\d+) tuples in the replay memory, critic loss ([-\d.eE+naninf]+)", r.stdout)
    assert t, r.stdout
    assert int(t.group(1)) > 0 and int(t.group(2)) > 2000 and float(t.group(3)) == float(t.group(3))
    from deepterrainrl_b200 import model_io
    layers, scale = model_io.read_model("/tmp/terrainrl_b200_model.h5")
    assert len(layers) == 13 and scale is not None and layers["terr_ip0"][0].size == 64 * 5984
