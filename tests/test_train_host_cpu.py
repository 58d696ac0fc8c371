"""Host-side pieces of the training path that need no GPU: exploration / curriculum schedule (C ABI vs the reference formulas),
Caffe-layout HDF5 model write-back + scale file round trip, arg-file parsing, per-cycle record formats."""
import json
import os
import struct

import numpy as np
import pytest


def test_schedule_matches_reference_formulas():
    from deepterrainrl_b200.train import TrainSchedule
    # args/opt_args_train_mace.txt
    s = TrainSchedule(init_exp_rate=0.9, exp_rate=0.2, init_exp_temp=20, exp_temp=0.025, init_exp_base_rate=0.9, exp_base_rate=0.002,
                      trainer_num_anneal_iters=50000, exp_base_anneal_iters=50000, trainer_curriculum_iters=0)
    for it in (0, 1, 12500, 25000, 50000, 80000):
        lerp = min(max(it / 50000.0, 0.0), 1.0)
        out = s(it)
        assert out["exp_rate"] == pytest.approx((1 - lerp) * 0.9 + lerp * 0.2, abs=1e-15)
        assert out["exp_temp"] == pytest.approx((1 - lerp) * 20 + lerp * 0.025, abs=1e-13)
        assert out["exp_base_rate"] == pytest.approx((1 - lerp) * 0.9 + lerp * 0.002, abs=1e-15)
        assert out["curriculum_phase"] == 1.0                       # curriculum disabled -> phase 1
    c = TrainSchedule(trainer_curriculum_iters=1000)
    assert c(0)["curriculum_phase"] == 1.0                          # gInitCurriculumPhase at iteration 0
    assert c(250)["curriculum_phase"] == 0.25 and c(5000)["curriculum_phase"] == 1.0


def test_arg_file_parser(tmp_path):
    from deepterrainrl_b200.train import parse_arg_file, TrainSchedule
    p = tmp_path / "args.txt"
    p.write_text("-scenario= train_mace\n-exp_rate= 0.2 // comment\n-init_exp_temp= 20\n\n-trainer_num_anneal_iters= 50000\n")
    a = parse_arg_file(str(p))
    assert a["scenario"] == "train_mace" and a["exp_rate"] == "0.2"
    s = TrainSchedule.from_args(a)
    assert s.p["exp_rate"] == 0.2 and s.p["init_exp_temp"] == 20 and s.p["trainer_num_anneal_iters"] == 50000


def _pack_blobs(assets):
    from pack_scene import read_pack, NET_LAYERS
    p = read_pack(os.path.join(assets, "dog_slopes_mixed.trlpack"))
    shapes = {"terr_conv0": (16, 1, 1, 8), "terr_conv1": (32, 16, 1, 4), "terr_conv2": (32, 32, 1, 4), "terr_ip0": (64, 5984), "ip0": (256, 147)}
    blobs = {}
    for name in NET_LAYERS:
        w, b = p["net_" + name + "_w"], p["net_" + name + "_b"]
        blobs[name] = (w.reshape(shapes.get(name, (b.size, w.size // b.size))), b)
    return p, blobs


def test_model_write_back_round_trip(assets, tmp_path):
    from deepterrainrl_b200.model_io import H5File, MACE_LAYERS, read_model, write_model
    p, blobs = _pack_blobs(assets)
    path = str(tmp_path / "out" / "dog_model.h5")
    write_model(path, blobs, p["net_in_offset"], p["net_in_scale"], p["net_out_offset"], p["net_out_scale"], mtime=1459171656)
    layers, scale = read_model(path)
    assert sorted(layers) == sorted(blobs)
    for name, (w, b) in blobs.items():
        assert layers[name][0].shape == w.shape and layers[name][1].shape == b.shape
        np.testing.assert_array_equal(layers[name][0], w)
        np.testing.assert_array_equal(layers[name][1], b)
    np.testing.assert_array_equal(scale["InputOffset"], p["net_in_offset"])
    np.testing.assert_array_equal(scale["OutputScale"], p["net_out_scale"])
    assert list(json.load(open(path[:-3] + "_scale.txt"))) == ["InputOffset", "InputScale", "OutputOffset", "OutputScale"]
    raw = open(path, "rb").read()
    # superblock v0 as in the shipped files: sizes of offsets / lengths 8, group leaf K 4, internal K 16, EOF = file size
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and list(raw[8:16]) == [0, 0, 0, 0, 0, 8, 8, 0]
    assert struct.unpack_from("<HHI", raw, 16) == (4, 16, 0)
    base, free, eof, drv = struct.unpack_from("<4Q", raw, 24)
    assert base == 0 and free == drv == 0xFFFFFFFFFFFFFFFF and eof == len(raw)
    f = H5File(path)
    data = f._group_children(f.root["btree"], f.root["heap"])["data"]
    groups = f._group_children(data["btree"], data["heap"])
    assert sorted(groups) == sorted(MACE_LAYERS)                   # one group per layer, parameter-free ones empty
    assert f._group_children(groups["terr_relu0"]["btree"], groups["terr_relu0"]["heap"]) == {}
    # dataset object header: dataspace v1 with max dims, IEEE f64 LE, fill value v2, contiguous layout v3, mtime, NIL padding
    kids = f._group_children(groups["terr_conv0"]["btree"], groups["terr_conv0"]["heap"])
    types = [t for t, _, _ in f._messages(kids["0"]["ohdr"])]
    assert types == [0x1, 0x3, 0x5, 0x8, 0x12, 0x0]
    body = [f.b[b:b + s] for t, b, s in f._messages(kids["0"]["ohdr"]) if t == 0x3][0]
    assert body[:20].hex() == "11203f000800000000004000340b0034ff030000"


def test_native_writer_is_byte_identical(assets, tmp_path):
    """csrc/model_io.h (C ABI trl_write_model) and the Python writer produce the same bytes, HDF5 image and scale file."""
    from deepterrainrl_b200.model_io import write_model, write_model_native
    p, blobs = _pack_blobs(assets)
    a, b = str(tmp_path / "py" / "m.h5"), str(tmp_path / "cc" / "m.h5")
    args = (blobs, p["net_in_offset"], p["net_in_scale"], p["net_out_offset"], p["net_out_scale"])
    write_model(a, *args, mtime=1459171656)
    write_model_native(b, *args, mtime=1459171656)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert open(a[:-3] + "_scale.txt").read() == open(b[:-3] + "_scale.txt").read()


def test_cycle_recorder_formats(tmp_path):
    from deepterrainrl_b200.records import CycleRecorder
    S, A = 5, 3
    rec = CycleRecorder(2, S, A, str(tmp_path / "a.txt"), str(tmp_path / "s.txt"), str(tmp_path / "r.txt"))
    rows = np.arange(3 * (1 + 2 * S + A), dtype=float).reshape(3, -1) / 8
    rows[:, 1 + S] = [1, 2, 0]
    n = rec.consume(rows, np.zeros(3, np.uint32), np.array([2, 0, 2]))
    assert n == 2 and rec.cycles == 2
    a = open(tmp_path / "a.txt").read().splitlines()
    assert a[0] == "1,\t%f,\t%f" % (rows[0, 2 + S], rows[0, 3 + S]) and a[1].startswith("0,\t")
    s = open(tmp_path / "s.txt").read().splitlines()
    assert len(s) == 2 and s[0].split(",\t")[1] == "%f" % rows[0, 1]
    assert open(tmp_path / "r.txt").read().splitlines() == ["%f" % rows[0, 0], "%f" % rows[2, 0]]
