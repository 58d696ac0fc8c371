"""Scene packer: reference arg file + JSON assets + Caffe HDF5 weights -> one flat binary `.trlpack`.

The GPU box has no /root/reference, so everything a scene needs (skeleton, body boxes, PD gains, gait
controller parameter sets, initial state, terrain generator parameters, policy weights + scale vectors)
is baked into one file of named arrays that both the product library (csrc/host/scene_pack.h) and the
CPU oracle (oracle/pack_reader.h) read with their own ~60-line readers.

Semantics restated here (reference file:line):
  * arg files: `-key= value` tokens, `//` comments, first match wins   util/ArgParser.cpp:42-140
  * scenario defaults (20 update steps, 1 substep, scale 1)             scenarios/ScenarioSimChar.cpp:48-70
  * skeleton / body / PD tables                                         anim/KinTree.cpp:9-60,1000-1016; sim/PDController.cpp:7-80
  * gait controller files, action blends                                sim/DogController.cpp:470-712
  * terrain type + 40-entry parameter sets with defaults                sim/TerrainGen2D.cpp:8-70
  * MACE net blobs (Caffe ToHDF5 layout) + `_scale.txt`                 learning/NeuralNet.cpp:81-215,571-587

Format: magic b"TRLPACK1", u32 n_records, then per record:
  u32 name_len, name bytes, u32 dtype (0 = f64, 1 = i32), u64 count, raw little-endian data.

Usage: python tools/pack_scene.py <arg_file> <out.trlpack> [--ref-root /root/reference] [-key= value ...]
"""
import json
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from h5min import H5File  # noqa: E402

CHAR_NAMES = ["none", "dog", "raptor"]
CTRL_NAMES = ["none", "dog", "dog_cacla", "dog_mace", "goat_mace", "raptor", "raptor_cacla", "raptor_mace"]
TERRAIN_TYPES = ["flat", "gaps", "steps", "walls", "bumps", "mixed", "narrow_gaps", "slopes", "slopes_gaps",
                 "slopes_walls", "slopes_steps", "slopes_mixed", "slopes_narrow_gaps", "cliffs"]
TERRAIN_PARAM_DEFS = [
    ("GapSpacingMin", 4), ("GapSpacingMax", 7), ("GapWMin", 0.5), ("GapWMax", 2), ("GapHMin", -2), ("GapHMax", -2),
    ("WallSpacingMin", 6), ("WallSpacingMax", 8), ("WallWMin", 0.2), ("WallWMax", 0.2), ("WallHMin", 0.25),
    ("WallHMax", 0.5),
    ("StepSpacingMin", 5), ("StepSpacingMax", 7), ("StepH0Min", 0.1), ("StepH0Max", 0.4), ("StepH1Min", -0.4),
    ("StepH1Max", -0.1),
    ("BumpHMin", 0), ("BumpHMax", 0.03),
    ("NarrowGapSpacingMin", 3), ("NarrowGapSpacingMax", 6), ("NarrowGapDistMin", 0.1), ("NarrowGapDistMax", 0.4),
    ("NarrowGapWMin", 0.15), ("NarrowGapWMax", 0.5), ("NarrowGapDepthMin", -2), ("NarrowGapDepthMax", -2),
    ("NarrowGapCountMin", 1), ("NarrowGapCountMax", 4),
    ("CliffSpacingMin", 5), ("CliffSpacingMax", 7), ("CliffH0Min", 0.1), ("CliffH0Max", 0.4), ("CliffH1Min", -0.4),
    ("CliffH1Max", -0.1), ("CliffMiniCountMax", 0),
    ("SlopeDeltaRange", 0.25), ("SlopeDeltaMin", -0.35), ("SlopeDeltaMax", 0.35),
]
DOG_MISC = ["TransTime", "Cv", "BackForceX", "BackForceY", "FrontForceX", "FrontForceY"]
DOG_STATES = ["BackStance", "Extend", "FrontStance", "Gather"]
DOG_STATE_PARAMS = ["SpineCurve", "Shoulder", "Elbow", "Hip", "Knee", "Ankle"]
RAPTOR_MISC = ["TransTime", "Cv", "Cd", "ForceX", "ForceY"]
RAPTOR_STATES = ["Contact", "Down", "Passing", "Up"]
RAPTOR_STATE_PARAMS = ["RootPitch", "SpineCurve", "StanceHip", "StanceKnee", "StanceAnkle", "SwingHip", "SwingKnee",
                       "SwingAnkle"]
NET_LAYERS = ["terr_conv0", "terr_conv1", "terr_conv2", "terr_ip0", "ip0", "val_ip0", "val_ip1",
              "a0_ip0", "a0_ip1", "a1_ip0", "a1_ip1", "a2_ip0", "a2_ip1"]


def tokenize_arg_file(path):
    """cArgParser::AppendArgs(file): whitespace-separated tokens, `//` starts a comment to end of line."""
    toks = []
    with open(path, "r") as f:
        for line in f:
            i = line.find("//")
            if i >= 0:
                line = line[:i]
            toks.extend(line.split())
    return toks


class Args:
    def __init__(self, toks):
        self.toks = toks

    def get(self, key, default=None):
        k = "-" + key + "="
        for i, t in enumerate(self.toks):
            if t == k and i + 1 < len(self.toks):
                v = self.toks[i + 1]
                if not (v.startswith("-") and v.endswith("=")):
                    return v
                return default
        return default


def load_json(ref_root, rel):
    with open(os.path.join(ref_root, rel), "r") as f:
        return json.load(f)


def read_ctrl_params(ref_root, rel, misc, states, state_params):
    d = load_json(ref_root, rel)
    v = [float(d["MiscParams"][k]) for k in misc]
    for s in states:
        v.extend(float(d["StateParams"][s][k]) for k in state_params)
    return v


def build_pack(arg_file, ref_root, cli_toks=()):
    toks = list(cli_toks) + tokenize_arg_file(arg_file)  # CLI first => CLI wins (optimizer/Main.cpp:19-32)
    a = Args(toks)
    rec = {}

    char_file = a.get("character_file")
    state_file = a.get("state_file", "")
    char_type = CHAR_NAMES.index(a.get("char_type", "none"))
    ctrl = CTRL_NAMES.index(a.get("char_ctrl", "none"))
    ch = load_json(ref_root, char_file)

    joints = ch["Skeleton"]["Joints"]
    nj = len(joints)
    J = np.zeros((nj, 7))
    for i, j in enumerate(joints):
        J[i] = [j.get("Type", 0), j.get("Parent", -1), j.get("AttachX", 0), j.get("AttachY", 0), j.get("AttachZ", 0),
                j.get("LimLow", 1), j.get("LimHigh", 0)]
    # PostProcessJointMat zeroes the root attach point (anim/KinTree.cpp:1004-1016)
    J[0, 2:5] = 0
    ndof = int(sum(3 if int(t) in (1, 3) and int(p) < 0 else (3 if int(t) == 1 else (1 if int(t) in (0, 2) else 0))
                   for t, p in J[:, :2]))

    B = np.zeros((nj, 9))
    shapes = {"box": 0, "capsule": 1, "null": -1}
    for i, b in enumerate(ch["BodyDefs"]):
        B[i] = [shapes[b.get("Shape", "null")], b.get("Mass", 0), b.get("AttachX", 0), b.get("AttachY", 0),
                b.get("AttachZ", 0), b.get("Theta", 0), b.get("Param0", 0), b.get("Param1", 0), b.get("Param2", 0)]

    P = np.zeros((nj, 6))
    for i, p in enumerate(ch["PDControllers"]):
        P[i] = [p.get("Kp", 0), p.get("Kd", 0), p.get("TorqueLim", 0), p.get("TargetTheta", 0), p.get("TargetVel", 0),
                p.get("UseWorldCoord", 0)]

    cj = ch["Controllers"]
    is_raptor = char_type == 2
    misc, states, sparams = ((RAPTOR_MISC, RAPTOR_STATES, RAPTOR_STATE_PARAMS) if is_raptor
                             else (DOG_MISC, DOG_STATES, DOG_STATE_PARAMS))
    C = np.array([read_ctrl_params(ref_root, f, misc, states, sparams) for f in cj["Files"]])
    A = np.array([[x["ParamIdx0"], x["ParamIdx1"], x["Blend"], 1.0 if x["Cyclic"] else 0.0] for x in cj["Actions"]])
    default_action = int(cj.get("DefaultAction", 0))
    grav_comp = 1 if cj.get("EnableGravityCompensation", True) else 0
    virt_forces = 1 if cj.get("EnableVirtualForces", True) else 0

    if state_file:
        st = load_json(ref_root, state_file)
        pose0, vel0 = np.array(st["Pose"], float), np.array(st["Vel"], float)
    else:
        pose0, vel0 = np.zeros(ndof), np.zeros(ndof)
    assert pose0.size == ndof and vel0.size == ndof, (pose0.size, ndof)

    terrain_file = a.get("terrain_file", "")
    ttype, tparams = 0, np.zeros((0, len(TERRAIN_PARAM_DEFS)))
    if terrain_file:
        t = load_json(ref_root, terrain_file)
        ttype = TERRAIN_TYPES.index(t.get("Type", "flat") or "flat")
        sets = []
        for s in t.get("Params", []):
            sets.append([float(s.get(k, dflt)) for k, dflt in TERRAIN_PARAM_DEFS])
        tparams = np.array(sets, float).reshape(len(sets), len(TERRAIN_PARAM_DEFS))
    default_tparams = np.array([d for _, d in TERRAIN_PARAM_DEFS], float)

    init_x = a.get("char_init_pos_x")
    model = a.get("policy_model", "")
    has_net = 1 if (a.get("policy_net", "") and model) else 0

    rec["meta_i32"] = np.array([
        char_type, ctrl, int(a.get("num_update_steps", 20)), int(a.get("num_sim_substeps", 1)),
        1 if init_x is not None else 0, ttype, tparams.shape[0], has_net, nj, ndof, C.shape[0], A.shape[0],
        default_action, grav_comp, virt_forces, int(a.get("tuple_buffer_size", 16)),
    ], np.int32)
    rec["meta_f64"] = np.array([
        0.0, -9.8, float(init_x) if init_x is not None else 0.0, float(a.get("terrain_blend", 0)),
        float(a.get("exp_rate", 0.1)), float(a.get("exp_temp", 1)), float(a.get("exp_base_rate", 0.01)),
        float(a.get("world_scale", 1)),
    ])
    rec["joints"] = J
    rec["bodies"] = B
    rec["pd"] = P
    rec["ctrl_params"] = C
    rec["actions"] = A
    rec["pose0"] = pose0
    rec["vel0"] = vel0
    rec["terrain_params"] = tparams
    rec["terrain_default_params"] = default_tparams

    if has_net:
        ds = H5File(os.path.join(ref_root, model)).datasets()
        for name in NET_LAYERS:
            rec["net_" + name + "_w"] = ds["/data/" + name + "/0"]
            rec["net_" + name + "_b"] = ds["/data/" + name + "/1"]
        scale_path = os.path.splitext(os.path.join(ref_root, model))[0] + "_scale.txt"
        sc = json.load(open(scale_path))
        for k, n in (("InputOffset", "net_in_offset"), ("InputScale", "net_in_scale"),
                     ("OutputOffset", "net_out_offset"), ("OutputScale", "net_out_scale")):
            rec[n] = np.array(sc[k], float)
        n_in = rec["net_in_offset"].size
        n_out = rec["net_out_offset"].size
        frag = rec["net_a0_ip1_b"].size
        rec["net_dims"] = np.array([n_in, n_in - 200, n_out, n_out // (frag + 1), frag], np.int32)
    return rec


def write_pack(rec, path):
    with open(path, "wb") as f:
        f.write(b"TRLPACK1")
        f.write(struct.pack("<I", len(rec)))
        for name, arr in rec.items():
            arr = np.ascontiguousarray(arr)
            if arr.dtype == np.int32:
                dt = 1
            else:
                arr = arr.astype("<f8")
                dt = 0
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<IQ", dt, arr.size))
            f.write(arr.tobytes())


def read_pack(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"TRLPACK1"
        n, = struct.unpack("<I", f.read(4))
        for _ in range(n):
            ln, = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode()
            dt, cnt = struct.unpack("<IQ", f.read(12))
            if dt == 1:
                out[name] = np.frombuffer(f.read(4 * cnt), "<i4").copy()
            else:
                out[name] = np.frombuffer(f.read(8 * cnt), "<f8").copy()
    return out


if __name__ == "__main__":
    argv = sys.argv[1:]
    ref_root = "/root/reference"
    if "--ref-root" in argv:
        i = argv.index("--ref-root")
        ref_root = argv[i + 1]
        del argv[i:i + 2]
    arg_file, out = argv[0], argv[1]
    if not os.path.isabs(arg_file):
        arg_file = os.path.join(ref_root, arg_file)
    rec = build_pack(arg_file, ref_root, argv[2:])
    write_pack(rec, out)
    print(f"wrote {out}: {len(rec)} records, {os.path.getsize(out)} bytes")
