"""Model I/O parity with the reference: Caffe `ToHDF5` weight files + `<stem>_scale.txt`.

cNeuralNet::OutputModel (learning/NeuralNet.cpp:571-587) writes the solver net with caffe::Net::ToHDF5 -- HDF5 superblock v0,
old-style groups (/data/<layer name>/<blob index>), contiguous IEEE f64 datasets, one (possibly empty) group per layer of the
net -- and the offset/scale vectors as JSON (WriteOffsetScale, learning/NeuralNet.cpp:1182-1205).  h5py / libhdf5 are not in this
image, so both directions are implemented on the file format itself: `H5File` reads what the reference ships
(data/policies/*/models/*.h5), `write_model` emits the same object layout (same message sets, group B-tree / heap / symbol-node
structure and sizes as the shipped files) so the original viewer's Caffe can load policies trained here.
"""
import json
import os
import struct
import time

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        assert b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        ver = b[8]
        assert ver == 0, f"superblock v{ver} unsupported"
        self.so, self.sl = b[13], b[14]
        assert self.so == 8 and self.sl == 8
        # v0 superblock: 8 sig, 8 version bytes, 2+2 group K, 4 flags, then 4 addresses
        p = 24
        self.base, _, self.eof, _ = struct.unpack_from("<4Q", b, p)
        p += 32
        # root group symbol table entry
        self.root = self._ste(p)

    def _ste(self, p):
        name_off, ohdr, cache_type, _ = struct.unpack_from("<QQII", self.b, p)
        scratch = self.b[p + 24:p + 40]
        ent = {"name_off": name_off, "ohdr": ohdr, "cache": cache_type}
        if cache_type == 1:
            ent["btree"], ent["heap"] = struct.unpack_from("<QQ", scratch, 0)
        return ent

    def _heap_data(self, addr):
        assert self.b[addr:addr + 4] == b"HEAP"
        size, _, data_addr = struct.unpack_from("<QQQ", self.b, addr + 8)
        return data_addr

    def _name(self, heap_data, off):
        e = self.b.index(b"\0", heap_data + off)
        return self.b[heap_data + off:e].decode()

    def _btree_leaves(self, addr, out):
        assert self.b[addr:addr + 4] == b"TREE", self.b[addr:addr + 4]
        ntype, level, nent = struct.unpack_from("<BBH", self.b, addr + 4)
        assert ntype == 0
        p = addr + 8 + 16  # siblings
        p += 8  # key 0
        for _ in range(nent):
            child, = struct.unpack_from("<Q", self.b, p)
            p += 16  # child + next key
            if level > 0:
                self._btree_leaves(child, out)
            else:
                out.append(child)

    def _messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1
        msgs = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(msgs) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x10:  # continuation
                    off, ln = struct.unpack_from("<QQ", b, body)
                    blocks.append((off, ln))
                msgs.append((mtype, body, msize))
                p = body + msize
        return msgs

    def _group_children(self, btree, heap):
        hd = self._heap_data(heap)
        leaves = []
        self._btree_leaves(btree, leaves)
        out = {}
        for snod in leaves:
            assert self.b[snod:snod + 4] == b"SNOD"
            n, = struct.unpack_from("<H", self.b, snod + 6)
            for i in range(n):
                ent = self._ste(snod + 8 + 40 * i)
                out[self._name(hd, ent["name_off"])] = ent
        return out

    def _read_dataset(self, ohdr):
        b = self.b
        dims, addr, size = None, None, None
        for mtype, body, msize in self._messages(ohdr):
            if mtype == 0x1:
                ver, rank, flags = struct.unpack_from("<BBB", b, body)
                p = body + (8 if ver == 1 else 4)
                dims = struct.unpack_from(f"<{rank}Q", b, p)
            elif mtype == 0x3:
                cls = b[body] & 0x0F
                tsize, = struct.unpack_from("<I", b, body + 4)
                assert cls == 1 and tsize == 8, "only float64 datasets supported"
            elif mtype == 0x8:
                ver = b[body]
                assert ver == 3, f"layout v{ver}"
                lclass = b[body + 1]
                assert lclass == 1, "only contiguous layout supported"
                addr, size = struct.unpack_from("<QQ", b, body + 2)
        n = int(np.prod(dims)) if dims else 1
        assert size == 8 * n
        return np.frombuffer(b, dtype="<f8", count=n, offset=addr).reshape(dims).copy()

    def _walk(self, ent, prefix, out):
        if "btree" not in ent:
            # group object header may carry the symbol-table message instead of cached scratch
            for mtype, body, msize in self._messages(ent["ohdr"]):
                if mtype == 0x11:
                    ent["btree"], ent["heap"] = struct.unpack_from("<QQ", self.b, body)
        if "btree" in ent:
            for name, ch in self._group_children(ent["btree"], ent["heap"]).items():
                self._walk(ch, prefix + "/" + name, out)
        else:
            out[prefix] = self._read_dataset(ent["ohdr"])

    def datasets(self):
        out = {}
        self._walk(self.root, "", out)
        return out


# ------------------------------------------------------------------------------------------------------------ writer
# layers of data/policies/*/nets/*_mace3_deploy.prototxt in net order (ToHDF5 creates one group per layer, empty if no blobs)
MACE_LAYERS = ["slice0", "terr_conv0", "terr_relu0", "terr_conv1", "terr_relu1", "terr_conv2", "terr_relu2", "terr_ip0",
               "terr_relu3", "char_flatten0", "concat0", "ip0", "relu0", "relu0_relu0_0_split", "val_ip0", "val_relu0", "val_ip1",
               "a0_ip0", "a0_relu0", "a0_ip1", "a1_ip0", "a1_relu0", "a1_ip1", "a2_ip0", "a2_relu0", "a2_ip1", "output"]
_LEAF_K, _INT_K = 4, 16                                  # group B-tree ranks of the shipped files (superblock bytes 16..19)
_F64_TYPE = bytes.fromhex("11203f000800000000004000340b0034ff030000") + b"\0" * 4      # datatype message body, IEEE f64 LE


class _Image:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, n, align=8):
        while len(self.b) % align:
            self.b.append(0)
        a = len(self.b)
        self.b.extend(b"\0" * n)
        return a

    def put(self, addr, data):
        self.b[addr:addr + len(data)] = data


def _msg(mtype, body, flags=0):
    assert len(body) % 8 == 0
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(img, msgs, total=None):
    body = b"".join(msgs)
    if total is not None:                                # pad with a NIL message like the library does (fixed-size header)
        pad = total - len(body) - 8
        assert pad >= 0
        body += _msg(0x0, b"\0" * pad)
        nmsg = len(msgs) + 1
    else:
        nmsg = len(msgs)
    addr = img.alloc(16 + len(body))
    img.put(addr, struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(body)) + body)
    return addr


def _dataset(img, arr, mtime):
    arr = np.ascontiguousarray(arr, "<f8")
    rank = arr.ndim
    dims = struct.pack(f"<{rank}Q", *arr.shape)
    space = struct.pack("<BBB5x", 1, rank, 1) + dims + dims          # v1 dataspace, max dims present
    data_addr = img.alloc(arr.nbytes)
    img.put(data_addr, arr.tobytes())
    msgs = [_msg(0x1, space), _msg(0x3, _F64_TYPE, 1), _msg(0x5, bytes.fromhex("0202020100000000"), 1),
            _msg(0x8, struct.pack("<BBQQ6x", 3, 1, data_addr, arr.nbytes), 1), _msg(0x12, struct.pack("<B3xI", 1, mtime))]
    return _object_header(img, msgs, total=256)


def _group(img, children):
    """children: {name: (object header address, is_group, btree, heap)} -> (ohdr, btree, heap) of a new old-style group."""
    names = sorted(children)                              # symbol nodes hold their entries in strcmp order
    # local heap: offset 0 is the empty string, names 8-byte aligned, the rest one free block
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        raw = n.encode() + b"\0"
        heap_data += raw + b"\0" * (-len(raw) % 8)
    used = len(heap_data)
    size = max(88, used + 16 + (-(used + 16) % 8))        # room for one free-list block (next = 1: none, size)
    heap_data += struct.pack("<QQ", 1, size - used) + b"\0" * (size - used - 16)
    cap = 2 * _LEAF_K
    chunks = [names[i:i + cap] for i in range(0, len(names), cap)] or [[]]
    snods = []
    for ch in chunks:
        a = img.alloc(8 + 40 * cap)
        ent = b""
        for n in ch:
            oh, is_group, bt, hp = children[n]
            scratch = struct.pack("<QQ", bt, hp) if is_group else b"\0" * 16
            ent += struct.pack("<QQII", name_off[n], oh, 1 if is_group else 0, 0) + scratch
        img.put(a, b"SNOD" + struct.pack("<BBH", 1, 0, len(ch)) + ent)
        snods.append(a)
    assert len(snods) <= 2 * _INT_K
    bt = img.alloc(24 + 2 * _INT_K * 16 + 8)
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods) if names else 0, UNDEF, UNDEF) + struct.pack("<Q", 0)
    if names:
        for a, ch in zip(snods, chunks):
            node += struct.pack("<QQ", a, name_off[ch[-1]])          # child, then the key = heap offset of its largest name
    img.put(bt, node)
    hp = img.alloc(32)
    hd = img.alloc(size)
    img.put(hd, bytes(heap_data))
    img.put(hp, b"HEAP" + struct.pack("<B3xQQQ", 0, size, used, hd))
    oh = _object_header(img, [_msg(0x11, struct.pack("<QQ", bt, hp))])
    return oh, bt, hp


def write_model(path, blobs, in_off, in_scale, out_off, out_scale, layers=None, mtime=None):
    """blobs: {layer name: (weight array in Caffe blob shape, bias array)}; writes `path` (HDF5) and `<stem>_scale.txt`."""
    layers = layers or MACE_LAYERS
    mtime = int(time.time()) if mtime is None else mtime
    img = _Image()
    img.alloc(96)                                         # superblock v0 + root symbol-table entry, filled in last
    layer_groups = {}
    for name in layers:
        kids = {}
        if name in blobs:
            for idx, arr in enumerate(blobs[name]):
                kids[str(idx)] = (_dataset(img, arr, mtime), False, 0, 0)
        oh, bt, hp = _group(img, kids)
        layer_groups[name] = (oh, True, bt, hp)
    missing = set(blobs) - set(layers)
    assert not missing, f"layers without a group: {missing}"
    d_oh, d_bt, d_hp = _group(img, layer_groups)
    r_oh, r_bt, r_hp = _group(img, {"data": (d_oh, True, d_bt, d_hp)})
    eof = len(img.b) + (-len(img.b) % 8)
    img.b.extend(b"\0" * (eof - len(img.b)))
    sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", _LEAF_K, _INT_K, 0)
    sb += struct.pack("<4Q", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, r_oh, 1, 0) + struct.pack("<QQ", r_bt, r_hp)
    assert len(sb) == 96
    img.put(0, sb)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(bytes(img.b))
    write_scale(os.path.splitext(path)[0] + "_scale.txt", in_off, in_scale, out_off, out_scale)


def write_scale(path, in_off, in_scale, out_off, out_scale):
    """cNeuralNet::WriteOffsetScale (learning/NeuralNet.cpp:1182-1205): same keys and key order; %.17g numbers (round-trip exact)."""
    def vec(v):
        return "[" + ", ".join("%.17g" % float(x) for x in np.asarray(v).ravel()) + "]"
    with open(path, "w") as f:
        f.write("{\n\"InputOffset\": %s,\n\"InputScale\": %s,\n\"OutputOffset\": %s,\n\"OutputScale\": %s\n}" %
                (vec(in_off), vec(in_scale), vec(out_off), vec(out_scale)))


def read_model(path):
    """-> ({layer: [blob arrays]}, {InputOffset, InputScale, OutputOffset, OutputScale} or None)"""
    ds = H5File(path).datasets()
    layers = {}
    for k in sorted(ds):
        _, root, layer, idx = k.split("/")
        assert root == "data"
        layers.setdefault(layer, {})[int(idx)] = ds[k]
    out = {name: [d[i] for i in sorted(d)] for name, d in layers.items()}
    sp = os.path.splitext(path)[0] + "_scale.txt"
    scale = None
    if os.path.exists(sp):
        with open(sp) as f:
            scale = {k: np.array(v, float) for k, v in json.load(f).items()}
    return out, scale


def write_model_native(path, blobs, in_off, in_scale, out_off, out_scale, mtime=0):
    """The same file through the C ABI's native writer (csrc/model_io.h): blobs {layer: (w, b)} of the 13 parameter layers."""
    import ctypes as C
    from .scenario import load_library
    L = load_library()
    order = ["terr_conv0", "terr_conv1", "terr_conv2", "terr_ip0", "ip0", "val_ip0", "val_ip1", "a0_ip0", "a0_ip1", "a1_ip0", "a1_ip1",
             "a2_ip0", "a2_ip1"]
    arrs = []
    for name in order:
        arrs += [np.ascontiguousarray(blobs[name][0], np.float64), np.ascontiguousarray(blobs[name][1], np.float64)]
    ptrs = (C.c_void_p * 26)(*[a.ctypes.data_as(C.c_void_p) for a in arrs])
    n_char = blobs["ip0"][0].shape[1] - 64
    n_frags, frag = blobs["val_ip1"][0].shape[0], blobs["a0_ip1"][0].shape[0]
    vec = [np.ascontiguousarray(v, np.float64) for v in (in_off, in_scale, out_off, out_scale)]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    rc = L.trl_write_model(path.encode(), ptrs, n_char, n_frags, frag, *[v.ctypes.data_as(C.c_void_p) for v in vec], C.c_uint32(mtime))
    if rc != 0:
        raise RuntimeError(L.trl_last_error().decode())
