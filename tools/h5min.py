"""Minimal HDF5 reader (superblock v0, old-style groups, contiguous float64 datasets).

Enough to read the Caffe `ToHDF5` weight files the reference ships
(/root/reference/data/policies/*/models/*.h5; written by learning/NeuralNet.cpp:571-587).
h5py is not installed in this image, so the pack tool uses this instead.
Returns {"/data/<layer>/<idx>": np.ndarray(float64)}.
"""
import struct
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


if __name__ == "__main__":
    import sys
    d = H5File(sys.argv[1]).datasets()
    tot = 0
    for k in sorted(d):
        print(k, d[k].shape, float(d[k].ravel()[0]))
        tot += d[k].size
    print("total params", tot)
