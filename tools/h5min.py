"""Minimal HDF5 reader used by tools/pack_scene.py -- the implementation lives in deepterrainrl_b200/model_io.py (which also
holds the writer); this module keeps the old import name."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deepterrainrl_b200.model_io import H5File, UNDEF  # noqa: E402,F401

if __name__ == "__main__":
    d = H5File(sys.argv[1]).datasets()
    tot = 0
    for k in sorted(d):
        print(k, d[k].shape, float(d[k].ravel()[0]))
        tot += d[k].size
    print("total params", tot)
