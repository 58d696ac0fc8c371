"""oracle/net.h (the builder's restatement of the Caffe forward pass) against an independent implementation of the same published
layer semantics: torch's own f64 conv1d / linear (tools/torch_net.py).  Caffe itself is absent, so this is the strongest pin the
NN arithmetic can get here: two implementations that share no code agree to rounding on every shipped model."""
import os

import numpy as np
import pytest

REF = "/root/reference"
SCENES = {"dog": "args/dog_slopes_mixed_args.txt", "goat": "args/goat_cliffs_args.txt", "raptor": "args/raptor_narrow_gaps_args.txt"}


def _check(pack_path, rec, tol=1e-12):
    from pyoracle import Oracle
    import torch_net
    blobs, io, isc, oo, osc = torch_net.blobs_from_pack(rec)
    o = Oracle(pack_path, 1, 0)
    n_out = oo.size
    X = torch_net.typical_inputs(io, isc, 24, 3)
    X[0] = -io                                       # the normalised zero vector
    X[1] = 0.0
    Y = torch_net.forward(blobs, io, isc, oo, osc, X)
    worst = 0.0
    for x, y in zip(X, Y):
        yo = o.net_eval(x, n_out)
        worst = max(worst, float(np.max(np.abs(yo - y) / (1.0 + np.abs(y)))))
    assert worst <= tol, worst
    return worst


@pytest.mark.parametrize("scene", ["dog_slopes_mixed", "goat_cliffs", "raptor_narrow_gaps"])
def test_oracle_net_matches_torch_on_the_asset_packs(assets, scene):
    from pack_scene import read_pack
    path = os.path.join(assets, scene + ".trlpack")
    _check(path, read_pack(path))


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout holds the other shipped models")
def test_oracle_net_matches_torch_on_all_shipped_models(tmp_path):
    """all 8 `.h5` models under data/policies (dog x4, goat x1, raptor x3), each packed with its character's scene"""
    import glob
    from pack_scene import build_pack, write_pack
    models = sorted(glob.glob(os.path.join(REF, "data/policies/*/models/*_model.h5")))
    assert len(models) == 8
    for m in models:
        char = os.path.basename(os.path.dirname(os.path.dirname(m)))
        rel = os.path.relpath(m, REF)
        rec = build_pack(os.path.join(REF, SCENES[char]), REF, ["-policy_model=", rel])
        p = str(tmp_path / (os.path.basename(m) + ".trlpack"))
        write_pack(rec, p)
        _check(p, rec)
