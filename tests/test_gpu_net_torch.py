"""The decision kernel's forward pass (csrc/trl_decide.cuh) against torch's own f64 conv1d / linear on the same policy states
(tools/torch_net.py): an implementation that shares no code with the kernel or with oracle/net.h."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene,n_out", [("dog_slopes_mixed", 90), ("goat_cliffs", 90), ("raptor_narrow_gaps", 87)])
def test_decision_kernel_matches_torch(assets, scene, n_out):
    import deepterrainrl_b200 as trl
    import torch_net
    from pack_scene import read_pack
    path = os.path.join(assets, scene + ".trlpack")
    blobs, io, isc, oo, osc = torch_net.blobs_from_pack(read_pack(path))
    n = 96
    sc = trl.ScenarioPoliEval(path, n)
    for _ in range(30):                                   # 1 s: every env has taken at least one policy decision
        sc.Update()
    X = np.stack([sc.GetPoliState(e) for e in range(n)])
    Yg = np.stack([sc.GetNetOut(e, n_out) for e in range(n)])
    assert np.all(np.isfinite(X)) and np.abs(Yg).max() > 0
    Yt = torch_net.forward(blobs, io, isc, oo, osc, X)
    err = np.max(np.abs(Yg - Yt) / (1.0 + np.abs(Yt)))
    assert err <= 1e-10, err


def test_layer_state_probe_matches_torch(assets):
    """trl_get_layer_state (cNeuralNet::GetLayerState behind cScenarioPoliEval::RecordNNActivation): every blob of the deploy net for
    an env's last decision against torch's layers"""
    import deepterrainrl_b200 as trl
    import torch_net
    from pack_scene import read_pack
    path = os.path.join(assets, "dog_slopes_mixed.trlpack")
    blobs, io, isc, oo, osc = torch_net.blobs_from_pack(read_pack(path))
    sc = trl.ScenarioPoliEval(path, 8)
    for _ in range(30):
        sc.Update()
    acts = {}
    torch_net.forward(blobs, io, isc, oo, osc, sc.GetPoliState(5)[None], acts)
    assert len(acts) == 28
    for name, a in acts.items():
        g = sc.GetLayerState(name, 5)
        assert g.size == a.size, name
        assert np.max(np.abs(g - a.ravel()) / (1.0 + np.abs(a.ravel()))) <= 1e-12, name
    with pytest.raises(RuntimeError, match="Can't find layer"):
        sc.GetLayerState("no_such_layer", 0)
