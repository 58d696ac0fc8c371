"""TEST INFRASTRUCTURE ONLY: the MACE actor-critic forward pass written with torch's own f64 layers -- a second, independent
implementation of the published Caffe layer semantics of data/policies/dog/nets/dog_mace3_deploy.prototxt (Slice 200 | rest;
Convolution = cross-correlation, NCHW with H = 1 -> F.conv1d; InnerProduct over the C x W flattening -> F.linear; ReLU; Concat),
wrapped by cNeuralNet::Eval's normalisation (learning/NeuralNet.cpp:352-375,977-986,1027-1036).  It stands behind oracle/net.h
(tests/test_net_torch_cpu.py) and behind the decision kernel (tests/test_gpu_net_torch.py); nothing in the product imports it."""
import numpy as np
import torch
import torch.nn.functional as F

LAYERS = ["terr_conv0", "terr_conv1", "terr_conv2", "terr_ip0", "ip0", "val_ip0", "val_ip1",
          "a0_ip0", "a0_ip1", "a1_ip0", "a1_ip1", "a2_ip0", "a2_ip1"]


def blobs_from_pack(rec):
    """{layer: (w, b)} + the four offset / scale vectors from the records of a .trlpack (tools/pack_scene.read_pack)"""
    blobs = {n: (np.asarray(rec["net_" + n + "_w"], float).ravel(), np.asarray(rec["net_" + n + "_b"], float).ravel()) for n in LAYERS}
    vec = {k: np.asarray(rec[k], float) for k in ("net_in_offset", "net_in_scale", "net_out_offset", "net_out_scale")}
    return blobs, vec["net_in_offset"], vec["net_in_scale"], vec["net_out_offset"], vec["net_out_scale"]


def forward(blobs, in_off, in_scale, out_off, out_scale, x, activations=None):
    """x: [n, 200 + n_char] raw policy states -> [n, n_out] unnormalised outputs (what cNeuralNet::Eval returns).
    activations: optional dict that receives every blob of the deploy net under its prototxt name (pre-activation blobs and the
    ReLU layers' own top blobs; [n, ...] arrays) -- what cNeuralNet::GetLayerState returns for that name."""
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
    x = t(np.atleast_2d(x))
    n = x.shape[0]
    n_char = x.shape[1] - 200
    w = lambda name, shape: t(blobs[name][0]).reshape(shape)
    b = lambda name: t(blobs[name][1])
    xn = (x + t(in_off)) * t(in_scale)                                       # NormalizeInput
    terr, char = xn[:, :200].reshape(n, 1, 200), xn[:, 200:]                 # slice0 (axis 1, point 200)
    acts = {"data": xn, "data_terrain": terr, "data_char": char, "char_flatten0": char}
    acts["terr_conv0"] = F.conv1d(terr, w("terr_conv0", (16, 1, 8)), b("terr_conv0"))
    a0 = acts["terr_relu0"] = F.relu(acts["terr_conv0"])
    acts["terr_conv1"] = F.conv1d(a0, w("terr_conv1", (32, 16, 4)), b("terr_conv1"))
    a1 = acts["terr_relu1"] = F.relu(acts["terr_conv1"])
    acts["terr_conv2"] = F.conv1d(a1, w("terr_conv2", (32, 32, 4)), b("terr_conv2"))
    a2 = acts["terr_relu2"] = F.relu(acts["terr_conv2"])
    acts["terr_ip0"] = F.linear(a2.reshape(n, 32 * 187), w("terr_ip0", (64, 32 * 187)), b("terr_ip0"))
    tip = acts["terr_relu3"] = F.relu(acts["terr_ip0"])
    cat = acts["concat0"] = torch.cat([tip, char], dim=1)                    # concat0: terrain features first
    acts["ip0"] = F.linear(cat, w("ip0", (256, 64 + n_char)), b("ip0"))
    h = acts["relu0"] = F.relu(acts["ip0"])
    outs = []
    for head in ("val", "a0", "a1", "a2"):
        acts[head + "_ip0"] = F.linear(h, w(head + "_ip0", (128, 256)), b(head + "_ip0"))
        hh = acts[head + "_relu0"] = F.relu(acts[head + "_ip0"])
        nout = blobs[head + "_ip1"][1].size
        y = acts[head + "_ip1"] = F.linear(hh, w(head + "_ip1", (nout, 128)), b(head + "_ip1"))
        outs.append(y)
    yn = torch.cat(outs, dim=1)                                              # output: val | a0 | a1 | a2
    acts["output"] = yn
    y = yn / t(out_scale) - t(out_off)                                       # UnnormalizeOutput
    if activations is not None:
        activations.update({k: v.numpy() for k, v in acts.items()})
    return y.numpy()


def typical_inputs(in_off, in_scale, n, seed):
    """policy states drawn around the training distribution: x = N(0,1) / scale - offset per component"""
    rng = np.random.default_rng(seed)
    sc = np.where(in_scale == 0, 1.0, in_scale)
    return rng.normal(size=(n, in_off.size)) / sc - in_off
