"""EXPERIMENT (measurement tool, not on the product path): what would a tcgen05 policy forward pass cost in decisions?

The reference evaluates its policy in f64 (Caffe Net<double>, learning/NeuralNet.h:13; cNeuralNet::Eval, learning/NeuralNet.cpp:352-375)
and the library's decision path keeps f64 (csrc/trl_decide2.cuh: TMA + DMMA).  tcgen05.mma has no f64 kind, so a tensor-core pass has
to split every operand into narrow parts.  This tool measures, on the states of real decisions:

  1. collects the policy states of >= N decisions from an exploration run of the library (tuple stream, s_beg of every tuple);
  2. evaluates the network in f64 with torch (tools/torch_net.py's layer semantics on the GPU) -> reference critic values / arg-max;
  3. evaluates the wide inner product terr_ip0 (5984 -> 64, 77 % of the network's multiply-adds) with csrc/trl_tc_policy.cu --
     tcgen05.mma, TMA-fed, FP32 accumulation in TMEM -- in bf16 x {1,2,3} and tf32 x {1,2} split precision, the rest of the network
     in f64, and counts how often the arg-max over the critics (= the actor whose action is applied) differs;
  4. checks the tensor-core kernel against an f64 product of the same narrow parts (validates descriptors / layout);
  5. times the kernel at decision-path batch sizes (M = 16 .. 128 rows, one k-split cluster of CTAs) and at throughput sizes.

Prints one JSON object (profiles/tc_policy_r02.json is a copy)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decisions", type=int, default=1_000_000)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--chunk", type=int, default=32768)
    ap.add_argument("--scene", default="dog_slopes_mixed")
    ap.add_argument("--exp-rate", type=float, default=0.2)
    ap.add_argument("--max-seconds", type=float, default=240.0)
    args = ap.parse_args()

    import torch
    import torch.nn.functional as F
    import deepterrainrl_b200 as trl
    import torch_net
    from pack_scene import read_pack

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", 0)
    pack = os.path.join(ROOT, "assets", args.scene + ".trlpack")
    blobs, in_off, in_scale, out_off, out_scale = torch_net.blobs_from_pack(read_pack(pack))
    L = trl.load_library()
    L.trl_tc_last_error.restype = C.c_char_p

    # ---------------------------------------------------------------- 1. real decision states
    t0 = time.time()
    sc = trl.ScenarioExpMACE(pack, args.envs, rng_seed=11)
    sc.EnableExplore(True, args.exp_rate, 0.025, 0.01)
    S = 200 + (in_off.size - 200)
    states = []
    have = 0
    for _ in range(60):
        sc.Update()
    sc.ResetTupleBuffer()
    updates = 0
    while have < args.decisions and time.time() - t0 < args.max_seconds:
        sc.Update()
        updates += 1
        if updates % 4 == 0:
            rows, _, _ = sc.GetTuples(f64=True)
            sc.ResetTupleBuffer()
            if rows.shape[0]:
                states.append(rows[:, 1:1 + S].copy())
                have += rows.shape[0]
    dr = C.c_int64(0)
    L.trl_tuples_dropped(sc.h, C.byref(dr))
    dropped = int(dr.value)
    sc.close()
    X = np.concatenate(states)[:args.decisions]
    n_dec = X.shape[0]
    collect_s = time.time() - t0

    # ---------------------------------------------------------------- network pieces on the GPU, f64
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    w = lambda name, shape: tt(blobs[name][0]).reshape(shape)
    b = lambda name: tt(blobs[name][1])
    Wc = [(w("terr_conv0", (16, 1, 8)), b("terr_conv0")), (w("terr_conv1", (32, 16, 4)), b("terr_conv1")), (w("terr_conv2", (32, 32, 4)), b("terr_conv2"))]
    Wt, bt = w("terr_ip0", (64, 5984)), b("terr_ip0")
    n_char = in_off.size - 200
    Wi, bi = w("ip0", (256, 64 + n_char)), b("ip0")
    heads = [(w(h + "_ip0", (128, 256)), b(h + "_ip0"), tt(blobs[h + "_ip1"][0]).reshape(-1, 128), b(h + "_ip1")) for h in ("val", "a0", "a1", "a2")]
    io_, is_, oo_, os_ = tt(in_off), tt(in_scale), tt(out_off), tt(out_scale)

    def conv_stage(x):                     # raw states -> (conv2 activations [n, 5984], normalised character features)
        xn = (x + io_) * is_
        a = xn[:, :200].reshape(-1, 1, 200)
        for (cw, cb) in Wc:
            a = F.relu(F.conv1d(a, cw, cb))
        return a.reshape(a.shape[0], 5984), xn[:, 200:]

    def tail(tip_pre, char):               # terr_ip0 pre-activation (bias included) -> unnormalised outputs
        h = F.relu(F.linear(torch.cat([F.relu(tip_pre), char], dim=1), Wi, bi))
        outs = [F.linear(F.relu(F.linear(h, w0, b0)), w1, b1) for (w0, b0, w1, b1) in heads]
        return torch.cat(outs, dim=1) / os_ - oo_

    # ---------------------------------------------------------------- the tensor-core kernel
    def ck(rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {L.trl_tc_last_error().decode()}")

    err_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def split(x64, kind, parts):           # f64 [rows, K] -> planes [parts, rows, K] (bf16 bits as int16 / tf32-rounded f32)
        rows, K = x64.shape
        planes = torch.empty((parts, rows, K), dtype=torch.int16 if kind == 0 else torch.float32, device=dev)
        ck(L.trl_tc_split(C.c_void_p(x64.data_ptr()), C.c_longlong(rows), K, kind, parts, C.c_void_p(planes.data_ptr()), st), "trl_tc_split")
        return planes

    def planes_f64(planes, kind):
        if kind == 0:
            return (planes.to(torch.int32) << 16).view(torch.float32).double()
        return planes.double()

    def tc_fc(a_planes, b_planes, kind, pairs, ksplit=1, out=None):
        parts, M, K = a_planes.shape
        if out is None:
            out = torch.zeros((M, 64), dtype=torch.float32, device=dev) if ksplit > 1 else torch.empty((M, 64), dtype=torch.float32, device=dev)
        ck(L.trl_tc_fc(C.c_void_p(a_planes.data_ptr()), C.c_void_p(b_planes.data_ptr()), M, K, kind, parts, pairs, ksplit,
                       C.c_void_p(out.data_ptr()), C.c_void_p(err_flag.data_ptr()), st), "trl_tc_fc")
        return out

    def pair_mask(pairs):
        return sum(1 << (3 * i + j) for i, j in pairs)

    MODES = [                               # name, kind (0 bf16, 1 tf32), parts, part pairs (i of A, j of B)
        ("bf16x1", 0, 1, [(0, 0)]),
        ("bf16x2_3prod", 0, 2, [(0, 0), (0, 1), (1, 0)]),
        ("bf16x3_6prod", 0, 3, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]),
        ("bf16x3_9prod", 0, 3, [(i, j) for i in range(3) for j in range(3)]),
        ("tf32x1", 1, 1, [(0, 0)]),
        ("tf32x2_3prod", 1, 2, [(0, 0), (0, 1), (1, 0)]),
        ("tf32x2_4prod", 1, 2, [(0, 0), (0, 1), (1, 0), (1, 1)]),
    ]
    wplanes = {(k, p): split(Wt.contiguous(), k, p) for k in (0, 1) for p in (1, 2, 3) if not (k == 1 and p == 3)}

    # ---------------------------------------------------------------- 4. kernel check on one chunk: tensor cores vs f64 product of the same parts
    res = {"scene": args.scene, "decisions": int(n_dec), "envs": args.envs, "updates": updates, "collect_s": round(collect_s, 1),
           "tuples_dropped": dropped, "exploration_rate": args.exp_rate, "modes": {}, "kernel_check": {}, "timing": {}}
    xs = tt(X[:4096])
    a2, _ = conv_stage(xs)
    for name, kind, parts, pairs in MODES:
        ap_ = split(a2, kind, parts)
        got = tc_fc(ap_, wplanes[(kind, parts)], kind, pair_mask(pairs)).double()
        Af, Bf = planes_f64(ap_, kind), planes_f64(wplanes[(kind, parts)], kind)
        want = sum(Af[i] @ Bf[j].T for i, j in pairs)
        got_k = tc_fc(ap_, wplanes[(kind, parts)], kind, pair_mask(pairs), ksplit=8).double()
        torch.cuda.synchronize()
        scale = want.abs().max().item()
        res["kernel_check"][name] = {"max_abs_err_vs_f64_product_of_parts": (got - want).abs().max().item(), "ksplit8": (got_k - want).abs().max().item(),
                                     "max_abs_value": scale}
    if int(err_flag.item()) != 0:
        res["error"] = "a bounded barrier wait gave up inside trl_tc_fc_kernel"
        print(json.dumps(res))
        return 1
    bad = [k for k, v in res["kernel_check"].items() if not (v["max_abs_err_vs_f64_product_of_parts"] <= 1e-3 * max(1.0, v["max_abs_value"]))]
    if bad:
        res["error"] = f"tensor-core result differs from the f64 product of the same parts: {bad}"
        print(json.dumps(res))
        return 1

    # ---------------------------------------------------------------- 2 + 3. flip rates over all decisions
    acc = {name: {"flips": 0, "max_abs_err_tip": 0.0, "max_abs_err_value": 0.0, "sum_abs_err_value": 0.0, "max_abs_err_action": 0.0} for name, *_ in MODES}
    acc["fp32_whole_net_torch"] = {"flips": 0, "max_abs_err_tip": 0.0, "max_abs_err_value": 0.0, "sum_abs_err_value": 0.0, "max_abs_err_action": 0.0}
    margins = []
    W32 = [(cw.float(), cb.float()) for cw, cb in Wc]
    for c0 in range(0, n_dec, args.chunk):
        x = tt(X[c0:c0 + args.chunk])
        a2, char = conv_stage(x)
        tip_ref = F.linear(a2, Wt, bt)
        y_ref = tail(tip_ref, char)
        v_ref = y_ref[:, :3]
        am_ref = v_ref.argmax(dim=1)
        srt = v_ref.sort(dim=1, descending=True).values
        margins.append((srt[:, 0] - srt[:, 1]).cpu().numpy())
        n_frag = (y_ref.shape[1] - 3) // 3

        def account(name, tip, y):
            a = acc[name]
            v = y[:, :3]
            am = v.argmax(dim=1)
            a["flips"] += int((am != am_ref).sum().item())
            if tip is not None:
                a["max_abs_err_tip"] = max(a["max_abs_err_tip"], (tip - tip_ref).abs().max().item())
            dv = (v - v_ref).abs()
            a["max_abs_err_value"] = max(a["max_abs_err_value"], dv.max().item())
            a["sum_abs_err_value"] += dv.sum().item()
            same = am == am_ref
            if same.any():
                idx = 3 + am_ref[same, None] * n_frag + torch.arange(n_frag, device=dev)[None]
                a["max_abs_err_action"] = max(a["max_abs_err_action"], (y[same].gather(1, idx) - y_ref[same].gather(1, idx)).abs().max().item())

        for name, kind, parts, pairs in MODES:
            ap_ = split(a2, kind, parts)
            tip = tc_fc(ap_, wplanes[(kind, parts)], kind, pair_mask(pairs)).double() + bt
            account(name, tip, tail(tip, char))
            del ap_
        # context: the whole network in FP32 FFMA (torch / cuBLAS, TF32 off)
        xn = ((x + io_) * is_).float()
        a = xn[:, :200].reshape(-1, 1, 200)
        for (cw, cb) in W32:
            a = F.relu(F.conv1d(a, cw, cb))
        tip32 = F.linear(a.reshape(a.shape[0], 5984), Wt.float(), bt.float())
        h = F.relu(F.linear(torch.cat([F.relu(tip32), xn[:, 200:]], dim=1), Wi.float(), bi.float()))
        outs = [F.linear(F.relu(F.linear(h, w0.float(), b0.float())), w1.float(), b1.float()) for (w0, b0, w1, b1) in heads]
        y32 = torch.cat(outs, dim=1).double() / os_ - oo_
        account("fp32_whole_net_torch", tip32.double(), y32)
    torch.cuda.synchronize()
    if int(err_flag.item()) != 0:
        res["error"] = "a bounded barrier wait gave up inside trl_tc_fc_kernel"
    m = np.concatenate(margins)
    res["critic_margin"] = {"median": float(np.median(m)), "p01": float(np.quantile(m, 0.01)), "p0001": float(np.quantile(m, 1e-4)), "min": float(m.min()),
                            "frac_below_1e-3": float((m < 1e-3).mean()), "frac_below_1e-5": float((m < 1e-5).mean()), "frac_below_1e-7": float((m < 1e-7).mean())}
    for name, a in acc.items():
        res["modes"][name] = {"argmax_flips": a["flips"], "flip_rate": a["flips"] / n_dec, "max_abs_err_terr_ip0": a["max_abs_err_tip"],
                              "max_abs_err_critic_value": a["max_abs_err_value"], "mean_abs_err_critic_value": a["sum_abs_err_value"] / (3 * n_dec),
                              "max_abs_err_selected_action_param": a["max_abs_err_action"]}

    # ---------------------------------------------------------------- 5. timing (CUDA events, L2-warm weights: the decision path's situation)
    def time_call(fn, reps=20):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3          # us

    a2, _ = conv_stage(tt(X[:args.chunk]))
    for name, kind, parts, pairs in (MODES[0], MODES[2], MODES[5]):
        for M, ks in ((128, 1), (128, 8), (128, 16), (128, 47), (1024, 1), (1024, 8), (args.chunk, 1)):
            M = min(M, a2.shape[0])
            ap_ = split(a2[:M].contiguous(), kind, parts)
            out = torch.zeros((M, 64), dtype=torch.float32, device=dev)
            us = time_call(lambda: tc_fc(ap_, wplanes[(kind, parts)], kind, pair_mask(pairs), ksplit=ks, out=out))
            flop = 2.0 * M * 64 * 5984 * len(pairs)
            res["timing"][f"{name}_M{M}_ksplit{ks}"] = {"us": round(us, 2), "tensor_tflops": round(flop / us * 1e-6, 2),
                                                         "f64_equivalent_gflops": round(2.0 * M * 64 * 5984 / us * 1e-3, 1)}
            del ap_
    # the f64 GEMM torch / cuBLAS runs for the same product (context for the throughput figure)
    for M in (128, args.chunk):
        xa = a2[:M].contiguous()
        us = time_call(lambda: F.linear(xa, Wt))
        res["timing"][f"f64_cublas_M{M}"] = {"us": round(us, 2), "f64_gflops": round(2.0 * M * 64 * 5984 / us * 1e-3, 1)}
    if int(err_flag.item()) != 0:
        res["error"] = "a bounded barrier wait gave up inside trl_tc_fc_kernel"
    res["total_s"] = round(time.time() - t0, 1)
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
