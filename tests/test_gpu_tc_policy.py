"""The experimental tcgen05 split-precision kernel for the policy's wide inner product (csrc/trl_tc_policy.cu): the TMA /
shared-memory-descriptor / TMEM plumbing is right iff the tensor-core result equals an f64 product of the SAME narrow parts up to the
FP32 accumulation of 5984 x pairs terms.  (What the narrow arithmetic does to real decisions is tools/tc_policy_probe.py's subject.)"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,parts,pairs", [(0, 1, [(0, 0)]), (0, 3, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]), (1, 2, [(0, 0), (0, 1), (1, 0)])])
def test_tcgen05_product_equals_f64_product_of_the_parts(kind, parts, pairs):
    import torch
    import deepterrainrl_b200 as trl
    L = trl.load_library()
    L.trl_tc_last_error.restype = C.c_char_p
    dev = torch.device("cuda", 0)
    K, M = 5984, 200                                     # two 128-row tiles, the second one ragged
    g = torch.Generator(device=dev).manual_seed(5 + kind)
    A = torch.randn(M, K, dtype=torch.float64, device=dev, generator=g)
    B = torch.randn(64, K, dtype=torch.float64, device=dev, generator=g) * 0.05
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    err = torch.zeros(1, dtype=torch.int32, device=dev)

    def split(x):
        p = torch.empty((parts,) + tuple(x.shape), dtype=torch.int16 if kind == 0 else torch.float32, device=dev)
        assert L.trl_tc_split(C.c_void_p(x.data_ptr()), C.c_longlong(x.shape[0]), K, kind, parts, C.c_void_p(p.data_ptr()), st) == 0
        return p

    def as_f64(p):
        return (p.to(torch.int32) << 16).view(torch.float32).double() if kind == 0 else p.double()

    Ap, Bp = split(A), split(B)
    Af, Bf = as_f64(Ap), as_f64(Bp)
    # the parts reproduce the operand: bf16 x 3 / tf32 x 2 carry ~24 / ~22 bits
    assert (Af.sum(0) - A).abs().max().item() <= (2.0 ** -7 if parts == 1 else 2.0 ** -20) * A.abs().max().item()
    want = sum(Af[i] @ Bf[j].T for i, j in pairs)
    mask = sum(1 << (3 * i + j) for i, j in pairs)
    for ksplit in (1, 8):
        out = torch.zeros((M, 64), dtype=torch.float32, device=dev)
        rc = L.trl_tc_fc(C.c_void_p(Ap.data_ptr()), C.c_void_p(Bp.data_ptr()), M, K, kind, parts, mask, ksplit, C.c_void_p(out.data_ptr()),
                         C.c_void_p(err.data_ptr()), st)
        assert rc == 0, L.trl_tc_last_error().decode()
        torch.cuda.synchronize()
        assert int(err.item()) == 0, "a bounded barrier wait gave up inside trl_tc_fc_kernel"
        scale = max(1.0, want.abs().max().item())
        assert (out.double() - want).abs().max().item() <= 2e-4 * scale, (kind, parts, ksplit)
    # bad arguments are refused, not launched
    assert L.trl_tc_fc(C.c_void_p(Ap.data_ptr()), C.c_void_p(Bp.data_ptr()), M, K, 2, parts, mask, 1, C.c_void_p(out.data_ptr()), C.c_void_p(err.data_ptr()), st) != 0
