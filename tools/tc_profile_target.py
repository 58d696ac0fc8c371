"""Target for ncu: a few launches of the experimental tcgen05 terr_ip0 kernel (csrc/trl_tc_policy.cu) at throughput size and at the
decision path's size, on random operands.

  ncu --set full --clock-control none --import-source on -k regex:trl_tc_fc_kernel -c 4 -o gpurun_out/tc python tools/tc_profile_target.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import deepterrainrl_b200 as trl  # noqa: E402

L = trl.load_library()
L.trl_tc_last_error.restype = C.c_char_p
dev = torch.device("cuda", 0)
K = 5984
err = torch.zeros(1, dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
pairs6 = sum(1 << (3 * i + j) for i, j in [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)])


def planes(rows, parts):
    x = torch.randn(rows, K, dtype=torch.float64, device=dev)
    p = torch.empty((parts, rows, K), dtype=torch.int16, device=dev)
    assert L.trl_tc_split(C.c_void_p(x.data_ptr()), C.c_longlong(rows), K, 0, parts, C.c_void_p(p.data_ptr()), st) == 0
    return p


B = planes(64, 3)
for M, ks in ((32768, 1), (32768, 1), (128, 47), (128, 47)):       # launches 0/1: throughput size, 2/3: decision-path size with a k-split
    A = planes(M, 3)
    out = torch.zeros((M, 64), dtype=torch.float32, device=dev)
    rc = L.trl_tc_fc(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), M, K, 0, 3, pairs6, ks, C.c_void_p(out.data_ptr()), C.c_void_p(err.data_ptr()), st)
    assert rc == 0, L.trl_tc_last_error().decode()
    torch.cuda.synchronize()
    del A
print("bail flag", int(err.item()), "checksum", float(out.double().abs().sum().item()))
