"""TEST INFRASTRUCTURE ONLY: point the Python mirror (deepterrainrl_b200.scenario) at an emulator build of the library for the
duration of a test.  The product's own loader (scenario.load_library) knows nothing about this."""
import contextlib
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build as simt_build  # noqa: E402


def open_simt(defines=()):
    L = C.CDLL(simt_build.build(defines))
    L.trl_create_from_pack.restype = C.c_void_p
    L.trl_create_from_pack.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    L.trl_last_error.restype = C.c_char_p
    L.trl_kernel_launches.restype = C.c_int64
    L.trl_kernel_launches.argtypes = [C.c_void_p]
    L.simt_counter.restype = C.c_longlong
    return L


@contextlib.contextmanager
def simt_library(defines=()):
    from deepterrainrl_b200 import scenario
    saved = scenario._LIB
    scenario._LIB = open_simt(defines)
    try:
        yield scenario._LIB
    finally:
        scenario._LIB = saved
