import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class RowOut(ctypes.Structure):
    """Mirror of nm_row_out (nanomod_b200/csrc/nm_lane.cuh)."""
    _fields_ = [("dnum", ctypes.c_int), ("ks_d", ctypes.c_double), ("ks_p", ctypes.c_double),
                ("two_u", ctypes.c_longlong), ("u_stat", ctypes.c_double), ("u_p", ctypes.c_double),
                ("t_stat", ctypes.c_double), ("t_p", ctypes.c_double), ("flags", ctypes.c_int)]


@pytest.fixture(scope="session", params=["float_imad", "float_keys", "int_keys"])
def emul(request):
    """g++ build of tests/host_emul/emul.cpp: the product's host/device headers compiled for the
    CPU so their logic can be checked without a GPU.  Test harness only.  Built three times: with
    -DNM_FLOAT_IMAD (the product's default: float32 keys, {fmin, a+b-min on the bit patterns}
    compare-exchanges), with plain float32 keys and with -DNM_INT_KEYS (int32 key images)."""
    src = os.path.join(ROOT, "tests", "host_emul", "emul.cpp")
    out_dir = os.path.join(ROOT, "tests", "host_emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "libemul_%s.so" % request.param)
    deps = [src] + [os.path.join(ROOT, "nanomod_b200", "csrc", f)
                    for f in ("nm_math.cuh", "nm_lane.cuh", "nm_deep.cuh", "nm_sortnet.inc", "nm_sortloop.inc")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        defs = {"int_keys": ["-DNM_INT_KEYS"], "float_imad": ["-DNM_FLOAT_IMAD"], "float_keys": []}[request.param]
        # hardware FMA for fmaf() where the host has it (the exhaustive grid-key test makes 3e10 of them);
        # no contraction of a*b+c anywhere else: results do not depend on the flag
        try:
            fma = ["-mfma", "-ffp-contract=off"] if " fma " in open("/proc/cpuinfo").read() else []
        except OSError:
            fma = []
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread"] + fma + defs +
                              ["-x", "c++", src, "-o", lib])
    L = ctypes.CDLL(lib)
    assert L.emul_sizeof_row_out() == ctypes.sizeof(RowOut)
    dbl = ctypes.c_double
    for name, args in (("emul_kolmogorov_sf", [dbl]), ("emul_student_t_two_sided", [dbl, dbl]),
                       ("emul_chi2_sf_even", [dbl, ctypes.c_int]), ("emul_norm_sf", [dbl]),
                       ("emul_norm_isf", [dbl])):
        getattr(L, name).restype = dbl
        getattr(L, name).argtypes = args
    L.emul_combine.restype = None
    L.emul_grid_exhaustive.restype = ctypes.c_longlong
    L.emul_grid_exhaustive.argtypes = [ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    L.emul_lane_position_grid.restype = ctypes.c_int
    L.emul_lane_position_grid.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.POINTER(ctypes.c_float),
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.variant = request.param
    return L


def rel_err(x, y):
    import math
    if x == y or (math.isnan(x) and math.isnan(y)):
        return 0.0
    return abs(x - y) / max(abs(y), 1e-300)
