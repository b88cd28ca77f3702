"""CPU suite, part 2: the C-ABI library loads without a GPU, exports every symbol include/dynprog_cuda.h
declares, mirrors the reference's argument checks, and refuses to solve anything without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gmap_gsnap_b200 import api
from oracle import checkers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dynprog_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dpc_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = C.CDLL(os.path.join(ROOT, "gmap-gsnap_b200", "csrc", "libdynprog_cuda.so"))
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libdynprog_cuda.so does not export %s" % n


def test_struct_layouts_match_header():
    assert api.PROBLEM_DT.itemsize == 112
    assert api.RESULT_DT.itemsize == 72
    assert api.PAIR_DT.itemsize == 16
    assert api.PROBLEM_DT.fields["defect_rate"][1] == 104
    assert api.RESULT_DT.fields["left_prob"][1] == 56


def test_maxlengths_follow_compute_maxlengths():
    lib = api.CudaLib()
    lib.init()
    a, b = C.c_int(), C.c_int()
    lib.lib.dpc_maxlengths(C.byref(a), C.byref(b))
    assert (a.value, b.value) == (611, 2000)           # dynprog.c:831-852 with GMAP's defaults
    lib.init(maxlookback=1200)
    lib.lib.dpc_maxlengths(C.byref(a), C.byref(b))
    assert (a.value, b.value) == (1211, 2000)
    lib.init()


def test_no_cpu_fallback():
    """Without a usable device the library must fail loudly; it never solves on the host."""
    lib = api.CudaLib()
    lib.init()
    if lib.lib.dpc_device_count() > 0:
        pytest.skip("a GPU is present")
    w = api.Workload(100_000, seed=3)
    lib.setup(w.make_setup())
    assert lib.lib.dpc_ctx_new(0) is None
    with pytest.raises(RuntimeError):
        lib.open(0)
    probs = w.single_gaps(4)
    res = np.zeros(4, dtype=api.RESULT_DT)
    assert lib.lib.dpc_solve(None, probs.ctypes.data_as(C.c_void_p), 4, res.ctypes.data_as(C.c_void_p), None, 0, None) < 0
    assert b"no CPU fallback" in lib.lib.dpc_strerror(-1)


def test_python_package_refuses_without_library(tmp_path):
    with pytest.raises(RuntimeError):
        api.CudaLib(path=str(tmp_path / "missing.so"))


def test_argument_errors_mirror_the_reference_aborts():
    """The reference abort()s on non-positive matrix sizes (dynprog.c:495-498) and on an unknown Endalign_T
    (5215); the host side of the library (shared with tests/emul) returns an error code instead of solving."""
    w = api.Workload(200_000, seed=9)
    em = checkers.EmulLib()
    em.init()
    em.setup(w.make_setup())

    def code(p):
        res = np.zeros(len(p), dtype=api.RESULT_DT)
        off = np.zeros(len(p) + 1, dtype=np.int64)
        return em.lib.emul_solve(p.ctypes.data_as(C.c_void_p), len(p), res.ctypes.data_as(C.c_void_p), None, 0,
                                 off.ctypes.data_as(C.c_void_p))

    p = w.single_gaps(1)
    assert code(p) == 0
    q = p.copy(); q["length1"] = 0
    assert code(q) == -2                                    # DPC_ERR_ARG
    q = p.copy(); q["kind"] = 9
    assert code(q) == -2
    q = p.copy(); q["extraband"] = -1
    assert code(q) == -2
    q = p.copy(); q["chroffset"] = 0; q["chrhigh"] = 10**9; q["chrpos"] = 190_000; q["genomiclength"] = 60_000    # past the registered genome
    assert code(q) == -2
    e = w.end_gaps(1); e["endalign"] = 7
    assert code(e) == -2
    g = w.genome_gaps(1, finalp_mode=1)                     # finalp needs the MaxEnt hook, none registered here
    assert code(g) == -5                                    # DPC_ERR_STATE
    buf = np.frombuffer(bytes([200] * 64), dtype=np.uint8).copy()
    q = p.copy(); q["seq1"] = buf.ctypes.data; q["length1"] = 20
    assert code(q) == -3                                    # DPC_ERR_ALPHABET
