"""CPU suite, part 2: the C-ABI library loads without a GPU, exports every symbol include/dynprog_cuda.h
declares, mirrors the reference's argument checks, and refuses to solve anything without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gmap_gsnap_b200 import api

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
