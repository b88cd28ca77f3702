"""oracle/checkers.py -- TEST INFRASTRUCTURE: ctypes drivers of the CPU checkers.

  PortOracle   oracle/liboracle_port.so          plain-C restatement of the reference's algorithm
  RefOracle    oracle/_ref/libdynprog_ref.so     the unmodified reference dynprog.c, compiled where it lies
  EmulLib      tests/emul/libdpc_emul.so         the device routines compiled for the CPU (lock-step lanes)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs import this module; the
product package (gmap-gsnap_b200/) never does.  Each checker library keeps its state (tables, genome, hooks) in
process-wide globals, so an instance re-registers its own init/setup arguments whenever another instance of the same
library was used in between.
"""
import ctypes as C
import os

import numpy as np

from gmap_gsnap_b200.api import (PROB_FN, RESULT_DT, Setup, _SolverLib, _ptr, _solve_common)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Checker(_SolverLib):
    _owner = {}          # library path -> the instance whose init/setup is currently registered

    def _claim(self):
        if _Checker._owner.get(self.path) is not self:
            _Checker._owner[self.path] = self
            if getattr(self, "_init_kw", None) is not None:
                self._do_init(**self._init_kw)
            if getattr(self, "_setup", None) is not None:
                self._do_setup(self._setup)

    def init(self, mode=0, maxlookback=600, extraquerygap=10, maxpeelback=11, end=10, paired=8):
        self._init_kw = dict(mode=mode, maxlookback=maxlookback, extraquerygap=extraquerygap, maxpeelback=maxpeelback, end=end, paired=paired)
        _Checker._owner[self.path] = self
        self._do_init(**self._init_kw)

    def setup(self, setup):
        self._setup = setup
        self._claim()
        self._do_setup(setup)


class PortOracle(_Checker):
    """oracle/liboracle_port.so -- TEST INFRASTRUCTURE (CPU restatement)."""

    def __init__(self, path=None):
        super().__init__(path or os.environ.get("DPC_PORT_LIB") or os.path.join(ROOT, "oracle", "liboracle_port.so"))
        L = self.lib
        L.port_init.argtypes = [C.c_int] * 6
        L.port_setup.argtypes = [C.POINTER(Setup)]
        L.port_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.port_pairdistance.argtypes = [C.c_int] * 3

    def _do_init(self, mode, maxlookback, extraquerygap, maxpeelback, end, paired):
        self.lib.port_init(maxlookback, extraquerygap, maxpeelback, end, paired, mode)

    def _do_setup(self, setup):
        self.lib.port_setup(C.byref(setup))

    def solve(self, problems, want_pairs=True):
        self._claim()
        return _solve_common(self.lib.port_solve, problems, want_pairs)


class RefOracle(_Checker):
    """oracle/_ref/libdynprog_ref.so -- the compiled, unmodified reference (TEST INFRASTRUCTURE)."""

    def __init__(self, path=None):
        super().__init__(path or os.path.join(ROOT, "oracle", "_ref", "libdynprog_ref.so"))
        L = self.lib
        L.ref_init.argtypes = [C.c_int] * 6
        L.ref_setup.argtypes = [C.POINTER(Setup)]
        L.ref_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.ref_solve_mt.restype = C.c_double
        L.ref_solve_mt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_pairdistance.argtypes = [C.c_int] * 2
        L.ref_splice_prob.restype = C.c_double
        self.splice_prob = C.cast(L.ref_splice_prob, PROB_FN)

    def _do_init(self, mode, maxlookback, extraquerygap, maxpeelback, end, paired):
        self.lib.ref_init(maxlookback, extraquerygap, maxpeelback, end, paired, mode)

    def _do_setup(self, setup):
        self.lib.ref_setup(C.byref(setup))

    def solve(self, problems, want_pairs=True):
        self._claim()
        return _solve_common(self.lib.ref_solve, problems, want_pairs)

    def solve_mt(self, problems, nthreads):
        self._claim()
        results = np.zeros(len(problems), dtype=RESULT_DT)
        secs = self.lib.ref_solve_mt(_ptr(problems), len(problems), _ptr(results), nthreads)
        return results, secs


class EmulLib(_Checker):
    """tests/emul/libdpc_emul.so -- TEST SCAFFOLDING: the device routines compiled single-lane for the CPU."""

    def __init__(self, path=None):
        super().__init__(path or os.environ.get("DPC_EMUL_LIB") or os.path.join(ROOT, "tests", "emul", "libdpc_emul.so"))
        L = self.lib
        L.emul_init.argtypes = [C.c_int] * 6
        L.emul_setup.argtypes = [C.POINTER(Setup)]
        L.emul_solve.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.emul_pairdistance.argtypes = [C.c_int] * 3
        L.emul_set_fill.argtypes = [C.c_int]
        L.emul_set_path.argtypes = [C.c_int]

    def set_fill(self, force_generic):
        """0: row-sweep fill + lane-parallel walk on 32 simulated lanes; 1: memory-state fill + serial walk."""
        self.lib.emul_set_fill(int(force_generic))

    def set_path(self, pipe):
        """0: host half (Batch packing / finalisation / rebuild); 1: the device pipeline's per-problem routines."""
        self.lib.emul_set_path(int(pipe))

    def _do_init(self, mode, maxlookback, extraquerygap, maxpeelback, end, paired):
        rc = self.lib.emul_init(maxlookback, extraquerygap, maxpeelback, end, paired, mode)
        if rc != 0:
            raise RuntimeError("emul_init failed: %d" % rc)

    def _do_setup(self, setup):
        self.lib.emul_setup(C.byref(setup))

    def solve(self, problems, want_pairs=True):
        self._claim()
        return _solve_common(self.lib.emul_solve, problems, want_pairs)


def arm_probability_mode(problems, solver):
    """Genome gaps the generator marked use_probabilities_p == 2 become the reference's second call
    (stage3.c:5833): use_probabilities_p = true with score_threshold = first-pass finalscore + QOPEN + 3*QINDEL
    (scores.h:7-8) = finalscore - 11.  `solver` is any of the libraries above."""
    out = problems.copy()
    sel = np.nonzero(out["use_probabilities_p"] == 2)[0]
    if len(sel) == 0:
        return out
    out["use_probabilities_p"][sel] = 0
    res, _, _ = solver.solve(out[sel], want_pairs=False)
    out["use_probabilities_p"][sel] = 1
    out["score_threshold"][sel] = res["finalscore"] - 11
    return out
