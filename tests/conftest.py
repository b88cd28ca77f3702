"""Shared fixtures.  `-m "not gpu"` runs here (no GPU): oracle vs reference vs golden vectors, host logic,
the single-lane emulation of the device routines, and the C-ABI symbol check.  `-m gpu` are the parity
tests proper: they call libdynprog_cuda through its C ABI on a B200 and compare with the oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    need = [
        os.path.join(ROOT, "gmap-gsnap_b200", "csrc", "libdynprog_cuda.so"),
        os.path.join(ROOT, "gmap-gsnap_b200", "host", "libdpc_synth.so"),
        os.path.join(ROOT, "oracle", "liboracle_port.so"),
        os.path.join(ROOT, "tests", "emul", "libdpc_emul.so"),
    ]
    if not all(os.path.exists(p) for p in need):
        subprocess.check_call(["make", "-C", ROOT, "all"], stdout=subprocess.DEVNULL)


_ensure_built()

from gmap_gsnap_b200 import api  # noqa: E402
from oracle import checkers  # noqa: E402


def has_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdynprog_ref.so"))


@pytest.fixture(scope="session")
def workload():
    return api.Workload(4_000_000, seed=11, n_frac=0.0005, nchr=4)


@pytest.fixture(scope="session")
def ref(workload):
    if not has_ref():
        pytest.skip("oracle/_ref/libdynprog_ref.so not built (needs /root/reference)")
    r = checkers.RefOracle()
    r.init()
    r.setup(workload.make_setup(splice_prob=r.splice_prob))
    return r


_hook_owner = []


def splice_prob_hook(workload):
    """MaxEnt probabilities for the hook: the compiled reference when present (its Maxent_hr needs the genome
    blocks registered first), else a deterministic stand-in."""
    if has_ref():
        r = checkers.RefOracle()
        r.init()
        r.setup(workload.make_setup())
        workload._keep.append(r)
        _hook_owner.append(r)
        return r.splice_prob
    return api.PROB_FN(lambda which, pos, chroffset, user: ((pos * 2654435761 + which * 97) % 1000003) / 1000003.0)


@pytest.fixture(autouse=True)
def _reclaim_hook_library():
    """The MaxEnt hook handed to the other libraries is a raw function of the compiled-reference library, whose genome
    registration is process-wide: a test that registered another genome there must not leak into the next test."""
    if _hook_owner:
        _hook_owner[-1]._claim()
    yield


@pytest.fixture(scope="session")
def prob_hook(workload):
    return splice_prob_hook(workload)


@pytest.fixture(scope="session")
def port(workload, prob_hook):
    o = checkers.PortOracle()
    o.init()
    o.setup(workload.make_setup(splice_prob=prob_hook))
    return o


@pytest.fixture(scope="session")
def emul(workload, prob_hook):
    e = checkers.EmulLib()
    e.init()
    e.setup(workload.make_setup(splice_prob=prob_hook))
    return e


@pytest.fixture(scope="session")
def cuda(workload, prob_hook):
    lib = api.CudaLib()
    lib.init()
    lib.setup(workload.make_setup(splice_prob=prob_hook))
    lib.open(0)
    yield lib
    lib.close()
