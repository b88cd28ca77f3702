"""BASELINE config 1: the reference's own `make check` case align.test (`gmap -A -g ss.chr17test ss.her2`, golden
tests/align.test.ok in the reference tree), run through two GMAP binaries built by oracle/build_gmap.sh:
gmap_ref (unmodified reference) and gmap_cuda (the five gap-fill solvers replaced by libdynprog_cuda through
gmap-gsnap_b200/host/dynprog_dropin.c).  All 250 matrix fills of this alignment then happen on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
DATA = os.path.join(REFDIR, "align_test")


def run(binary):
    exe = os.path.join(REFDIR, binary)
    if not os.path.exists(exe) or not os.path.exists(os.path.join(DATA, "align.test.ok")):
        pytest.skip("oracle/_ref/%s not built (oracle/build_gmap.sh needs /root/reference)" % binary)
    return subprocess.run([exe, "-A", "-g", "ss.chr17test", "ss.her2"], cwd=DATA, capture_output=True, text=True, timeout=300)


def golden():
    return open(os.path.join(DATA, "align.test.ok")).read()


def test_reference_build_reproduces_align_test():
    r = run("gmap_ref")
    assert r.returncode == 0
    assert r.stdout == golden()


def test_offloaded_build_refuses_without_gpu():
    from gmap_gsnap_b200 import api
    if api.CudaLib().lib.dpc_device_count() > 0:
        pytest.skip("a GPU is present")
    r = run("gmap_cuda")
    assert r.returncode == 9                      # the reference's fatal-error exit code (gmap.c:2287-2308)
    assert "libdynprog_cuda" in r.stderr


@pytest.mark.gpu
def test_offloaded_build_reproduces_align_test():
    r = run("gmap_cuda")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == golden()
