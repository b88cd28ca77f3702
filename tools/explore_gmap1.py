"""Exploration aid: single-worker-thread comparison (no inter-thread contention)."""
import os, sys, filecmp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmap_gsnap_b200 import gmap_e2e as g
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
case = g.prepare("/tmp/gmap_case1", 20_000_000, 4, n)
for t in (1, 2):
    dt, ref_out, err = g.run_gmap("gmap_ref", case, t)
    print("gmap_ref -t %d: %.2f s %s cpu %s" % (t, dt, err.strip().splitlines()[-1], g.run_gmap.last_cpu), flush=True)
for t, f in ((1, 16), (1, 4), (1, 64), (2, 16)):
    dt, out, err = g.run_gmap("gmap_cuda", case, t, fibers=f, out="/tmp/gmap_case1/c.out")
    st = [l for l in err.splitlines() if "device batches" in l]
    print("gmap_cuda -t %d x %d: %.2f s %s cpu %s identical=%s | %s" % (t, f, dt, err.strip().splitlines()[-1], g.run_gmap.last_cpu, filecmp.cmp(ref_out, out, shallow=False), st[0] if st else ""), flush=True)
