"""Exploration aid: does the output thread (format of the printed alignments) bound either binary?"""
import os, sys, filecmp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmap_gsnap_b200 import gmap_e2e as g
case = g.prepare("/tmp/gmap_case2", 100_000_000, 4, 16000)
for fmt in ("-A", "-S", "-f samse"):
    os.environ["GMAP_OUTFMT"] = fmt
    dt, ref_out, err = g.run_gmap("gmap_ref", case, 16)
    print("fmt %s gmap_ref: %.2f s %s cpu %s" % (fmt, dt, err.strip().splitlines()[-1], g.run_gmap.last_cpu), flush=True)
    dt, out, err = g.run_gmap("gmap_cuda", case, 16, fibers=16, out="/tmp/gmap_case2/c.out")
    st = [l for l in err.splitlines() if "device batches" in l]
    print("fmt %s gmap_cuda: %.2f s %s cpu %s identical=%s" % (fmt, dt, [l for l in err.strip().splitlines() if l.startswith("Processed")][-1], g.run_gmap.last_cpu, filecmp.cmp(ref_out, out, shallow=False)), flush=True)
