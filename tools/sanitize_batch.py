"""A small mixed batch of every solver through every path of dpc_solve (host half, device pipeline with device / host
Pair expansion) and both fill routines, compared with the oracle.  Meant to be run under compute-sanitizer
(tools/sanitize_gpu.sh): small enough for memcheck / racecheck, yet it reaches every kernel of the library."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmap_gsnap_b200 import api  # noqa: E402
from oracle import checkers  # noqa: E402

w = api.Workload(2_000_000, seed=3, nchr=2)
hook = api.PROB_FN(lambda which, pos, chroffset, user: ((pos * 2654435761 + which * 97) % 1000003) / 1000003.0)
setup = w.make_setup(splice_prob=hook)
probs = np.concatenate([
    w.single_gaps(300, extraband=30, edge_frac_pm=50, lower_case=1, iupac_pm=10), w.single_gaps(100, extraband=3),
    w.single_gaps(40, extraband=64, len_lo=60, len_hi=100), w.end_gaps(400, edge_frac_pm=100),
    w.genome_gaps(200, finalp_mode=2, long_frac=0.1, long_hi=611), w.cdna_gaps(40), w.splicejunction_gaps(60)])
plain = probs[(probs["kind"] <= api.END3_GAP)]
port = checkers.PortOracle()
port.init()
port.setup(setup)
want_all, want_plain = port.solve(probs), port.solve(plain)
lib = api.CudaLib()
lib.init()
lib.setup(setup)
lib.open(0)
bad = []
for fill in (0, 1):
    lib.lib.dpc_set_fill(fill)
    for path, batch, want in ((0, probs, want_all), (1, probs, want_all), (3, plain, want_plain), (4, plain, want_plain)):
        lib.lib.dpc_set_path(path)
        bad += api.compare(*want, *lib.solve(batch), rtol=1e-6)
lib.lib.dpc_set_path(0)
lib.lib.dpc_set_fill(0)
lib.close()
print("sanitize batch: %d problems x 2 fills x 4 paths, mismatches: %d %s" % (len(probs), len(bad), bad[:3]))
sys.exit(1 if bad else 0)
