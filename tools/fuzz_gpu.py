"""Randomised differential run on the GPU box: libdynprog_cuda vs the compiled reference over many seeds and
parameter corners (bands 0-40, tiny and long sequences, lower case / IUPAC / N, segment edges, all kinds incl. the
splice-junction solvers).  Not part of the pytest suite (minutes of reference CPU time); prints one line per seed."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmap_gsnap_b200 import api
from oracle import checkers

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
w = api.Workload(8_000_000, seed=17, n_frac=0.001, nchr=4)
ref = checkers.RefOracle(); ref.init(); ref.setup(w.make_setup())
hook = ref.splice_prob
lib = api.CudaLib(); lib.init(); lib.setup(w.make_setup(splice_prob=hook)); lib.open(0)
ref.setup(w.make_setup(splice_prob=hook))
bad_total = 0
t0 = time.time()
for seed in range(1000, 1000 + nseeds):
    rng = np.random.default_rng(seed)
    eb = int(rng.choice([0, 1, 3, 7, 15, 30, 40]))
    sets = [
        w.single_gaps(4000, extraband=eb, seed=seed, len_lo=int(rng.choice([1, 2, 8])), len_hi=int(rng.choice([8, 40, 100, 300])),
                    edge_frac_pm=30, lower_case=1, iupac_pm=10),
        w.end_gaps(4000, extraband=int(rng.choice([0, 3, 10])), seed=seed + 1, len_hi=int(rng.choice([5, 40, 120])), edge_frac_pm=30, lower_case=1, iupac_pm=10),
        w.genome_gaps(2000, extraband=int(rng.choice([3, 7, 12])), seed=seed + 2, finalp_mode=2, long_frac=0.05, long_hi=int(rng.choice([200, 611]))),
        w.cdna_gaps(300, seed=seed + 3),
        w.splicejunction_gaps(1500, seed=seed + 4, extraband=int(rng.choice([0, 3, 8])), len_hi=int(rng.choice([20, 60, 150]))),
    ]
    # generators that plant splice sites change the genome: register it again on both sides
    lib.setup(w.make_setup(splice_prob=hook)); ref.setup(w.make_setup(splice_prob=hook))
    probs = np.concatenate(sets)
    probs = probs[rng.permutation(len(probs))]
    probs = checkers.arm_probability_mode(probs, ref)
    want = ref.solve(probs)
    bad = api.compare(*want, *lib.solve(probs), rtol=1e-6)
    # the device pipeline (both Pair-expansion routes) on the hook-free part of the batch
    plain = probs[(probs["use_probabilities_p"] == 0) & (probs["kind"] <= api.END3_GAP)]
    want_plain = ref.solve(plain)
    for path in (3, 4):
        lib.lib.dpc_set_path(path)
        bad += api.compare(*want_plain, *lib.solve(plain), rtol=1e-6)
    lib.lib.dpc_set_path(0)
    bad_total += len(bad)
    print("seed %d: %d problems, extraband %d, mismatches %d %s" % (seed, len(probs), eb, len(bad), bad[:2]), flush=True)
print("FUZZ %s: %d seeds, %d mismatching fields, %.0f s" % ("OK" if bad_total == 0 else "FAILED", nseeds, bad_total, time.time() - t0))
lib.close()
sys.exit(1 if bad_total else 0)
