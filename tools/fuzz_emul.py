"""Randomised differential run on the CPU: the lock-step emulation of the device source (tests/emul: the same dpc_core.h /
dpc_rows.h / dpc_pipe.h the kernels are compiled from) vs the compiled reference, over many seeds and parameter corners.
Same generators as tools/fuzz_gpu.py, other seeds; not part of the pytest suite (minutes).  Prints one line per seed."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmap_gsnap_b200 import api
from oracle import checkers

nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
w = api.Workload(8_000_000, seed=23, n_frac=0.001, nchr=4)
ref = checkers.RefOracle(); ref.init(); ref.setup(w.make_setup())
hook = ref.splice_prob
emul = checkers.EmulLib(); emul.init(); emul.setup(w.make_setup(splice_prob=hook))
ref.setup(w.make_setup(splice_prob=hook))
bad_total = 0
t0 = time.time()
for seed in range(5000, 5000 + nseeds):
    rng = np.random.default_rng(seed)
    eb = int(rng.choice([0, 1, 3, 7, 15, 30, 40, 70]))
    sets = [
        w.single_gaps(1200, extraband=eb, seed=seed, len_lo=int(rng.choice([1, 2, 8])), len_hi=int(rng.choice([8, 40, 100, 300])),
                    edge_frac_pm=30, lower_case=1, iupac_pm=10),
        w.end_gaps(1200, extraband=int(rng.choice([0, 3, 10])), seed=seed + 1, len_hi=int(rng.choice([5, 40, 120])), edge_frac_pm=30, lower_case=1, iupac_pm=10),
        w.genome_gaps(600, extraband=int(rng.choice([3, 7, 12, 20])), seed=seed + 2, finalp_mode=2, long_frac=0.05, long_hi=int(rng.choice([200, 611]))),
        w.cdna_gaps(100, seed=seed + 3),
        w.splicejunction_gaps(400, seed=seed + 4, extraband=int(rng.choice([0, 3, 8])), len_hi=int(rng.choice([20, 60, 150]))),
    ]
    emul.setup(w.make_setup(splice_prob=hook)); ref.setup(w.make_setup(splice_prob=hook))
    probs = np.concatenate(sets)
    probs = probs[rng.permutation(len(probs))]
    probs = checkers.arm_probability_mode(probs, ref)
    want = ref.solve(probs)
    bad = api.compare(*want, *emul.solve(probs), rtol=1e-6)
    bad_total += len(bad)
    print("seed %d: %d problems, extraband %d, mismatches %d %s" % (seed, len(probs), eb, len(bad), bad[:2]), flush=True)
print("FUZZ (emulation) %s: %d seeds, %d mismatching fields, %.0f s" % ("OK" if bad_total == 0 else "FAILED", nseeds, bad_total, time.time() - t0))
sys.exit(1 if bad_total else 0)
