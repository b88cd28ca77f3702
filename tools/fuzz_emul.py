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

# second part: genome gaps under every flavour of splicing IIT (splice-site / intron level x novel splicing or not)
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import util  # noqa: E402
w = api.Workload(4_000_000, seed=11, n_frac=0.0005, nchr=4)
bad_iit, n_iit, t0 = 0, 0, time.time()
for seed in range(7000, 7000 + max(1, nseeds // 6)):
    for intron_level, novel in util.SPLICING_IIT_MODES:
        rng = np.random.default_rng(seed * 4 + intron_level * 2 + novel)
        known, intron = util.splicing_iit_hooks(known_mod=int(rng.choice([2, 3, 5, 17])), intron_mod=int(rng.choice([2, 3, 5])))
        r = checkers.RefOracle(); r.init()
        s = w.make_setup(splice_prob=r.splice_prob, splice_known=known, novelsplicingp=novel, splice_intron=intron, intron_level=intron_level)
        e = checkers.EmulLib(); e.init(); r.setup(s); e.setup(s)
        probs = w.genome_gaps(300, seed=seed, extraband=int(rng.choice([3, 7, 12])), finalp_mode=2, prob_mode_pm=100, long_frac=0.03,
                              long_hi=int(rng.choice([150, 400])))
        r.setup(s); e.setup(s)
        probs = checkers.arm_probability_mode(probs, r)
        bad = api.compare(*r.solve(probs), *e.solve(probs), rtol=1e-6)
        bad_iit += len(bad); n_iit += len(probs)
        if bad:
            print("seed %d intron_level %d novel %d: %s" % (seed, intron_level, novel, bad[:2]), flush=True)
print("FUZZ splicing-IIT modes (emulation) %s: %d problems, %d mismatching fields, %.0f s" % ("OK" if bad_iit == 0 else "FAILED", n_iit, bad_iit, time.time() - t0))
sys.exit(1 if bad_total or bad_iit else 0)
