"""Helpers shared by the CPU and GPU tests."""
import os

import numpy as np

from gmap_gsnap_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden_v1.npz")
GOLDEN_SETS = ["single30", "single3", "end", "genome", "cdna", "edge"]


class Golden:
    """tests/golden/golden_v1.npz: inputs and outputs of the compiled reference (see make_golden.py)."""

    def __init__(self):
        z = np.load(GOLDEN)
        self.z = z
        self.workload = api.Workload.__new__(api.Workload)
        self.workload.nbases = int(z["nbases"])
        self.workload.nchr = 3
        self.workload.seed = 0
        self.workload.blocks = z["blocks"].copy()
        self.workload._keep = []
        probs = dict(zip(map(tuple, z["prob_keys"].tolist()), z["prob_vals"].tolist()))
        self.missing = []

        def hook(which, pos, chroffset, user):
            v = probs.get((which, pos, chroffset))
            if v is None:
                self.missing.append((which, pos, chroffset))
                return 0.0
            return v

        self.hook = api.PROB_FN(hook)
        self._keep = []

    def setup(self):
        return self.workload.make_setup(splice_prob=self.hook)

    def problems(self, name):
        qbuf = self.z[name + "_qbuf"].copy()
        self._keep.append(qbuf)
        return api.attach(self.z[name + "_problems"], qbuf, self.z[name + "_qoff"])

    def expected(self, name):
        return self.z[name + "_results"], self.z[name + "_pairs"], self.z[name + "_pairoff"]


def mixed_problems(w, n_each, seed, **kw):
    """One array with every kind, shuffled, as a batch from stage3 would look."""
    sets = [
        w.single_gaps(n_each, extraband=30, seed=seed, edge_frac_pm=20, lower_case=1, iupac_pm=10),
        w.single_gaps(n_each, extraband=3, seed=seed + 1, edge_frac_pm=20),
        w.end_gaps(n_each, seed=seed + 2, edge_frac_pm=30, lower_case=1, iupac_pm=10),
        w.genome_gaps(n_each, seed=seed + 3, finalp_mode=2, long_frac=kw.get("long_frac", 0.05), long_hi=kw.get("long_hi", 300)),
        w.cdna_gaps(max(n_each // 4, 1), seed=seed + 4),
    ]
    allp = np.concatenate(sets)
    rng = np.random.default_rng(seed)
    return allp[rng.permutation(len(allp))]
