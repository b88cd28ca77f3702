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


def splicing_iit_hooks(known_mod=3, intron_mod=2):
    """Deterministic stand-ins for the splicing IIT lookups behind dpc_setup_t.splice_known / splice_intron."""
    known = api.KNOWN_FN(lambda which, chrnum, pos, sign, user: int((pos * 7 + which + chrnum) % known_mod == 0))
    intron = api.INTRON_FN(lambda chrnum, pos1, pos2, sign, user: int((pos1 + 3 * pos2 + sign + chrnum) % intron_mod == 0))
    return known, intron


SPLICING_IIT_MODES = [(0, 1), (0, 0), (1, 1), (1, 0)]        # (intron_level, novelsplicingp)


def long_nogaps_ends(w, lengths=(16383, 16384, 20000, 40000), seed=5):
    """QUERYEND_NOGAPS end gaps longer than one traceback op can carry (the reference does not clip these ends,
    dynprog.c:5179-5186): both ends, both strands, random query (a quarter of the columns match)."""
    rng = np.random.default_rng(seed)
    base = w.end_gaps(64, seed=seed)
    out = []
    for n in lengths:
        for kind in (api.END3_GAP, api.END5_GAP):
            for watson in (0, 1):
                p = base[(base["kind"] == kind)][:1].copy()
                p["watsonp"] = watson
                p["chrpos"], p["genomiclength"] = 1000, 90000
                off = 30000 if kind == api.END3_GAP else 30000 + n
                q = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
                w._keep.append(q)
                p["endalign"] = api.QUERYEND_NOGAPS
                p["length1"], p["length2"] = n, n + 10
                p["offset2"] = off
                p["seq1"] = q.ctypes.data + (n - 1 if kind == api.END5_GAP else 0)
                out.append(p)
    return np.concatenate(out)
