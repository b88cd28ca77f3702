"""Generates tests/golden/golden_v1.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libdynprog_ref.so,
built by oracle/Makefile from /root/reference/src).  Run in the build container only:

    python tests/golden/make_golden.py

Each problem set is stored with its inputs (genome blocks, problems, query bytes), the reference's outputs
(results, Pair records), and every MaxEnt probability the reference's hook returned while the restatement
solved the set, so that the golden tests do not need the reference at run time."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gmap_gsnap_b200 import api  # noqa: E402
from oracle import checkers


def edge_cases(w):
    """Early returns and odd inputs of the five entry points (SURVEY.md 8a "quirks")."""
    import ctypes as C
    rng = np.random.default_rng(5)
    sg = w.single_gaps(60, extraband=3, seed=77)
    gg = w.genome_gaps(40, long_frac=0.0, finalp_mode=2, seed=78)
    eg = w.end_gaps(120, seed=79)
    cg = w.cdna_gaps(12, seed=80)
    sg["length1"][0] = 612          # > maxlength1: finalscore -10000, index bumped, NULL (dynprog.c:4509)
    sg["length2"][1] = 2001         # > maxlength2
    sg["widebandp"][2:12] = 0       # unwidened band
    sg["length2"][2:12] = np.clip(sg["length2"][2:12], sg["length1"][2:12] - 3, sg["length1"][2:12] + 3)
    sg["chrhigh"][12:16] = sg["chroffset"][12:16]     # outside the chromosome: every genomic nt is '*'
    sg["defect_rate"][16:20] = [0.0029, 0.003, 0.0139, 0.014]
    gg["length1"][0] = 1            # L1 <= 1: NEG, index NOT bumped (4855)
    gg["length1"][1] = 0
    gg["length1"][2] = 612; gg["length2"][2] = 620; gg["length2R"][2] = 620    # too long (4922)
    gg["splicingp"][3:8] = 0
    gg["maxpeelback"][8:12] = 5     # L1 > 4*maxpeelback: SINGLE penalties (4862)
    gg["chrhigh"][12:14] = gg["chroffset"][12:14]
    gg["finalp"][12:14] = 0
    eg["length1"][0] = 0            # (5140)
    eg["length2"][1] = 0
    eg["length1"][2] = -3
    for i in range(3, 40):          # unrelated query: best end point is (0,0) or the list gets rejected (5259)
        n = int(eg["length1"][i])
        addr = int(eg["seq1"][i]) - (n - 1 if eg["kind"][i] == api.END5_GAP else 0)
        C.memmove(addr, bytes(rng.choice(list(b"ACGT"), size=n).tolist()), n)
    eg["endalign"][3:15] = api.QUERYEND_GAP
    eg["endalign"][15:27] = api.QUERYEND_NOGAPS
    eg["endalign"][27:33] = api.BEST_LOCAL
    eg["endalign"][33:40] = api.QUERYEND_INDELS
    eg["endalign"][40:60] = api.QUERYEND_INDELS
    eg["chrhigh"][60:64] = eg["chroffset"][60:64]
    cg["length2"][0] = 1            # (4605)
    cg["length2"][1] = 0
    return np.concatenate([sg, gg, eg, cg])


def main():
    w = api.Workload(300_000, seed=20121, n_frac=0.001, nchr=3)
    ref = checkers.RefOracle()
    ref.init()
    ref.setup(w.make_setup(splice_prob=ref.splice_prob))
    calls = {}

    def rec(which, pos, chroffset, user):
        v = ref.lib.ref_splice_prob(which, pos, chroffset, None)
        calls[(which, pos, chroffset)] = v
        return v

    hook = api.PROB_FN(rec)
    port = checkers.PortOracle()
    port.init()
    port.setup(w.make_setup(splice_prob=hook))

    sets = {
        "single30": w.single_gaps(400, extraband=30, edge_frac_pm=30, lower_case=1, iupac_pm=15),
        "single3": w.single_gaps(400, extraband=3, edge_frac_pm=30, lower_case=1),
        "end": w.end_gaps(600, edge_frac_pm=40, lower_case=1, iupac_pm=15),
        "genome": w.genome_gaps(400, finalp_mode=2, prob_mode_pm=150, long_frac=0.05, long_hi=300, iupac_pm=5),
        "cdna": w.cdna_gaps(200),
    }
    sets["edge"] = edge_cases(w)
    out = {"blocks": w.blocks, "nbases": np.int64(w.nbases)}
    for name, probs in sets.items():
        probs = checkers.arm_probability_mode(probs, ref)
        res, pairs, off = ref.solve(probs)
        bad = api.compare(res, pairs, off, *port.solve(probs))   # also records the hook calls
        assert not bad, (name, bad)
        det, qbuf, offs = api.detach(probs)
        out[name + "_problems"] = det
        out[name + "_qbuf"] = qbuf
        out[name + "_qoff"] = offs
        out[name + "_results"] = res
        out[name + "_pairs"] = pairs
        out[name + "_pairoff"] = off
        print(name, len(probs), "problems", len(pairs), "pairs; null lists:", int(res["null_list"].sum()))
    keys = np.array(sorted(calls.keys()), dtype=np.int64).reshape(-1, 3)
    out["prob_keys"] = keys
    out["prob_vals"] = np.array([calls[tuple(k)] for k in keys.tolist()], dtype=np.float64)
    path = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(keys), "probabilities")


if __name__ == "__main__":
    main()
