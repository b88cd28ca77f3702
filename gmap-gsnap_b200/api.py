"""ctypes mirror of include/dynprog_cuda.h plus the synthetic workload generator.

Host-side plumbing only: loads libdynprog_cuda.so (the product), and builds
problem arrays as numpy structured arrays laid out exactly like dpc_problem_t.
The reference interface this mirrors is src/dynprog.h (Dynprog_single_gap 71-82,
Dynprog_cdna_gap 84-97, Dynprog_genome_gap 99-117, Dynprog_end5_gap 119-132,
Dynprog_end3_gap 148-161).  Nothing here falls back to a CPU solver: if the
CUDA library is missing, loading raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

DPC_UNSET = -2000000001
SINGLE_GAP, GENOME_GAP, CDNA_GAP, END5_GAP, END3_GAP, END5_SPLICEJUNCTION, END3_SPLICEJUNCTION, MICROEXON_INT = range(8)
QUERYEND_GAP, QUERYEND_INDELS, QUERYEND_NOGAPS, BEST_LOCAL = range(4)
KIND_NAMES = ["single_gap", "genome_gap", "cdna_gap", "end5_gap", "end3_gap"]


class Problem(C.Structure):
    _fields_ = [
        ("seq1", C.c_void_p), ("seq1R", C.c_void_p),
        ("kind", C.c_int32), ("endalign", C.c_int32),
        ("length1", C.c_int32), ("length1R", C.c_int32), ("length2", C.c_int32), ("length2R", C.c_int32),
        ("offset1", C.c_int32), ("offset1R", C.c_int32), ("offset2", C.c_int32), ("offset2R", C.c_int32),
        ("chroffset", C.c_uint32), ("chrhigh", C.c_uint32), ("chrpos", C.c_uint32), ("genomiclength", C.c_uint32),
        ("chrnum", C.c_int32), ("cdna_direction", C.c_int32), ("extraband", C.c_int32),
        ("maxpeelback", C.c_int32), ("score_threshold", C.c_int32), ("dynprogindex", C.c_int32),
        ("watsonp", C.c_uint8), ("jump_late_p", C.c_uint8), ("widebandp", C.c_uint8), ("halfp", C.c_uint8),
        ("finalp", C.c_uint8), ("use_probabilities_p", C.c_uint8), ("splicingp", C.c_uint8), ("reserved", C.c_uint8),
        ("defect_rate", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("null_list", C.c_int32), ("dynprogindex_out", C.c_int32), ("finalscore", C.c_int32),
        ("nmatches", C.c_int32), ("nmismatches", C.c_int32), ("nopens", C.c_int32), ("nindels", C.c_int32),
        ("new_leftgenomepos", C.c_int32), ("new_rightgenomepos", C.c_int32),
        ("exonhead", C.c_int32), ("introntype", C.c_int32), ("incompletep", C.c_int32),
        ("npairs", C.c_int32), ("reserved", C.c_int32),
        ("left_prob", C.c_double), ("right_prob", C.c_double),
    ]


class Pair(C.Structure):
    _fields_ = [
        ("querypos", C.c_int32), ("genomepos", C.c_int32), ("dynprogindex", C.c_int32),
        ("cdna", C.c_char), ("comp", C.c_char), ("genome", C.c_char), ("gapp", C.c_uint8),
    ]


PROB_FN = C.CFUNCTYPE(C.c_double, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p)
KNOWN_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p)
INTRON_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p)


class Setup(C.Structure):
    _fields_ = [
        ("genome_blocks", C.c_void_p), ("genome_nwords", C.c_uint64),
        ("novelsplicingp", C.c_int32), ("intron_level", C.c_int32),
        ("splice_prob", C.c_void_p), ("splice_known", C.c_void_p), ("user", C.c_void_p),
        ("splice_intron", C.c_void_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("nproblems", C.c_int64), ("nmatrices", C.c_int64), ("cells", C.c_int64),
        ("fill_bytes", C.c_int64), ("pipeline_chunks", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("launches", C.c_int32), ("host_chunks", C.c_int32),
    ]


class SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("genome_nbases", C.c_uint64),
        ("nchr", C.c_int32), ("extraband", C.c_int32), ("len_lo", C.c_int32), ("len_hi", C.c_int32),
        ("intron_lo", C.c_int32), ("intron_hi", C.c_int32),
        ("p_sub", C.c_double), ("p_del", C.c_double), ("p_ins", C.c_double),
        ("long_frac", C.c_double), ("long_hi", C.c_int32), ("edge_frac_pm", C.c_int32),
        ("finalp_mode", C.c_int32), ("prob_mode_pm", C.c_int32), ("lower_case", C.c_int32),
        ("iupac_pm", C.c_int32), ("reserved", C.c_int32),
    ]


PROBLEM_DT = np.dtype(Problem)
RESULT_DT = np.dtype(Result)
PAIR_DT = np.dtype(Pair)
assert PROBLEM_DT.itemsize == 112 and RESULT_DT.itemsize == 72 and PAIR_DT.itemsize == 16


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- synth
_synth = None


def synth_lib():
    global _synth
    if _synth is None:
        _synth = C.CDLL(os.path.join(HERE, "host", "libdpc_synth.so"))
        _synth.synth_genome_nwords.restype = C.c_uint64
        _synth.synth_genome_nwords.argtypes = [C.c_uint64]
        _synth.synth_genome.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_double]
        _synth.synth_genome_from_ascii.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        _synth.synth_genome_char.restype = C.c_char
        _synth.synth_genome_char.argtypes = [C.c_void_p, C.c_uint32]
        for name in ("synth_single_gaps", "synth_end_gaps", "synth_genome_gaps", "synth_cdna_gaps"):
            f = getattr(_synth, name)
            f.restype = C.c_int64
            f.argtypes = [C.POINTER(SynthParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    return _synth


class Workload:
    """A genome plus problem arrays.  Keeps the query buffers alive (the problems point into them)."""

    def __init__(self, nbases, seed=1, n_frac=0.0, nchr=4):
        lib = synth_lib()
        self.nbases = int(nbases)
        self.nchr = nchr
        self.seed = seed
        self.blocks = np.zeros(int(lib.synth_genome_nwords(self.nbases)), dtype=np.uint32)
        lib.synth_genome(_ptr(self.blocks), self.nbases, seed, n_frac)
        self._keep = []
        self.version = 0        # bumped whenever a generator plants splice sites into the genome

    @classmethod
    def from_ascii(cls, seq):
        lib = synth_lib()
        self = cls.__new__(cls)
        self.nbases = len(seq)
        self.nchr = 1
        self.seed = 0
        self.blocks = np.zeros(int(lib.synth_genome_nwords(self.nbases)), dtype=np.uint32)
        lib.synth_genome_from_ascii(_ptr(self.blocks), seq.encode(), self.nbases)
        self._keep = []
        self.version = 0
        return self

    def params(self, **kw):
        sp = SynthParams()
        sp.seed = self.seed
        sp.genome_nbases = self.nbases
        sp.nchr = self.nchr
        sp.extraband = 3
        sp.len_lo, sp.len_hi = 10, 100
        sp.intron_lo, sp.intron_hi = 50, 20000
        sp.p_sub, sp.p_del, sp.p_ins = 0.05, 0.03, 0.03
        sp.long_frac, sp.long_hi = 0.0, 600
        for k, v in kw.items():
            setattr(sp, k, v)      # includes seed=... for an independent stream
        return sp

    def _gen(self, fname, n, qbytes_per, sp):
        lib = synth_lib()
        probs = np.zeros(n, dtype=PROBLEM_DT)
        qbuf = np.zeros(int(n) * int(qbytes_per) + 4096, dtype=np.uint8)
        used = getattr(lib, fname)(C.byref(sp), _ptr(self.blocks), n, _ptr(probs), _ptr(qbuf), qbuf.size)
        if used < 0:
            raise RuntimeError("%s: query buffer too small" % fname)
        self._keep.append(qbuf)
        self.last_qbuf = qbuf          # the contiguous query buffer the problems just generated point into
        return probs

    def single_gaps(self, n, extraband=30, **kw):
        d = dict(extraband=extraband, len_lo=10, len_hi=100, p_sub=0.05, p_del=0.03, p_ins=0.03)
        d.update(kw)
        sp = self.params(**d)
        return self._gen("synth_single_gaps", n, 2 * sp.len_hi + 16, sp)

    def end_gaps(self, n, extraband=3, **kw):
        d = dict(extraband=extraband, len_lo=1, len_hi=40, p_sub=0.03, p_del=0.0025, p_ins=0.0025)
        d.update(kw)
        sp = self.params(**d)
        return self._gen("synth_end_gaps", n, 2 * (sp.len_hi + 11) + 16, sp)

    def genome_gaps(self, n, extraband=7, long_frac=0.1, long_hi=600, **kw):
        d = dict(extraband=extraband, len_lo=22, len_hi=60, p_sub=0.02, p_del=0.005, p_ins=0.005,
                 long_frac=long_frac, long_hi=long_hi)
        d.update(kw)
        sp = self.params(**d)
        per = 2 * (sp.long_hi if sp.long_frac > 0 else sp.len_hi) + 32
        self.version = getattr(self, "version", 0) + 1     # plants GT..AG etc. into the genome
        return self._gen("synth_genome_gaps", n, per, sp)

    def cdna_gaps(self, n, extraband=7, **kw):
        d = dict(extraband=extraband, len_lo=4, len_hi=40, intron_lo=10, intron_hi=60, p_sub=0.02, p_del=0.005, p_ins=0.005)
        d.update(kw)
        sp = self.params(**d)
        return self._gen("synth_cdna_gaps", n, 2 * sp.len_hi + sp.intron_hi + 64, sp)

    def splicejunction_gaps(self, n, extraband=3, seed=77, len_lo=8, len_hi=60, p_sub=0.04, p_indel=0.01, long_run_pm=30):
        """Problems for Dynprog_end5/3_splicejunction (dynprog.c:5411-5552, 5869-6012): the genomic side is a
        splice-junction STRING (anchor exon piece of `contlength` bases + the far exon piece), not a genome window;
        the query is a mutated copy of it.  A few problems get a genome-only run of 9+ flanked by GT..AG / CT..AC
        so that the gapholder path of add_genomeskip (2448-2451) is exercised.  Pure numpy; test workloads only."""
        rng = np.random.default_rng(seed)
        probs = np.zeros(n, dtype=PROBLEM_DT)
        qbuf = np.zeros(n * (2 * len_hi + 64) * 2 + 4096, dtype=np.uint8)
        base = qbuf.ctypes.data
        at = 0
        acgt = np.frombuffer(b"ACGT", np.uint8)
        for i in range(n):
            five = bool(rng.integers(0, 2))
            L2 = int(rng.integers(len_lo, len_hi + 1)) + 10
            cont = int(rng.integers(0, L2 + 1)) if rng.random() < 0.1 else int(rng.integers(1, max(2, L2 // 2)))
            g = acgt[rng.integers(0, 4, L2)]
            if rng.random() < 0.02:
                g[rng.integers(0, L2)] = ord("N")
            q = []
            j = 0
            skip_at = int(rng.integers(2, max(3, L2 - 14))) if rng.integers(0, 1000) < long_run_pm and L2 > 30 else -1
            while j < L2 - 10:
                if j == skip_at:
                    run = int(rng.integers(9, 13))
                    if rng.random() < 0.7:
                        g[j:j + 2] = (71, 84) if rng.random() < 0.5 else (67, 84)
                        g[j + run - 2:j + run] = (65, 71) if g[j] == 71 else (65, 67)
                    j += run
                    continue
                u = rng.random()
                if u < p_indel:
                    j += 1
                    continue
                if u < 2 * p_indel:
                    q.append(int(acgt[rng.integers(0, 4)]))
                ch = int(g[j])
                if u > 1 - p_sub:
                    ch = int(acgt[rng.integers(0, 4)])
                if rng.random() < 0.01:
                    ch |= 0x20
                q.append(ch)
                j += 1
            if not q:
                q = [65]
            q = np.array(q, np.uint8)
            L1 = len(q)
            # both strings are laid out in reading order; the END5 variant hands over pointers to their LAST characters
            qs, gs = (q[::-1], g[::-1]) if five else (q, g)
            qbuf[at:at + L1] = qs
            qa = at; at += L1 + 1
            qbuf[at:at + L2] = gs
            ga = at; at += L2 + 1
            pr = probs[i]
            pr["kind"] = END5_SPLICEJUNCTION if five else END3_SPLICEJUNCTION
            pr["seq1"] = base + qa + (L1 - 1 if five else 0)
            pr["seq1R"] = base + ga + (L2 - 1 if five else 0)
            pr["length1"], pr["length2"], pr["length2R"] = L1, L2, cont
            o1, oa, far = int(rng.integers(20, 2000)), int(rng.integers(1000, 100000)), int(rng.integers(200, 50000))
            pr["offset1"] = o1
            pr["offset2"] = oa
            pr["offset2R"] = oa - far if five else oa + far
            pr["chroffset"], pr["chrhigh"], pr["chrpos"], pr["genomiclength"] = 0, self.nbases, 1000, 200000
            pr["cdna_direction"] = int(rng.integers(-1, 2))
            pr["extraband"] = extraband
            pr["maxpeelback"] = 11
            pr["dynprogindex"] = int(rng.integers(1, 50)) * (1 if rng.random() < 0.5 else -1)
            pr["watsonp"], pr["jump_late_p"], pr["widebandp"], pr["splicingp"] = int(rng.integers(0, 2)), int(rng.integers(0, 2)), 1, 1
            pr["defect_rate"] = float(rng.choice([0.0, 0.01, 0.05]))
        self._keep.append(qbuf)
        return probs

    def _set_base(self, pos, ch):
        """Writes one base (A C G T) into the 3-word genome blocks (high, low, flags per 32 nt, genome.c:9325)."""
        code = "ACGT".index(ch)
        w, bit = 3 * (pos >> 5), pos & 31
        if bit < 16:
            self.blocks[w + 1] = (int(self.blocks[w + 1]) & ~(3 << (2 * bit))) | (code << (2 * bit))
        else:
            self.blocks[w] = (int(self.blocks[w]) & ~(3 << (2 * bit - 32))) | (code << (2 * bit - 32))
        self.blocks[w + 2] = int(self.blocks[w + 2]) & ~(1 << bit)

    def microexon_problems(self, n, seed=91, span_hi=4000):
        """Problems for Dynprog_microexon_int (dynprog.c:7127-7429): a query piece = left exon end + microexon + right
        exon start, against a genomic segment where the three parts are separated by two introns (GT..AG, or CT..AC
        for the antisense direction).  Most problems get the microexon planted (some with a mismatch in a flank, some
        with a decoy copy of the middle piece without splice sites); the rest have none.  Plants into the genome:
        register it again afterwards (the `version` counter tells CudaLib.refresh_genome).  Pure numpy; test workloads."""
        rng = np.random.default_rng(seed)
        probs = np.zeros(n, dtype=PROBLEM_DT)
        qbuf = np.zeros(n * 64 + 4096, dtype=np.uint8)
        base, at = qbuf.ctypes.data, 0
        chrlen = self.nbases // self.nchr
        comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
        for i in range(n):
            chrn = int(rng.integers(0, self.nchr))
            glen = int(rng.integers(30000, 60000))
            chrpos = int(rng.integers(0, chrlen - glen))
            watson = bool(rng.integers(0, 2))
            cdna = 1 if rng.random() < 0.5 else -1
            chroffset = chrn * chrlen
            a, mid, b = int(rng.integers(4, 16)), int(rng.integers(3, 13)), int(rng.integers(4, 16))
            span = int(rng.integers(120, span_hi))
            lo = int(rng.integers(200, glen - span - 200))
            ro = lo + span
            i1len = int(rng.integers(20, span - a - b - mid - 30))
            m0 = lo + a + i1len
            r0 = ro - b + 1

            def put(s, text):
                for k, ch in enumerate(text):
                    gpos = chroffset + chrpos + (s + k if watson else glen - 1 - (s + k))
                    self._set_base(gpos, ch if watson else comp[ch])

            q = "".join("ACGT"[x] for x in rng.integers(0, 4, a + mid + b))
            don, acc = ("GT", "AG") if cdna > 0 else ("CT", "AC")
            planted = rng.random() < 0.8
            put(lo, q[:a])
            put(r0, q[a + mid:])
            if planted:
                put(lo + a, don); put(m0 - 2, acc); put(m0, q[a:a + mid]); put(m0 + mid, don); put(r0 - 2, acc)
                if rng.random() < 0.3 and m0 + mid + 40 < r0 - 20:      # a decoy copy of the middle piece further right
                    put(m0 + mid + 20, q[a:a + mid])
            ql = list(q)
            if rng.random() < 0.15:
                k = int(rng.integers(0, a)); ql[k] = "ACGT"[("ACGT".index(ql[k]) + 1) % 4]
            if rng.random() < 0.15:
                k = a + mid + int(rng.integers(0, b)); ql[k] = "ACGT"[("ACGT".index(ql[k]) + 2) % 4]
            if rng.random() < 0.1:
                ql[int(rng.integers(0, len(ql)))] = "N"
            if rng.random() < 0.1:
                k = int(rng.integers(0, len(ql))); ql[k] = ql[k].lower()
            L1 = len(ql)
            qbuf[at:at + L1] = np.frombuffer("".join(ql).encode(), np.uint8)
            pr = probs[i]
            pr["kind"] = MICROEXON_INT
            pr["seq1"] = base + at
            at += L1 + 1
            pr["length1"], pr["length2"], pr["length2R"] = L1, L1 + 8, L1 + 8
            pr["offset1"], pr["offset2"], pr["offset2R"] = int(rng.integers(20, 3000)), lo, ro
            pr["chroffset"], pr["chrhigh"], pr["chrpos"], pr["genomiclength"] = chroffset, chroffset + chrlen, chrpos, glen
            pr["chrnum"] = chrn + 1
            pr["cdna_direction"] = cdna
            pr["maxpeelback"] = 11
            pr["dynprogindex"] = int(rng.integers(1, 50)) * (1 if rng.random() < 0.5 else -1)
            pr["watsonp"], pr["jump_late_p"], pr["widebandp"], pr["splicingp"] = int(watson), int(not watson), 1, 1
            pr["defect_rate"] = float(rng.choice([0.0, 0.01, 0.05]))
        self._keep.append(qbuf)
        self.version = getattr(self, "version", 0) + 1
        return probs

    def make_setup(self, splice_prob=None, splice_known=None, novelsplicingp=1, splice_intron=None, intron_level=0):
        s = Setup()
        s.genome_blocks = self.blocks.ctypes.data
        s.genome_nwords = self.blocks.size
        s.novelsplicingp = novelsplicingp
        s.splice_prob = C.cast(splice_prob, C.c_void_p).value if splice_prob is not None else None
        s.splice_known = C.cast(splice_known, C.c_void_p).value if splice_known is not None else None
        s.user = None
        s.intron_level = intron_level
        s.splice_intron = C.cast(splice_intron, C.c_void_p).value if splice_intron is not None else None
        self._keep.append((splice_prob, splice_known, splice_intron))
        s._workload = self
        return s


# --------------------------------------------------------------------------- solvers
class _SolverLib:
    """Common driver for libraries that speak dpc_problem_t / dpc_result_t (the product here; the CPU checkers in
    oracle/checkers.py)."""

    prefix = None

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.path = path

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)


def _solve_common(fn, problems, want_pairs=True, pair_cap=None):
    n = len(problems)
    results = np.zeros(n, dtype=RESULT_DT)
    pair_off = np.zeros(n + 1, dtype=np.int64)
    if want_pairs:
        if pair_cap is None:
            pair_cap = int((problems["length1"].astype(np.int64) + problems["length1R"] + problems["length2"]
                            + problems["length2R"] + 8).clip(min=8).sum()) + 64
        pairs = np.zeros(pair_cap, dtype=PAIR_DT)
        rc = fn(_ptr(problems), n, _ptr(results), _ptr(pairs), pair_cap, _ptr(pair_off))
    else:
        pairs = np.zeros(0, dtype=PAIR_DT)
        rc = fn(_ptr(problems), n, _ptr(results), None, 0, _ptr(pair_off))
    if rc != 0:
        raise RuntimeError("solve failed with code %d" % rc)
    return results, pairs[: pair_off[n]] if want_pairs else pairs, pair_off


class CudaLib(_SolverLib):
    """gmap-gsnap_b200/csrc/libdynprog_cuda.so -- THE PRODUCT."""

    def __init__(self, path=None):
        path = path or os.environ.get("DPC_LIB") or os.path.join(HERE, "csrc", "libdynprog_cuda.so")
        if not os.path.exists(path):
            raise RuntimeError("libdynprog_cuda.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        super().__init__(path)
        L = self.lib
        L.dpc_init.argtypes = [C.c_int] * 6
        L.dpc_setup.argtypes = [C.POINTER(Setup)]
        L.dpc_ctx_new.restype = C.c_void_p
        L.dpc_ctx_new.argtypes = [C.c_int]
        L.dpc_ctx_free.argtypes = [C.c_void_p]
        L.dpc_add_bulk.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dpc_add.argtypes = [C.c_void_p, C.c_void_p]
        for name in ("dpc_flush", "dpc_wait", "dpc_reset", "dpc_relaunch"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.dpc_result.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.dpc_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.dpc_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.dpc_sync.argtypes = [C.c_void_p]
        L.dpc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.dpc_set_fill.argtypes = [C.c_int]
        L.dpc_stream.restype = C.c_void_p
        L.dpc_stream.argtypes = [C.c_void_p]
        L.dpc_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.dpc_ctx_new_multi.restype = C.c_void_p
        L.dpc_ctx_new_multi.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.dpc_host_register.argtypes = [C.c_void_p, C.c_uint64]
        L.dpc_host_unregister.argtypes = [C.c_void_p]
        L.dpc_set_path.argtypes = [C.c_int]
        L.dpc_device_count.restype = C.c_int
        L.dpc_measure_int_peak.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.dpc_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.dpc_pairdistance.argtypes = [C.c_int] * 3
        L.dpc_strerror.restype = C.c_char_p
        L.dpc_strerror.argtypes = [C.c_int]
        L.dpc_maxlengths.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.ctx = None

    _epoch = 0      # the library's init/setup state is process-wide: remember who registered last

    def init(self, mode=0, maxlookback=600, extraquerygap=10, maxpeelback=11, end=10, paired=8):
        self._init_args = (mode, maxlookback, extraquerygap, maxpeelback, end, paired)
        rc = self.lib.dpc_init(maxlookback, extraquerygap, maxpeelback, end, paired, mode)
        if rc != 0:
            raise RuntimeError("dpc_init: %s" % self.lib.dpc_strerror(rc).decode())
        CudaLib._epoch += 1

    def setup(self, setup):
        self._setup = setup
        rc = self.lib.dpc_setup(C.byref(setup))
        if rc != 0:
            raise RuntimeError("dpc_setup: %s" % self.lib.dpc_strerror(rc).decode())
        wl = getattr(setup, "_workload", None)
        self._genome_version = getattr(wl, "version", 0) if wl is not None else 0
        CudaLib._epoch += 1
        self._my_epoch = CudaLib._epoch

    def open(self, device=0):
        self.ctx = self.lib.dpc_ctx_new(device)
        if not self.ctx:
            raise RuntimeError("dpc_ctx_new(%d) failed: no usable CUDA device (no CPU fallback)" % device)
        return self

    def close(self):
        if self.ctx:
            self.lib.dpc_ctx_free(self.ctx)
            self.ctx = None

    def check(self, rc, what):
        if rc < 0:
            raise RuntimeError("%s: %s" % (what, self.lib.dpc_strerror(rc).decode()))
        return rc

    def refresh_genome(self):
        """The library mirrors the genome into HBM at dpc_setup; the synthetic generator for genome gaps
        edits the host genome afterwards (planted splice sites), so register it again when that happened."""
        wl = getattr(self._setup, "_workload", None)
        if self._my_epoch != CudaLib._epoch:          # another wrapper re-initialised the shared library
            mode, a, b, c, d, e = self._init_args
            self.init(mode, a, b, c, d, e)
            self.setup(self._setup)
        elif wl is not None and getattr(wl, "version", 0) != self._genome_version:
            self.setup(self._setup)

    def solve(self, problems, want_pairs=True):
        self.refresh_genome()

        def fn(p, n, r, pr, cap, off):
            return self.lib.dpc_solve(self.ctx, p, n, r, pr, cap, off)
        return _solve_common(fn, problems, want_pairs)

    def load(self, problems):
        """Ticket-API path on the context's own stream: add_bulk + flush + wait.  Leaves the batch resident in
        HBM (dpc_relaunch re-runs its kernels) and returns the results."""
        self.refresh_genome()
        L = self.lib
        self.check(L.dpc_reset(self.ctx), "dpc_reset")
        self.check(L.dpc_add_bulk(self.ctx, _ptr(problems), len(problems)), "dpc_add_bulk")
        self.check(L.dpc_flush(self.ctx), "dpc_flush")
        self.check(L.dpc_wait(self.ctx), "dpc_wait")
        res = np.zeros(len(problems), dtype=RESULT_DT)
        for i in range(len(problems)) if len(problems) <= 4096 else ():
            L.dpc_result(self.ctx, i, res[i:i + 1].ctypes.data_as(C.c_void_p))
        return res

    def solve_into(self, problems, results, pairs, pair_off):
        """dpc_solve into caller-owned (reusable) output arrays; returns the number of pair records."""
        self.refresh_genome()
        rc = self.lib.dpc_solve(self.ctx, _ptr(problems), len(problems), _ptr(results),
                                _ptr(pairs) if pairs is not None else None, len(pairs) if pairs is not None else 0, _ptr(pair_off))
        self.check(rc, "dpc_solve")
        return int(pair_off[len(problems)])

    def stats(self):
        s = Stats()
        self.check(self.lib.dpc_get_stats(self.ctx, C.byref(s)), "dpc_get_stats")
        return s

    def kernel_ms(self):
        ms = C.c_float()
        self.check(self.lib.dpc_last_kernel_ms(self.ctx, C.byref(ms)), "dpc_last_kernel_ms")
        return float(ms.value)

    def int_peak(self, device=0):
        """(ALU-pipe, fill-mix) giga lane-operations per second measured on `device` right now."""
        a, m = C.c_double(), C.c_double()
        self.check(self.lib.dpc_measure_int_peak(device, C.byref(a), C.byref(m)), "dpc_measure_int_peak")
        return float(a.value), float(m.value)

    def open_multi(self, devices):
        arr = (C.c_int * len(devices))(*devices)
        self.ctx = self.lib.dpc_ctx_new_multi(arr, len(devices))
        if not self.ctx:
            raise RuntimeError("dpc_ctx_new_multi(%r) failed: no usable CUDA device (no CPU fallback)" % (devices,))
        return self

    def register(self, array):
        """Page-locks a numpy array so that dpc_solve can copy results into it in place."""
        self.check(self.lib.dpc_host_register(_ptr(array), array.nbytes), "dpc_host_register")

    def unregister(self, array):
        self.check(self.lib.dpc_host_unregister(_ptr(array)), "dpc_host_unregister")


# --------------------------------------------------------------------------- comparison
RESULT_FIELDS = [n for n, _ in Result._fields_ if n != "reserved"]


def compare(res_a, pairs_a, off_a, res_b, pairs_b, off_b, rtol=1e-6):
    """Returns a list of human-readable mismatches (empty = identical)."""
    bad = []
    for f in RESULT_FIELDS:
        a, b = res_a[f], res_b[f]
        if f in ("left_prob", "right_prob"):
            ok = np.isclose(a, b, rtol=rtol, atol=0.0) | (a == b)
        else:
            ok = a == b
        for i in np.nonzero(~ok)[0][:5]:
            bad.append("problem %d field %s: %r vs %r" % (i, f, a[i], b[i]))
    if not np.array_equal(off_a, off_b):
        i = int(np.nonzero(off_a != off_b)[0][0])
        bad.append("pair offsets differ first at problem %d" % (i - 1))
    elif len(pairs_a):
        neq = pairs_a != pairs_b
        if neq.any():
            j = int(np.nonzero(neq)[0][0])
            i = int(np.searchsorted(off_a, j, side="right") - 1)
            bad.append("pair %d (problem %d, #%d): %r vs %r" % (j, i, j - off_a[i], pairs_a[j], pairs_b[j]))
    return bad


# --------------------------------------------------------------------------- (de)serialisation of problem sets
def _span(p):
    """(address of the first byte, length) of the query bytes problem p points at."""
    kind = int(p["kind"])
    if kind == END5_GAP:
        n = max(int(p["length1"]), 0)
        return int(p["seq1"]) - (n - 1), n
    if kind == CDNA_GAP:
        n = int(p["offset1R"]) - int(p["offset1"]) + 1
        return int(p["seq1"]), max(n, 0)
    return int(p["seq1"]), max(int(p["length1"]), 0)


def detach(problems):
    """Copies the query bytes out of process memory: returns (problems with pointers zeroed, qbuf, offsets)."""
    out = problems.copy()
    chunks, offs, at = [], np.zeros(len(problems), dtype=np.int64), 0
    for i, p in enumerate(problems):
        addr, n = _span(p)
        offs[i] = at
        if n > 0:
            chunks.append(C.string_at(addr, n))
            at += n
    out["seq1"] = 0
    out["seq1R"] = 0
    qbuf = np.frombuffer(b"".join(chunks) + b"\0" * 8, dtype=np.uint8).copy()
    return out, qbuf, offs


def attach(problems, qbuf, offs):
    """Inverse of detach: points the problems at qbuf (which the caller must keep alive)."""
    out = problems.copy()
    base = qbuf.ctypes.data
    for i in range(len(out)):
        kind = int(out["kind"][i])
        if kind == END5_GAP:
            out["seq1"][i] = base + offs[i] + max(int(out["length1"][i]), 0) - 1
        elif kind == CDNA_GAP:
            out["seq1"][i] = base + offs[i]
            out["seq1R"][i] = base + offs[i] + int(out["offset1R"][i]) - int(out["offset1"][i])
        else:
            out["seq1"][i] = base + offs[i]
    return out
