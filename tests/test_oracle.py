"""CPU suite, part 1: the oracle is pinned (restatement == compiled reference == golden vectors) and the
device routines, compiled single-lane for the CPU (tests/emul), agree with it."""
import numpy as np
import pytest

from gmap_gsnap_b200 import api
from oracle import checkers
from conftest import has_ref
from util import GOLDEN_SETS, SPLICING_IIT_MODES, Golden, long_nogaps_ends, mixed_problems, splicing_iit_hooks


@pytest.fixture(scope="module")
def golden():
    return Golden()


@pytest.fixture(scope="module")
def golden_port(golden):
    o = checkers.PortOracle()
    o.init()
    o.setup(golden.setup())
    return o


@pytest.fixture(scope="module")
def golden_emul(golden):
    e = checkers.EmulLib()
    e.init()
    e.setup(golden.setup())
    return e


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_restatement_matches_golden(golden, golden_port, name):
    got = golden_port.solve(golden.problems(name))
    assert not api.compare(*golden.expected(name), *got)
    assert not golden.missing


@pytest.mark.parametrize("fill", [0, 1], ids=["row_sweep_32_lanes", "memory_fill_1_lane"])
@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_device_routines_on_cpu_match_golden(golden, golden_emul, name, fill):
    golden_emul.set_fill(fill)
    try:
        got = golden_emul.solve(golden.problems(name))
    finally:
        golden_emul.set_fill(0)
    assert not api.compare(*golden.expected(name), *got)
    assert not golden.missing


def test_golden_has_negative_cases(golden):
    res, _, _ = golden.expected("edge")
    assert res["null_list"].sum() >= 20
    assert (res["finalscore"] == -10000).sum() >= 2          # "too long" sentinel, dynprog.c:4512
    assert (res["finalscore"] == -1000000).sum() >= 2        # NEG_INFINITY sentinel, dynprog.c:4857


@pytest.mark.parametrize("seed", [1, 2])
def test_restatement_matches_compiled_reference(workload, ref, port, seed):
    probs = checkers.arm_probability_mode(mixed_problems(workload, 1200, seed, long_frac=0.03, long_hi=611), ref)
    assert not api.compare(*ref.solve(probs), *port.solve(probs))


def test_probability_mode_matches_reference(workload, ref, port, emul):
    probs = workload.genome_gaps(800, seed=5, finalp_mode=2, prob_mode_pm=1000, long_frac=0.0)
    probs = checkers.arm_probability_mode(probs, ref)
    assert (probs["use_probabilities_p"] == 1).all()
    want = ref.solve(probs)
    assert not api.compare(*want, *port.solve(probs))
    assert not api.compare(*want, *emul.solve(probs))


@pytest.mark.parametrize("fill", [0, 1], ids=["row_sweep_32_lanes", "memory_fill_1_lane"])
def test_device_routines_on_cpu_match_oracle(workload, port, emul, fill):
    probs = checkers.arm_probability_mode(mixed_problems(workload, 1500, 9, long_frac=0.03, long_hi=611), port)
    emul.set_fill(fill)
    try:
        got = emul.solve(probs)
    finally:
        emul.set_fill(0)
    assert not api.compare(*port.solve(probs), *got)


def test_band_widths_around_the_chunk_limits(workload, port, emul):
    """The row sweep gives every lane 1, 2 or 4 diagonals: bands of 31..34, 63..66 and 127..130 diagonals (the last
    ones fall back to the memory-state fill), with and without jump_late_p."""
    sets = []
    for eb in (15, 16, 31, 32, 63, 64):
        p = workload.single_gaps(120, extraband=eb, seed=1000 + eb)
        p["length2"] = np.maximum(p["length1"] + np.arange(len(p)) % 4 - 1, 1)      # W = 2 eb + 1 + |L2 - L1|
        sets.append(p)
    allp = np.concatenate(sets)
    assert not api.compare(*port.solve(allp), *emul.solve(allp))


def test_wide_bands_and_max_sizes(workload, port, emul):
    """Bands of 64+ diagonals and the largest matrices Dynprog_T allows (611 x 2000, dynprog.c:831-852)."""
    probs = workload.single_gaps(40, extraband=30, seed=21)
    probs["length2"][:20] = np.minimum(probs["length1"][:20] + np.arange(20) * 7 + 30, 2000)
    probs["extraband"][20:30] = 64
    big = workload.single_gaps(3, extraband=3, seed=22, len_lo=580, len_hi=600, p_del=0.002, p_ins=0.002)
    big["length2"][0] = 2000
    allp = np.concatenate([probs, big])
    assert not api.compare(*port.solve(allp), *emul.solve(allp))


def test_known_splice_sites(workload, prob_hook):
    """splicing_iit != NULL: +20 at known sites, and with novelsplicingp == false only known-known introns pass."""
    known = api.KNOWN_FN(lambda which, chrnum, pos, sign, user: int((pos * 7 + which) % 11 == 0))
    for novel in (1, 0):
        s = workload.make_setup(splice_prob=prob_hook, splice_known=known, novelsplicingp=novel)
        o, e = checkers.PortOracle(), checkers.EmulLib()
        o.init(); e.init()
        o.setup(s); e.setup(s)
        probs = workload.genome_gaps(300, seed=31 + novel, finalp_mode=2, long_frac=0.0)
        want, got = o.solve(probs), e.solve(probs)
        assert not api.compare(*want, *got)
        if not novel:
            assert want[0]["null_list"].sum() > 0


def test_long_nogaps_ends(workload, ref, port, emul):
    """QUERYEND_NOGAPS ends are not clipped to maxlength (dynprog.c:5179): 16 384+ columns need more than one
    traceback op; beyond 38 ops the problem is refused, never truncated."""
    probs = long_nogaps_ends(workload)
    want = ref.solve(probs)
    assert (want[0]["npairs"] >= 16383).all()
    assert not api.compare(*want, *port.solve(probs))
    assert not api.compare(*want, *emul.solve(probs))
    probs["length1"][0] = probs["length2"][0] = 38 * 16383 + 1
    with pytest.raises(RuntimeError):
        emul.solve(probs[:1])


@pytest.mark.parametrize("intron_level,novel", SPLICING_IIT_MODES)
def test_splicing_iit_modes_match_compiled_reference(workload, intron_level, novel):
    """Every flavour of splicing_iit != NULL in bridge_intron_gap: splice-site level (dynprog.c:3377-3458) and intron
    level (3460-3542), novel splicing allowed or not -- the latter with an intron-level IIT is the scan constrained
    to the given introns (3552-3696).  Compiled reference == restatement == device routines on the CPU; a tenth of
    the problems ask for use_probabilities_p (ignored by the constrained scan)."""
    if not has_ref():
        pytest.skip("compiled reference not built")
    known, intron = splicing_iit_hooks(known_mod=(3 if novel else 17) if intron_level else 3, intron_mod=3)
    r = checkers.RefOracle()
    r.init()
    s = workload.make_setup(splice_prob=r.splice_prob, splice_known=known, novelsplicingp=novel,
                            splice_intron=intron, intron_level=intron_level)
    o, e = checkers.PortOracle(), checkers.EmulLib()
    o.init(); e.init()
    r.setup(s); o.setup(s); e.setup(s)
    probs = workload.genome_gaps(400, seed=41 + 2 * intron_level + novel, finalp_mode=2, prob_mode_pm=100, long_frac=0.02, long_hi=300)
    probs = checkers.arm_probability_mode(probs, o)
    want = r.solve(probs)
    assert not api.compare(*want, *o.solve(probs))
    assert not api.compare(*want, *e.solve(probs))
    nulls = int(want[0]["null_list"].sum())
    assert len(probs) - nulls > 20 and (novel or nulls > 0), nulls


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_pairdistance_tables(mode):
    """pairdistance_init (dynprog.c:1127-1226) for every Mode_T: product table == restatement table,
    including the 'z' / 'Z' loop-bound quirks."""
    o, lib = checkers.PortOracle(), api.CudaLib()
    o.init(mode=mode)
    lib.init(mode=mode)
    for t in range(4):
        a = np.array([[o.lib.port_pairdistance(t, i, j) for j in range(128)] for i in range(128)])
        b = np.array([[lib.lib.dpc_pairdistance(t, i, j) for j in range(128)] for i in range(128)])
        assert (a == b).all()
    assert lib.lib.dpc_pairdistance(0, ord("N"), ord("N")) == 3
    assert lib.lib.dpc_pairdistance(0, ord("A"), ord("z")) == 0
    assert lib.lib.dpc_pairdistance(0, ord("Z"), ord("Z")) == -3
    assert lib.lib.dpc_pairdistance(3, ord("A"), ord("C")) == -5


def test_pairdistance_matches_compiled_reference(ref):
    lib = api.CudaLib()
    lib.init(mode=0)
    ref.init(mode=0)
    for i in range(128):
        for j in range(128):
            assert ref.lib.ref_pairdistance(i, j) == lib.lib.dpc_pairdistance(0, i, j)   # Dynprog_pairdistance = HIGHQ table


# ---- Dynprog_end5/3_splicejunction (kinds 5, 6) -------------------------------------------
def test_splicejunction_restatement_matches_compiled_reference(workload, ref, port):
    probs = workload.splicejunction_gaps(3000, seed=21)
    want = ref.solve(probs)
    assert (want[1]["gapp"] == 2).sum() == len(probs)            # one known gapholder per solution (dynprog.c:5518)
    assert (want[1]["gapp"] == 1).sum() > 0                      # and some intron-like genome runs (2448-2507)
    assert not api.compare(*want, *port.solve(probs))


@pytest.mark.parametrize("fill", [0, 1], ids=["row_sweep_32_lanes", "memory_fill_1_lane"])
def test_splicejunction_device_routines_on_cpu(workload, port, emul, fill):
    probs = workload.splicejunction_gaps(3000, seed=22, len_hi=90)
    emul.set_fill(fill)
    try:
        got = emul.solve(probs)
    finally:
        emul.set_fill(0)
    assert not api.compare(*port.solve(probs), *got)


def test_splicejunction_edges(workload, port, emul):
    """contlength 0 / beyond the string, too-long inputs (early return without index bump, dynprog.c:5452-5465),
    junction strings with a byte outside ACGTN (rejected by the library, never guessed)."""
    probs = workload.splicejunction_gaps(64, seed=23)
    probs["length2R"][:16] = 0
    probs["length2R"][16:32] = probs["length2"][16:32] + 5
    want = port.solve(probs)
    assert not api.compare(*want, *emul.solve(probs))
    big = probs[:4].copy()
    big["length1"] = 700
    w2, g2 = port.solve(big), emul.solve(big)
    assert not api.compare(*w2, *g2)
    assert (w2[0]["null_list"] == 1).all() and (w2[0]["finalscore"] == 0).all()
    assert (w2[0]["dynprogindex_out"] == big["dynprogindex"]).all()
    bad = probs[32:33].copy()
    import ctypes as C
    addr = int(bad["seq1R"][0]) - (int(bad["length2"][0]) - 1 if int(bad["kind"][0]) == api.END5_SPLICEJUNCTION else 0)
    C.memset(addr, ord("*"), 1)
    with pytest.raises(RuntimeError):
        emul.solve(bad)


def test_rebuild_without_the_staged_genome_stream(workload, port, emul):
    """The Pair rebuild normally reads the genome characters the device staged and returned; without that stream it
    decodes the 2-bit genome itself (gather_genome).  Both give the reference's records."""
    probs = mixed_problems(workload, 800, 41, long_frac=0.03, long_hi=400)
    probs = probs[probs["use_probabilities_p"] == 0]
    emul.set_fill(2)            # bit 1: no staged-genome stream
    try:
        got = emul.solve(probs)
    finally:
        emul.set_fill(0)
    assert not api.compare(*port.solve(probs), *got)


def _hookless(probs):
    """Problems the device pipeline takes: no use_probabilities_p (needs the host hook per position)."""
    return probs[probs["use_probabilities_p"] == 0]


@pytest.mark.parametrize("fill", [0, 1], ids=["row_sweep", "memory_fill"])
def test_device_pipeline_routines_on_cpu(workload, port, emul, golden, golden_port, golden_emul, fill):
    """dpc_pipe.h -- the per-problem routines of dpc_solve's device pipeline (argument checks and early returns,
    descriptor packing, result finalisation, Pair-record expansion with 1 and with 32 lanes, staged and re-decoded
    genome characters) -- compiled for the CPU, against the restatement: mixed batches, every golden set including
    the 232 edge cases, long NOGAPS ends, band widths around the lane-chunk limits."""
    try:
        for e, o, sets in ((emul, port, None), (golden_emul, golden_port, GOLDEN_SETS)):
            e.setup(e._setup)          # the checkers' state is process-wide: register this pair's genome
            o.setup(o._setup)
            e.set_fill(fill)
            e.set_path(1)
            if sets is None:
                probs = _hookless(mixed_problems(workload, 1500, 909, long_frac=0.04, long_hi=611))
                assert not api.compare(*o.solve(probs), *e.solve(probs))
                probs = long_nogaps_ends(workload, lengths=(16383, 16384, 20000))
                assert not api.compare(*o.solve(probs), *e.solve(probs))
                probs = workload.single_gaps(300, extraband=30, seed=77, edge_frac_pm=100, lower_case=1, iupac_pm=20)
                probs["extraband"] = np.arange(300) % 70
                assert not api.compare(*o.solve(probs), *e.solve(probs))
                probs = workload.end_gaps(3000, seed=78, edge_frac_pm=300, lower_case=1, iupac_pm=20)
                for ea in range(4):
                    probs["endalign"][ea::5] = ea
                assert not api.compare(*o.solve(probs), *e.solve(probs))
            else:
                for name in sets:
                    probs = golden.problems(name)
                    keep = probs["use_probabilities_p"] == 0
                    want = golden.expected(name)
                    got = e.solve(probs[keep])
                    ref = o.solve(probs[keep])
                    assert not api.compare(*ref, *got), name
                    assert (want[0][keep] == got[0]).all() or name == "genome", name
    finally:
        emul.set_fill(0)
        emul.set_path(0)
        emul.setup(emul._setup)
        port.setup(port._setup)


def test_microexon_search(ref, port, emul):
    """Dynprog_microexon_int (dynprog.c:7127-7429): compiled reference == restatement == the
    library's host logic around the (here CPU-run) device scan.  Planted microexons of 3-12 nt with GT..AG / CT..AC
    introns, decoy copies, flank mismatches, N and lower-case query letters, problems without a microexon."""
    w = api.Workload(3_000_000, seed=23, nchr=3)
    probs = w.microexon_problems(600, seed=92)
    r = checkers.RefOracle()
    r.init()
    r.setup(w.make_setup())                                   # Maxent_hr needs the genome registered first
    s = w.make_setup(splice_prob=r.splice_prob)
    r.setup(s)
    o, e = checkers.PortOracle(), checkers.EmulLib()
    o.init(); e.init()
    o.setup(s); e.setup(s)
    want = r.solve(probs)
    found = int((want[0]["null_list"] == 0).sum())
    assert 300 < found < 600 and ((want[1]["gapp"] == 1) & ((want[1]["comp"] == b">") | (want[1]["comp"] == b"<"))).sum() == 2 * found
    assert not api.compare(*want, *o.solve(probs))
    assert not api.compare(*want, *e.solve(probs))
