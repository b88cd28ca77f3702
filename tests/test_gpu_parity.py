"""GPU suite: libdynprog_cuda, called through its C ABI, against the oracle on the same seeded inputs.
Bit-exact for scores, end points, intron boundaries, counts and Pair records; 1e-6 relative for the
splice-site probabilities (the tolerance BASELINE.json's north_star states)."""
import ctypes as C
import os

import numpy as np
import pytest

from gmap_gsnap_b200 import api
from oracle import checkers
from util import GOLDEN_SETS, SPLICING_IIT_MODES, Golden, long_nogaps_ends, mixed_problems, splicing_iit_hooks

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def golden():
    return Golden()


@pytest.fixture(scope="module", autouse=True, params=[0, 1], ids=["auto_path", "host_half"])
def solve_path(request):
    """dpc_solve serves a chunk with its device pipeline (kernels for packing / finalisation / Pair expansion) or
    with the host half (hooks, splice-junction solvers).  Every test of this module runs once with the library's own
    choice per chunk and once with the host half forced; both must match the oracle."""
    lib = api.CudaLib()
    lib.lib.dpc_set_path(request.param)
    yield request.param
    lib.lib.dpc_set_path(0)


@pytest.fixture(scope="module", params=[0, 1], ids=["register_fill", "memory_fill"])
def fill_mode(request, cuda):
    cuda.lib.dpc_set_fill(request.param)
    yield request.param
    cuda.lib.dpc_set_fill(0)


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_golden_vectors(golden, name):
    lib = api.CudaLib()
    lib.init()
    lib.setup(golden.setup())
    lib.open(0)
    try:
        for force in (0, 1):
            lib.lib.dpc_set_fill(force)
            got = lib.solve(golden.problems(name))
            assert not api.compare(*golden.expected(name), *got, rtol=RTOL)
    finally:
        lib.lib.dpc_set_fill(0)
        lib.close()
    assert not golden.missing


def test_mixed_batch_matches_oracle(workload, port, cuda, fill_mode):
    probs = checkers.arm_probability_mode(mixed_problems(workload, 4000, 101, long_frac=0.03, long_hi=611), port)
    assert not api.compare(*port.solve(probs), *cuda.solve(probs), rtol=RTOL)


def test_matches_compiled_reference(workload, ref, cuda):
    probs = checkers.arm_probability_mode(mixed_problems(workload, 3000, 202), ref)
    assert not api.compare(*ref.solve(probs), *cuda.solve(probs), rtol=RTOL)


@pytest.mark.parametrize("extraband", [0, 3, 30])
def test_single_gap_bands(workload, port, cuda, fill_mode, extraband):
    probs = workload.single_gaps(6000, extraband=extraband, seed=300 + extraband, edge_frac_pm=20, lower_case=1, iupac_pm=5)
    assert not api.compare(*port.solve(probs), *cuda.solve(probs))


def test_end_gaps_all_endalign(workload, port, cuda, fill_mode):
    probs = workload.end_gaps(12000, seed=401, edge_frac_pm=30, lower_case=1, iupac_pm=10)
    for ea in range(4):
        probs["endalign"][ea::7] = ea
    assert not api.compare(*port.solve(probs), *cuda.solve(probs))


def test_genome_gaps_modes(workload, port, cuda, fill_mode):
    probs = workload.genome_gaps(5000, seed=501, finalp_mode=2, prob_mode_pm=200, long_frac=0.08, long_hi=611, iupac_pm=3)
    probs = checkers.arm_probability_mode(probs, port)
    want, got = port.solve(probs), cuda.solve(probs)
    assert not api.compare(*want, *got, rtol=RTOL)
    assert want[0]["null_list"].sum() > 0 and (want[0]["null_list"] == 0).sum() > 0


def test_cdna_gaps(workload, port, cuda, fill_mode):
    probs = workload.cdna_gaps(1500, seed=601)
    assert not api.compare(*port.solve(probs), *cuda.solve(probs))


def test_wide_bands_and_max_sizes(workload, port, cuda):
    """Bands of 64+ diagonals (memory-state fill inside the default kernel), problems that need HBM scratch,
    and the largest matrices Dynprog_T allows (611 x 2000)."""
    probs = workload.single_gaps(60, extraband=30, seed=21)
    probs["length2"][:20] = np.minimum(probs["length1"][:20] + np.arange(20) * 7 + 30, 2000)
    probs["extraband"][20:30] = 64
    big = workload.single_gaps(4, extraband=3, seed=22, len_lo=580, len_hi=600, p_del=0.002, p_ins=0.002)
    big["length2"][0] = 2000
    gg = workload.genome_gaps(6, seed=23, long_frac=1.0, long_hi=611, len_hi=560)
    allp = np.concatenate([probs, big, gg])
    assert not api.compare(*port.solve(allp), *cuda.solve(allp))


def test_known_splice_sites(workload, prob_hook):
    known = api.KNOWN_FN(lambda which, chrnum, pos, sign, user: int((pos * 7 + which) % 11 == 0))
    for novel in (1, 0):
        s = workload.make_setup(splice_prob=prob_hook, splice_known=known, novelsplicingp=novel)
        o, lib = checkers.PortOracle(), api.CudaLib()
        o.init(); lib.init()
        o.setup(s); lib.setup(s)
        lib.open(0)
        try:
            probs = workload.genome_gaps(400, seed=31 + novel, finalp_mode=2, long_frac=0.0)
            assert not api.compare(*o.solve(probs), *lib.solve(probs), rtol=RTOL)
        finally:
            lib.close()


@pytest.mark.parametrize("intron_level,novel", SPLICING_IIT_MODES)
def test_splicing_iit_modes(workload, ref, intron_level, novel):
    """splicing_iit != NULL in all four flavours (splice-site / intron level x novel splicing allowed or not); the
    last one is the bridge constrained to the given introns, dynprog.c:3552-3696.  Against the compiled reference."""
    known, intron = splicing_iit_hooks(known_mod=(3 if novel else 17) if intron_level else 3, intron_mod=3)
    r = checkers.RefOracle()
    r.init()
    s = workload.make_setup(splice_prob=r.splice_prob, splice_known=known, novelsplicingp=novel,
                            splice_intron=intron, intron_level=intron_level)
    r.setup(s)
    lib = api.CudaLib()
    lib.init()
    lib.setup(s)
    lib.open(0)
    try:
        probs = workload.genome_gaps(1500, seed=51 + 2 * intron_level + novel, finalp_mode=2, prob_mode_pm=100, long_frac=0.03, long_hi=611)
        probs = checkers.arm_probability_mode(probs, r)
        for force in (0, 1):
            lib.lib.dpc_set_fill(force)
            assert not api.compare(*r.solve(probs), *lib.solve(probs), rtol=RTOL)
    finally:
        lib.lib.dpc_set_fill(0)
        lib.close()


def test_long_nogaps_ends(workload, ref, cuda):
    """QUERYEND_NOGAPS ends of 16 383 .. 40 000 columns (several M ops per problem), bulk and ticket API."""
    probs = long_nogaps_ends(workload)
    want = ref.solve(probs)
    got = cuda.solve(probs)
    assert not api.compare(*want, *got)
    L = cuda.lib
    cuda.refresh_genome()
    assert L.dpc_reset(cuda.ctx) == 0
    assert L.dpc_add_bulk(cuda.ctx, probs.ctypes.data_as(C.c_void_p), len(probs)) == 0
    assert L.dpc_flush(cuda.ctx) == 0 and L.dpc_wait(cuda.ctx) == 0
    buf = np.zeros(50000, dtype=api.PAIR_DT)
    for t in range(len(probs)):
        n = L.dpc_pairs(cuda.ctx, t, buf.ctypes.data_as(C.c_void_p), len(buf))
        assert n == want[2][t + 1] - want[2][t] and (buf[:n] == want[1][want[2][t]:want[2][t + 1]]).all()
    assert L.dpc_reset(cuda.ctx) == 0


def test_empty_and_degenerate_batches(workload, port, cuda):
    probs = workload.end_gaps(8, seed=7)
    res, pairs, off = cuda.solve(probs[:0])
    assert len(res) == 0 and len(pairs) == 0 and off.tolist() == [0]
    probs["length1"][:] = 0                      # every problem resolves on the host side of the boundary (5140)
    assert not api.compare(*port.solve(probs), *cuda.solve(probs))


def test_ticket_api_equals_bulk_api(workload, cuda):
    """dpc_add / dpc_flush / dpc_wait / dpc_result / dpc_pairs (what a modified stage3.c calls) == dpc_solve."""
    probs = mixed_problems(workload, 200, 77)
    want_res, want_pairs, want_off = cuda.solve(probs)
    L = cuda.lib
    cuda.refresh_genome()
    assert L.dpc_reset(cuda.ctx) == 0
    tickets = [L.dpc_add(cuda.ctx, probs[i:i + 1].ctypes.data_as(C.c_void_p)) for i in range(len(probs))]
    assert tickets == list(range(len(probs)))
    assert L.dpc_flush(cuda.ctx) == 0 and L.dpc_wait(cuda.ctx) == 0
    r = np.zeros(1, dtype=api.RESULT_DT)
    buf = np.zeros(8192, dtype=api.PAIR_DT)
    for t in tickets:
        assert L.dpc_result(cuda.ctx, t, r.ctypes.data_as(C.c_void_p)) == 0
        assert r[0] == want_res[t]
        n = L.dpc_pairs(cuda.ctx, t, buf.ctypes.data_as(C.c_void_p), len(buf))
        assert n == want_off[t + 1] - want_off[t]
        assert (buf[:n] == want_pairs[want_off[t]:want_off[t + 1]]).all()
    assert L.dpc_reset(cuda.ctx) == 0


def test_full_size_properties(workload, port, cuda):
    """BASELINE config 2 at full size (1 M single gaps, band 30): too big for the oracle in seconds, so
    check size-independent properties plus an exact comparison on a strided sample."""
    n = 1_000_000
    probs = workload.single_gaps(n, extraband=30, seed=4242)
    res, _, off = cuda.solve(probs, want_pairs=False)
    assert (res["null_list"] == 0).all()
    assert (res["dynprogindex_out"] == probs["dynprogindex"] + np.sign(probs["dynprogindex"])).all()
    # every aligned column is a match or a mismatch and every run is an indel, so the pairs add up to the path and
    # the global path consumes both sequences exactly -- except where a 9+ genome run became a gapholder or a
    # column fell off the segment ('*'); those few are compared with the oracle one by one
    cols = res["nmatches"] + res["nmismatches"]
    plain = (cols + res["nindels"] == res["npairs"]) & (2 * cols + res["nindels"] == probs["length1"] + probs["length2"])
    assert plain.mean() > 0.99
    odd = probs[~plain][:2000]
    if len(odd):
        assert not api.compare(*port.solve(odd), *cuda.solve(odd))
    # score bound: 3 per match, at least -3 per mismatch (HIGHQ/MEDQ/LOWQ), gaps cost open + n*extend <= -13
    assert (res["finalscore"] <= 3 * res["nmatches"]).all()
    # idempotence: solving again gives the same answer
    res2, _, _ = cuda.solve(probs, want_pairs=False)
    assert (res == res2).all()
    sample = probs[:: n // 5000]
    assert not api.compare(*port.solve(sample), *cuda.solve(sample))


@pytest.mark.parametrize("kind,n", [("single", 1_000_000), ("genome", 500_000), ("genome_mix", 500_000), ("end", 1_000_000)])
def test_full_size_exact_against_compiled_reference(kind, n, solve_path):
    """BASELINE configs[1..3] at their full sizes, EXACT: the unmodified reference (oracle/_ref/libdynprog_ref.so,
    all host threads) solves the same problems as the GPU and every output field must be equal -- scores, counts,
    intron boundaries, introntype, splice-site probabilities (rtol 1e-6), npairs, dynprogindex; the Pair records are
    compared on the first 100 000.  genome_mix is config 3 with its stated mix: finalp and halfp both ways and a
    10 % subset re-solved with use_probabilities_p and score_threshold = first-pass finalscore - 11 (stage3.c:5833)."""
    import bench
    if not os.path.exists(os.path.join(bench.ROOT, "oracle", "_ref", "libdynprog_ref.so")):
        pytest.skip("compiled reference not built")
    mix = kind == "genome_mix"
    w, probs = bench.make_workload(0, n, "genome" if mix else kind, genome_mix=mix)
    ref = checkers.RefOracle()
    ref.init()
    ref.setup(w.make_setup())
    hook = ref.splice_prob if mix else None        # finalp / probability mode need the MaxEnt hook
    ref.setup(w.make_setup(splice_prob=hook))
    lib = api.CudaLib()
    lib.init()
    lib.setup(w.make_setup(splice_prob=hook))
    lib.open(0)
    if mix:
        probs = checkers.arm_probability_mode(probs, lib)
        assert (probs["use_probabilities_p"] == 1).sum() > n // 20 and probs["finalp"].sum() > n // 3
    try:
        got, _, _ = lib.solve(probs, want_pairs=False)
        want, _ = ref.solve_mt(probs, os.cpu_count() or 1)
        empty = np.zeros(0, dtype=api.PAIR_DT)
        z = np.zeros(1, dtype=np.int64)
        assert not api.compare(want, empty, z, got, empty, z, rtol=RTOL)
        head = probs[:100_000]
        want_head = ref.solve(head)
        assert not api.compare(*want_head, *lib.solve(head), rtol=RTOL)
        if solve_path == 0 and not mix:
            # the two output routes of the device pipeline, each forced for every chunk
            for path in (3, 4):
                lib.lib.dpc_set_path(path)
                assert not api.compare(*want_head, *lib.solve(head), rtol=RTOL)
            lib.lib.dpc_set_path(0)
    finally:
        lib.lib.dpc_set_path(solve_path)
        lib.close()


def test_splicejunction_solvers(workload, ref, port, cuda, fill_mode):
    """Dynprog_end5/3_splicejunction: junction string as characters, end point on the last
    row, two-part traceback around the known gapholder; against the compiled reference and the restatement."""
    probs = workload.splicejunction_gaps(6000, seed=31, len_hi=100)
    got = cuda.solve(probs)
    assert not api.compare(*ref.solve(probs), *got)
    assert not api.compare(*port.solve(probs), *got)
    mixed = np.concatenate([probs[:500], workload.end_gaps(500, seed=32), workload.single_gaps(500, seed=33)])
    assert not api.compare(*port.solve(mixed), *cuda.solve(mixed))


def test_device_pipeline_is_taken_and_exact(workload, ref, solve_path):
    """Hook-free batches must actually run the device pipeline (stats), with scattered queries (gathered), with one
    contiguous query buffer (range copy), with page-locked output arrays (copied in place) and without pair output."""
    if solve_path != 0:
        pytest.skip("host half forced")
    lib = api.CudaLib()
    lib.init()
    lib.setup(workload.make_setup())
    lib.open(0)
    o = checkers.RefOracle()
    o.init()
    o.setup(workload.make_setup())
    try:
        mixed = mixed_problems(workload, 3000, 4711, long_frac=0.03, long_hi=611)
        mixed = mixed[(mixed["use_probabilities_p"] == 0) & (mixed["finalp"] == 0)]
        want = o.solve(mixed)
        one = workload.single_gaps(40000, extraband=30, seed=4712, edge_frac_pm=10, lower_case=1, iupac_pm=5)
        want_one = o.solve(one)
        for path in (3, 4, 2):                           # pairs expanded on the device / by host threads / per chunk
            lib.lib.dpc_set_path(path)
            got = lib.solve(mixed)
            st = lib.stats()
            assert st.pipeline_chunks > 0 and st.host_chunks == 0
            assert not api.compare(*want, *got)
            assert not api.compare(*want_one, *lib.solve(one))
        want = want_one
        # caller-owned, page-locked outputs: the copy engine writes results and pairs in place
        res = np.zeros(len(one), dtype=api.RESULT_DT)
        pairs = np.zeros(len(want[1]) + 16, dtype=api.PAIR_DT)
        off = np.zeros(len(one) + 1, dtype=np.int64)
        lib.register(pairs); lib.register(res)
        try:
            for _ in range(2):
                pairs[:] = np.zeros(1, dtype=api.PAIR_DT)[0]
                assert lib.solve_into(one, res, pairs, off) == len(want[1])
                assert not api.compare(*want, res, pairs[:len(want[1])], off)
        finally:
            lib.unregister(pairs); lib.unregister(res)
        got = lib.solve(one, want_pairs=False)
        assert (got[0] == want[0]).all() and np.array_equal(got[2], want[2])
        # a chunk that needs a hook refuses the forced pipeline and is served by the host half otherwise
        probs = workload.genome_gaps(50, seed=4713, finalp_mode=0)
        probs["use_probabilities_p"] = 1
        with pytest.raises(RuntimeError):
            lib.solve(probs)
    finally:
        lib.lib.dpc_set_path(0)
        lib.close()


def test_multi_device_context(workload, ref):
    """dpc_ctx_new_multi: chunks dealt round-robin to the devices of the box, outputs in input order."""
    lib = api.CudaLib()
    ndev = lib.lib.dpc_device_count()
    if ndev < 2:
        pytest.skip("one device visible")
    lib.init()
    lib.setup(workload.make_setup())
    lib.open_multi(list(range(min(ndev, 4))))
    try:
        probs = np.concatenate([workload.single_gaps(60000, extraband=30, seed=4801), workload.end_gaps(30000, seed=4802)])
        o = checkers.RefOracle()
        o.init()
        o.setup(workload.make_setup())
        assert not api.compare(*o.solve(probs), *lib.solve(probs))
    finally:
        lib.close()


def test_microexon_search(solve_path):
    """Dynprog_microexon_int: exact-match scans on the device, flanks / MaxEnt / pair assembly on
    the host; against the compiled reference, alone and mixed into a batch of other kinds."""
    w = api.Workload(6_000_000, seed=29, nchr=3)
    probs = w.microexon_problems(1500, seed=93, span_hi=30000)
    r = checkers.RefOracle()
    r.init()
    r.setup(w.make_setup())
    s = w.make_setup(splice_prob=r.splice_prob)
    r.setup(s)
    lib = api.CudaLib()
    lib.init()
    lib.setup(s)
    lib.open(0)
    try:
        want = r.solve(probs)
        assert (want[0]["null_list"] == 0).sum() > 700
        assert not api.compare(*want, *lib.solve(probs), rtol=RTOL)
        mixed = np.concatenate([probs[:300], w.single_gaps(500, seed=94), w.end_gaps(500, seed=95)])
        mixed = mixed[np.random.default_rng(5).permutation(len(mixed))]
        assert not api.compare(*r.solve(mixed), *lib.solve(mixed), rtol=RTOL)
    finally:
        lib.close()
