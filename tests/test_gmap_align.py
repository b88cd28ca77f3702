"""BASELINE config 1: the reference's own `make check` case align.test (`gmap -A -g ss.chr17test ss.her2`, golden
tests/align.test.ok in the reference tree), run through two GMAP binaries built by oracle/build_gmap.sh:
gmap_ref (unmodified reference) and gmap_cuda (the five gap-fill solvers replaced by libdynprog_cuda through
gmap-gsnap_b200/host/dynprog_dropin.c).  All 250 matrix fills of this alignment then happen on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
DATA = os.path.join(REFDIR, "align_test")


def run(binary):
    exe = os.path.join(REFDIR, binary)
    if not os.path.exists(exe) or not os.path.exists(os.path.join(DATA, "align.test.ok")):
        pytest.skip("oracle/_ref/%s not built (oracle/build_gmap.sh needs /root/reference)" % binary)
    return subprocess.run([exe, "-A", "-g", "ss.chr17test", "ss.her2"], cwd=DATA, capture_output=True, text=True, timeout=300)


def golden():
    return open(os.path.join(DATA, "align.test.ok")).read()


def test_reference_build_reproduces_align_test():
    r = run("gmap_ref")
    assert r.returncode == 0
    assert r.stdout == golden()


def test_offloaded_build_refuses_without_gpu():
    from gmap_gsnap_b200 import api
    if api.CudaLib().lib.dpc_device_count() > 0:
        pytest.skip("a GPU is present")
    r = run("gmap_cuda")
    assert r.returncode == 9                      # the reference's fatal-error exit code (gmap.c:2287-2308)
    assert "libdynprog_cuda" in r.stderr


@pytest.mark.gpu
def test_offloaded_build_reproduces_align_test():
    r = run("gmap_cuda")
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout == golden()


def synthetic_case(tmp_path, seed, n_transcripts=24, genome_len=150_000):
    """A random genomic segment and spliced, mutated transcripts of it (exons 80-400 bp, GT..AG introns 60-3000 bp,
    substitutions, small indels, both strands): FASTA files for `gmap -g`."""
    import numpy as np
    rng = np.random.default_rng(seed)
    nt = np.array(list("ACGT"))
    genome = nt[rng.integers(0, 4, genome_len)]
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    records = []
    for t in range(n_transcripts):
        pos = int(rng.integers(1000, genome_len - 40_000))
        exons = []
        for e in range(int(rng.integers(3, 9))):
            elen = int(rng.integers(80, 400))
            exons.append((pos, pos + elen))
            ilen = int(rng.integers(60, 3000))
            # canonical GT..AG (mostly), sometimes GC..AG or none
            kind = rng.integers(0, 10)
            if kind < 8:
                genome[pos + elen:pos + elen + 2] = list("GT"); genome[pos + elen + ilen - 2:pos + elen + ilen] = list("AG")
            elif kind == 8:
                genome[pos + elen:pos + elen + 2] = list("GC"); genome[pos + elen + ilen - 2:pos + elen + ilen] = list("AG")
            pos += elen + ilen
        records.append(exons)
    seqs = []
    for t, exons in enumerate(records):
        s = "".join("".join(genome[a:b]) for a, b in exons)
        out = []
        for ch in s:                                   # 1 % substitutions, 0.3 % deletions, 0.3 % insertions
            u = rng.random()
            if u < 0.003:
                continue
            if u < 0.006:
                out.append("ACGT"[rng.integers(0, 4)])
            if 0.006 <= u < 0.016:
                ch = "ACGT"[("ACGT".index(ch) + 1 + rng.integers(0, 3)) % 4]
            out.append(ch)
        s = "".join(out)
        if t % 2:
            s = "".join(comp[c] for c in reversed(s))
        if t % 5 == 0:
            s = "ACGTTGCA" * 3 + s + "TTGACCAG" * 2           # unalignable ends
        seqs.append(s)
    gfile, qfile = tmp_path / "genome.fa", tmp_path / "transcripts.fa"
    g = "".join(genome)
    gfile.write_text(">seg\n" + "\n".join(g[i:i + 60] for i in range(0, len(g), 60)) + "\n")
    qfile.write_text("".join(">t%d\n%s\n" % (i, "\n".join(s[j:j + 60] for j in range(0, len(s), 60))) for i, s in enumerate(seqs)))
    return str(gfile), str(qfile)


def run_files(binary, gfile, qfile, extra=()):
    exe = os.path.join(REFDIR, binary)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s not built (oracle/build_gmap.sh needs /root/reference)" % binary)
    return subprocess.run([exe, "-A", *extra, "-g", gfile, qfile], capture_output=True, text=True, timeout=900)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2])
def test_synthetic_transcripts_identical_to_reference_gmap(tmp_path, seed):
    """Whole-program differential test: every Dynprog_* call stage 3 makes for spliced, mutated transcripts
    (single gaps, genome gaps incl. finalp with the real MaxEnt hook, cDNA gaps, end gaps) goes to the GPU in
    gmap_cuda; the alignments must be byte-identical to the unmodified reference binary's."""
    gfile, qfile = synthetic_case(tmp_path, seed)
    ref = run_files("gmap_ref", gfile, qfile)
    got = run_files("gmap_cuda", gfile, qfile, extra=("-O",))       # -O: ordered output (the worker runs many requests at once)
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    assert ref.stdout.count(">t") >= 20 and "Alignments:" in ref.stdout
    assert got.stdout == ref.stdout


def test_synthetic_case_runs_on_reference_gmap(tmp_path):
    gfile, qfile = synthetic_case(tmp_path, 1, n_transcripts=4)
    ref = run_files("gmap_ref", gfile, qfile)
    assert ref.returncode == 0
    assert ref.stdout.count("Alignments:") == 4


@pytest.mark.gpu
def test_worker_threads_each_with_their_own_context(tmp_path):
    """gmap -t 4: four worker threads, each with its own thread-local dpc_ctx_t (like the per-thread Dynprog_T of
    gmap.c:2270), ordered output: identical to the single-threaded reference."""
    gfile, qfile = synthetic_case(tmp_path, 3)
    ref = run_files("gmap_ref", gfile, qfile)
    got = run_files("gmap_cuda", gfile, qfile, extra=("-t", "4", "-O"))
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    assert got.stdout == ref.stdout


# ---- the drop-in's fiber scheduler (host side of stage 3 batching) -------------------------------------------
MOCK = os.path.join(ROOT, "tests", "emul", "_mock")


def run_files_env(binary, gfile, qfile, extra=(), env_extra=None):
    exe = os.path.join(REFDIR, binary)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/%s not built (oracle/build_gmap.sh needs /root/reference)" % binary)
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([exe, "-A", *extra, "-g", gfile, qfile], capture_output=True, text=True, timeout=900, env=env)


def test_fiber_scheduler_on_cpu_test_double(tmp_path):
    """Host logic of the batched drop-in without a GPU: gmap_cuda is run against tests/emul/_mock (the ticket API
    answered by the compiled reference).  Two worker threads x 8 fibers collect stage 3's gaps into batches
    (double-buffered contexts, ticket bookkeeping, per-fiber stacks and Pairpools); the alignments must be
    byte-identical to the unmodified reference binary's, and batches must hold more than one gap."""
    if not os.path.exists(os.path.join(MOCK, "libdynprog_cuda.so")) or not os.path.exists(os.path.join(REFDIR, "libdynprog_ref.so")):
        pytest.skip("test double not built")
    gfile, qfile = synthetic_case(tmp_path, 4, n_transcripts=16, genome_len=120_000)
    ref = run_files("gmap_ref", gfile, qfile)
    env = {"LD_LIBRARY_PATH": MOCK, "DPC_FIBERS": "8", "DPC_FIBER_STATS": "1"}
    got = run_files_env("gmap_cuda", gfile, qfile, extra=("-t", "2", "-O"), env_extra=env)
    assert ref.returncode == 0 and got.returncode == 0, got.stderr[-2000:]
    assert got.stdout == ref.stdout
    stats = [l for l in got.stderr.splitlines() if "device batches" in l]
    assert len(stats) == 2
    per_batch = [float(l.split("(")[1].split(" per batch")[0]) for l in stats]
    assert max(per_batch) > 1.5
    # DPC_FIBERS=1: the plain worker thread, one gap per flush
    env["DPC_FIBERS"] = "1"
    one = run_files_env("gmap_cuda", gfile, qfile, extra=("-t", "2", "-O"), env_extra=env)
    assert one.returncode == 0 and one.stdout == ref.stdout


@pytest.mark.gpu
def test_database_run_with_fibers_identical_to_reference(tmp_path):
    """BASELINE configs[4] in small: a synthetic genome database (built with the reference's gmapindex), 2-kb
    spliced transcripts, gmap -t 4 with 16 fibers per thread on the GPU vs the unmodified reference."""
    import filecmp
    from gmap_gsnap_b200 import gmap_e2e as g
    if not g.have_binaries():
        pytest.skip("oracle/_ref binaries not built")
    case = g.prepare(str(tmp_path), 4_000_000, 4, 300, seed=9)
    _, ref_out, _ = g.run_gmap("gmap_ref", case, 4)
    _, got_out, err = g.run_gmap("gmap_cuda", case, 4, fibers=16)
    assert filecmp.cmp(ref_out, got_out, shallow=False)
    assert "device batches" in err


def _known_splicing_case(tmp_path):
    from gmap_gsnap_b200 import gmap_e2e as g
    if not g.have_binaries() or not os.path.exists(os.path.join(g.BIN, "iit_store")):
        pytest.skip("oracle/_ref binaries not built")
    return g, g.prepare(str(tmp_path), 4_000_000, 4, 100, seed=13, known_sites_frac=0.7)


def _junction_solves(err):
    return sum(int(l.split("; ")[-1].split(" splice-junction")[0]) for l in err.splitlines() if "splice-junction solves" in l)


def test_known_splicing_run_on_cpu_test_double(tmp_path):
    """gmap -s <known splice sites>: Splicetrie_solve_end5/3 (splicetrie.c:306-520) call Dynprog_end5/3_splicejunction
    once per candidate far site, and the genome gaps get their known-site rewards through the splice_known hook.
    Host logic on the CPU test double; byte-identical to the unmodified reference."""
    import filecmp
    if not os.path.exists(os.path.join(MOCK, "libdynprog_cuda.so")) or not os.path.exists(os.path.join(REFDIR, "libdynprog_ref.so")):
        pytest.skip("test double not built")
    g, case = _known_splicing_case(tmp_path)
    _, ref_out, _ = g.run_gmap("gmap_ref", case, 2)
    old = os.environ.get("LD_LIBRARY_PATH")
    os.environ["LD_LIBRARY_PATH"] = MOCK
    try:
        _, got_out, err = g.run_gmap("gmap_cuda", case, 2, fibers=8, out=os.path.join(str(tmp_path), "mock.out"))
    finally:
        if old is None:
            del os.environ["LD_LIBRARY_PATH"]
        else:
            os.environ["LD_LIBRARY_PATH"] = old
    assert filecmp.cmp(ref_out, got_out, shallow=False)
    assert _junction_solves(err) > 0


@pytest.mark.gpu
def test_known_splicing_run_identical_to_reference(tmp_path):
    """The same run with the real library: the splice-junction solvers (kinds 5/6) and the known-site genome gaps
    on the GPU, inside an unmodified stage 3 / splicetrie."""
    import filecmp
    g, case = _known_splicing_case(tmp_path)
    _, ref_out, _ = g.run_gmap("gmap_ref", case, 4)
    _, got_out, err = g.run_gmap("gmap_cuda", case, 4, fibers=16)
    assert filecmp.cmp(ref_out, got_out, shallow=False)
    assert _junction_solves(err) > 0
