"""Whole-program workload for BASELINE configs[4]: synthetic spliced transcripts against a synthetic genome database,
aligned by two GMAP binaries built by oracle/build_gmap.sh --

  oracle/_ref/gmap_ref    the unmodified reference (gmap.c worker threads, CPU dynprog.c), run with -t <host cores>;
  oracle/_ref/gmap_cuda   the same program with the five gap-fill solvers served by libdynprog_cuda through
                          gmap-gsnap_b200/host/dynprog_dropin.c: every worker thread runs DPC_FIBERS copies of the
                          worker loop, and the gaps stage 3 reaches are collected into device batches.

The database is built on the spot with the reference's own tools (oracle/_ref/bin: fa_coords, gmap_process,
gmapindex, driven like util/gmap_build.pl.in:100-215 does), with k-mer = base size 12 as the reference's
tests/setup1.test.in:12 does.  Host-side bookkeeping only; no alignment code lives here.
"""
import os
import subprocess
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
BIN = os.path.join(REFDIR, "bin")
_COMP = np.zeros(256, np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def have_binaries():
    need = [os.path.join(REFDIR, "gmap_ref"), os.path.join(REFDIR, "gmap_cuda")] + [os.path.join(BIN, b) for b in ("fa_coords", "gmap_process", "gmapindex")]
    return all(os.path.exists(p) for p in need)


def make_genome(total_bases, n_chr, seed):
    rng = np.random.default_rng(seed)
    per = total_bases // n_chr
    return [np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, per, dtype=np.uint8)].copy() for _ in range(n_chr)]


def make_transcripts(chroms, n, seed, length=2000, sub=0.01, indel=0.002):
    """4-10 exons, GT-AG introns of 50 bp - 20 kb (log-uniform), 1 % substitutions, 0.2 %
    indels, both strands.  The intron dinucleotides are planted in the genome (so call this BEFORE writing it)."""
    rng = np.random.default_rng(seed + 1)
    out = []
    sites = []          # (chromosome index, 1-based last exon base, "donor") / (.., last intron base, "acceptor")
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for t in range(n):
        c = int(rng.integers(0, len(chroms)))
        g = chroms[c]
        nex = int(rng.integers(4, 11))
        cuts = np.sort(rng.choice(np.arange(40, length - 40), nex - 1, replace=False))
        elens = np.diff(np.concatenate(([0], cuts, [length])))
        elens = np.maximum(elens, 30)
        ilens = np.exp(rng.uniform(np.log(50), np.log(20000), nex - 1)).astype(np.int64)
        span = int(elens.sum() + ilens.sum())
        pos = int(rng.integers(1000, len(g) - span - 1000))
        parts = []
        for e in range(nex):
            parts.append(g[pos:pos + elens[e]])
            pos += int(elens[e])
            if e < nex - 1:
                g[pos:pos + 2] = (71, 84)                              # GT
                g[pos + ilens[e] - 2:pos + ilens[e]] = (65, 71)         # AG
                sites.append((c, pos, "donor"))
                pos += int(ilens[e])
                sites.append((c, pos, "acceptor"))
        s = np.concatenate(parts)
        u = rng.random(len(s))
        subs = u < sub
        s = s.copy()
        s[subs] = acgt[(np.searchsorted(acgt, s[subs]) + rng.integers(1, 4, int(subs.sum()))) % 4]
        keep = ~((u >= sub) & (u < sub + indel / 2))                   # deletions
        ins = (u >= sub + indel / 2) & (u < sub + indel)               # insertions (one random base before)
        rep = np.where(ins, 2, 1) * keep
        s2 = np.repeat(s, rep)
        if ins.any():
            idx = np.cumsum(rep)[ins & keep] - 2
            s2[idx] = acgt[rng.integers(0, 4, len(idx))]
        if t & 1:
            s2 = _COMP[s2[::-1]]
        out.append(s2)
    make_transcripts.sites = sites
    return out


def write_fasta(path, names, seqs, width=60):
    with open(path, "wb") as f:
        for name, s in zip(names, seqs):
            f.write(b">" + name.encode() + b"\n")
            n = len(s) // width * width
            if n:
                body = np.empty((n // width, width + 1), np.uint8)
                body[:, :width] = s[:n].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n < len(s):
                f.write(s[n:].tobytes() + b"\n")


def build_db(workdir, genome_fa, dbname="synth", k=12):
    """fa_coords | gmap_process | gmapindex -A / -G / -O / -P, then the install step (util/gmap_build.pl.in:100-235)."""
    env = dict(os.environ)
    b = workdir
    def sh(cmd):
        r = subprocess.run(cmd, shell=True, cwd=workdir, env=env, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("%s\n%s" % (cmd, r.stderr[-2000:]))
    sh(f"{BIN}/fa_coords -o {b}/{dbname}.coords {genome_fa}")
    with open(os.path.join(b, dbname + ".version"), "w") as f:
        f.write(dbname + "\n")
    proc = f"{BIN}/gmap_process -c {b}/{dbname}.coords {genome_fa}"
    sh(f"{proc} | {BIN}/gmapindex -d {dbname} -D {b} -A")
    sh(f"{proc} | {BIN}/gmapindex -d {dbname} -F {b} -D {b} -G")
    sh(f"cat {b}/{dbname}.genomecomp | {BIN}/gmapindex -b {k} -k {k} -q 3 -d {dbname} -F {b} -D {b} -O")
    sh(f"cat {b}/{dbname}.genomecomp | {BIN}/gmapindex -b {k} -k {k} -q 3 -d {dbname} -F {b} -D {b} -P")
    dest = os.path.join(b, dbname)
    os.makedirs(os.path.join(dest, dbname + ".maps"), exist_ok=True)
    for fn in os.listdir(b):
        if fn.startswith(dbname + ".") and os.path.isfile(os.path.join(b, fn)) and not fn.endswith(".coords"):
            os.replace(os.path.join(b, fn), os.path.join(dest, fn))
    return b, dbname


def prepare(workdir, genome_bases, n_chr, n_transcripts, seed=5, known_sites_frac=0.0):
    os.makedirs(workdir, exist_ok=True)
    chroms = make_genome(genome_bases, n_chr, seed)
    tx = make_transcripts(chroms, n_transcripts, seed)
    gfa = os.path.join(workdir, "genome.fa")
    qfa = os.path.join(workdir, "transcripts.fa")
    write_fasta(gfa, ["chr%d" % (i + 1) for i in range(n_chr)], chroms)
    write_fasta(qfa, ["t%d" % i for i in range(n_transcripts)], tx)
    t0 = time.time()
    dbdir, dbname = build_db(workdir, gfa)
    case = {"dbdir": dbdir, "dbname": dbname, "queries": qfa, "n": n_transcripts, "bases": int(sum(len(s) for s in tx)),
            "db_build_s": time.time() - t0}
    if known_sites_frac > 0 and os.path.exists(os.path.join(BIN, "iit_store")):
        # a known-splice-site map (gmap -s): the format gtf_splicesites / gff3_splicesites print, stored with iit_store
        rng = np.random.default_rng(seed + 2)
        txt = os.path.join(workdir, "sites.txt")
        with open(txt, "w") as f:
            for i, (c, pos, kind) in enumerate(make_transcripts.sites):
                if rng.random() < known_sites_frac:
                    f.write(">site%d chr%d:%d..%d %s\n" % (i, c + 1, pos, pos + 1, kind))
        maps = os.path.join(dbdir, dbname, dbname + ".maps")
        r = subprocess.run(f"cat {txt} | {BIN}/iit_store -o {maps}/sites.iit", shell=True, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("iit_store: " + r.stderr[-1000:])
        # full path: this reference version dereferences a NULL user_splicingdir when the file is not found
        # "locally" first (gmap.c:3287)
        case["splicing"] = os.path.join(maps, "sites.iit")
    return case


def run_gmap(binary, case, threads, fibers=None, device=None, extra=(), out=None, timeout=3600):
    """Runs one binary over the case; returns (wall seconds of the whole process, stdout path, stderr text)."""
    env = dict(os.environ)
    if fibers is not None:
        env["DPC_FIBERS"] = str(fibers)
    if device is not None:
        env["DPC_DEVICE"] = str(device)
    env["DPC_FIBER_STATS"] = "1"
    out = out or os.path.join(case["dbdir"], os.path.basename(binary) + ".out")
    if case.get("splicing"):
        extra = ("-s", case["splicing"], *extra)
    cmd = [os.path.join(REFDIR, binary), "-D", case["dbdir"], "-d", case["dbname"], "-t", str(threads), "-O", *os.environ.get("GMAP_OUTFMT", "-A").split(), *extra, case["queries"]]
    import resource
    ru0 = resource.getrusage(resource.RUSAGE_CHILDREN)
    t0 = time.time()
    with open(out, "wb") as f:
        r = subprocess.run(cmd, stdout=f, stderr=subprocess.PIPE, env=env, timeout=timeout)
    dt = time.time() - t0
    ru1 = resource.getrusage(resource.RUSAGE_CHILDREN)
    run_gmap.last_cpu = (ru1.ru_utime - ru0.ru_utime, ru1.ru_stime - ru0.ru_stime, ru1.ru_minflt - ru0.ru_minflt)
    if r.returncode != 0:
        raise RuntimeError("%s exited %d\n%s" % (binary, r.returncode, r.stderr.decode()[-3000:]))
    return dt, out, r.stderr.decode()


def processed_seconds(err, default):
    """GMAP's own stopwatch ("Processed N queries in S seconds", gmap.c:3940): starts when the index is loaded and
    the workers are created, stops when the last result is out; excludes process start-up and tear-down of either arm."""
    for line in err.splitlines():
        if line.startswith("Processed ") and " seconds" in line:
            try:
                return float(line.split(" in ")[1].split(" seconds")[0])
            except (IndexError, ValueError):
                pass
    return default
