#!/usr/bin/env python
"""bench.py -- banded gap-fill DP throughput (GCUPS, gap-fills/s) of libdynprog_cuda on B200.

Headline workload (BASELINE.json configs[1]): batched Dynprog_single_gap, 1 M synthetic gap fills of 10-100 bp,
band (extraband_single) 30, widebandp, on ONE B200; random genome, queries = genomic windows with 5 %
substitutions, 3 % deletions, 3 % insertions.  Under torchrun each rank owns one GPU and its own
1 M problems (sharded by read, no collective on the data path): weak scaling.  The other two hot-path configs
(configs[2] genome gaps with their stated finalp / halfp / probability mix, configs[3] end gaps: the 10 M reads x 2
ends cut 8 ways = 2.5 M problems per GPU) run after the headline and are reported under `other_workloads`.

A step = one pass of the hot path over the batch.
  value   GCUPS with the batch resident in HBM (solve kernels only, CUDA events on the library's stream).
  e2e     GCUPS through the C ABI with HOST buffers: dpc_solve = copy-in, prepare / solve / finish / expand kernels,
          copy-out of results and Pair records (or host expansion of the compact device records), wall clock.
  --impl reference   the reference's own dynprog.c (oracle/_ref, compiled unmodified) on all host cores, same
          problems, same config object.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gmap_gsnap_b200 import api  # noqa: E402
from oracle import checkers  # noqa: E402   (the CPU reference: cpu_baseline, --impl reference and the parity sample only)

UNIT = "GCUPS"
GENOME_BASES = 64_000_000
EXTRABAND = 30
OPS_PER_CELL = 20          # SURVEY.md 8(d) / DESIGN.md section 3: integer operations of one 3-state cell update

WORKLOADS = {
    "single": "Dynprog_single_gap, 1M synthetic 10-100 bp gap fills, band 30, on 1 B200 per rank (BASELINE configs[1])",
    "genome": "Dynprog_genome_gap, 500K synthetic cDNA/genomic gaps across GT-AG/GC-AG/AT-AC introns 50 bp-20 kb, band 7, 10% long, "
              "finalp and halfp both ways, 10% re-solved in probability mode (BASELINE configs[2])",
    "end": "Dynprog_end5_gap/end3_gap, synthetic 250-bp read ends, tails 1-40 + 11 peeled, band 3; 10M reads x 2 ends cut 8 ways "
           "= 2.5M problems per GPU (BASELINE configs[3])",
    "gmap": "whole-program GMAP (stage 1-3) on synthetic 2-kb spliced transcripts vs a synthetic genome database, "
            "gap fills collected from stage 3 into device batches (BASELINE configs[4], bounded sample)",
}
DEFAULT_PROBLEMS = {"single": 1_000_000, "genome": 500_000, "end": 2_500_000}
METRICS = {"single": "banded_dp_gcups_single_gap", "genome": "banded_dp_gcups_genome_gap", "end": "banded_dp_gcups_end_gap"}
STANDIN_HOOK = api.PROB_FN(lambda which, pos, chroffset, user: ((pos * 2654435761 + which * 97) % 1000003) / 1000003.0)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, n):
    """DRAM bytes per step of the solve kernels from the committed ncu capture of the same workload and size
    (profiles/r2_traffic.json, written by tools/ncu_summary.py from the raw CSV next to it)."""
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    if not os.path.exists(path):
        return None, None
    t = json.load(open(path)).get(workload)
    if not t or int(t["problems"]) != int(n):
        return None, None
    return float(t["dram_bytes_per_step"]), t["source"]


def band_cells(L1, L2, eb):
    """In-band cells of matrices with widened bands."""
    L1, L2, eb = (np.asarray(x, dtype=np.int64) for x in (L1, L2, eb))
    rband = np.where(L2 >= L1, L2 - L1 + eb, eb)
    lband = np.where(L2 >= L1, eb, L1 - L2 + eb)
    total = np.zeros(len(L1), dtype=np.int64)
    for c in range(1, int(L2.max()) + 1 if len(L2) else 1):
        lo = np.maximum(1, c - rband)
        hi = np.minimum(L1, c + lband)
        total += np.where(c <= L2, np.maximum(hi - lo + 1, 0), 0)
    return total


def cells_numpy(probs, workload):
    """In-band cells of a batch in numpy (the reference arm has no device; the CUDA arm cross-checks the library's count)."""
    L1, L2, eb = probs["length1"], probs["length2"], probs["extraband"]
    if workload == "single":
        return int(band_cells(L1, L2, eb).sum())
    if workload == "genome":
        ok = L1 > 1
        return int(band_cells(L1[ok], L2[ok], eb[ok]).sum() + band_cells(L1[ok], probs["length2R"][ok], eb[ok]).sum())
    filled = (probs["endalign"] != api.QUERYEND_NOGAPS) & (L1 > 0) & (L2 > 0)      # NOGAPS ends fill no matrix
    return int(band_cells(L1[filled], L2[filled], eb[filled]).sum())


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_workload(rank, n, kind="single", genome_mix=False):
    """The synthetic inputs of BASELINE configs[1..3].  genome_mix: config 3's stated mix -- finalp and
    halfp both ways, a 10 % subset marked for the probability-mode second call (arm it with checkers.arm_probability_mode)."""
    w = api.Workload(GENOME_BASES, seed=0x9E3779B9 + rank, nchr=4)
    if kind == "genome":
        probs = w.genome_gaps(n, extraband=7, seed=0x5EED0003 + 1000 * rank, finalp_mode=2 if genome_mix else 0,
                              prob_mode_pm=100 if genome_mix else 0, long_frac=0.1, long_hi=600)
    elif kind == "end":
        probs = w.end_gaps(n, extraband=3, seed=0x5EED0004 + 1000 * rank)
    else:
        probs = w.single_gaps(n, extraband=EXTRABAND, seed=0x5EED0002 + 1000 * rank)
    return w, probs


def config_of(workload, n, extraband, cells):
    """The `config` object: identical for the CUDA arm and for --impl reference."""
    return {"workload": WORKLOADS[workload], "problems_per_gpu": int(n), "extraband": int(extraband), "genome_bases": GENOME_BASES,
            "cells_per_gpu": int(cells),
            "l2": "inputs larger than L2: every step streams all problem records, query bytes and result records through HBM"}


class _MtSolver:
    """solve() facade over RefOracle.solve_mt (results only), for arm_probability_mode."""

    def __init__(self, ref):
        self.ref = ref

    def solve(self, problems, want_pairs=False):
        res, _ = self.ref.solve_mt(problems, os.cpu_count() or 1)
        return res, None, None


def reference_for(w, workload):
    """The compiled reference registered on workload w; genome gaps get its own MaxEnt tables as the hook."""
    ref = checkers.RefOracle()
    ref.init()
    ref.setup(w.make_setup())
    if workload == "genome":
        ref.setup(w.make_setup(splice_prob=ref.splice_prob))
    return ref


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, all host threads (gmap -t N shape, gmap.c:2254-2276), on the
    same problems and with the same config object as the CUDA arm."""
    if rank != 0:
        return
    workload = args.workload
    n = args.problems or DEFAULT_PROBLEMS[workload]
    w, probs = make_workload(0, n, workload, genome_mix=workload == "genome")
    ref = reference_for(w, workload)
    if workload == "genome":
        probs = checkers.arm_probability_mode(probs, _MtSolver(ref))
    cores = os.cpu_count() or 1
    cells = cells_numpy(probs, workload)
    for _ in range(args.warmup):
        ref.solve_mt(probs[: max(1000, n // 8)], cores)
    secs = 0.0
    for _ in range(args.steps):
        _, s = ref.solve_mt(probs, cores)
        secs += s
    gcups = cells * args.steps / secs / 1e9
    line = {
        "impl": "reference", "metric": METRICS[workload], "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "fills_per_s": n * args.steps / secs,
        "config": config_of(workload, n, probs["extraband"][0], cells),
        "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "all %d problems of the step, unmodified dynprog.c -O3, one Dynprog_T triple + Pairpool per thread, "
                                   "%d threads (scratch allocation outside the timed region)" % (n, cores)},
        "e2e": {"value": gcups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_gmap_workload(args, rank, world, local_rank):
    """BASELINE configs[4] on a bounded sample: gmap_ref -t <cores> against gmap_cuda (drop-in solvers, fibers).
    One step = the whole transcript file through one binary; every rank aligns its own file against its own
    GPU with cores/world worker threads.  Queries/s; the outputs of the two binaries must be identical."""
    import filecmp
    from gmap_gsnap_b200 import gmap_e2e as g
    if not g.have_binaries():
        if rank == 0:
            print(json.dumps({"metric": "gmap_queries_per_s", "unavailable": "oracle/_ref binaries not built (oracle/build_gmap.sh)"}))
        return
    cores = os.cpu_count() or 1
    threads = max(1, cores // max(1, world))
    n = args.problems or 16000
    case = g.prepare("/tmp/dpc_gmap_case_%d" % rank, args.genome_bases, 4, n, seed=5 + rank)
    line = {"metric": "gmap_queries_per_s", "unit": "queries/s", "n_gpus": world, "steps": 1, "warmup": 0,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOADS["gmap"], "transcripts_per_rank": n, "transcript_bp": 2000,
                       "genome_bases": args.genome_bases, "worker_threads_per_rank": threads, "fibers_per_thread": args.fibers}}
    if args.impl == "reference":
        if rank != 0:
            return
        dt, out, err = g.run_gmap("gmap_ref", case, cores)
        dt = g.processed_seconds(err, dt)
        line.update({"impl": "reference", "value": n / dt, "ms_per_step": 1e3 * dt,
                     "cpu_baseline": {"value": n / dt, "unit": "queries/s", "cores": cores, "kind": "reference",
                                      "sample": "%d transcripts, unmodified gmap -t %d" % (n, cores)},
                     "e2e": {"value": n / dt, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line), flush=True)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    # two passes of each binary, alternating (rank 0 also runs the reference arm); the faster pass of each counts:
    # the host CPUs of these boxes are shared and single passes vary by +-10 %
    times, ref_times, walls, ref_walls, rout = [], [], [], [], None
    for rep in range(2):
        dt1, out, err = g.run_gmap("gmap_cuda", case, threads, fibers=args.fibers, device=local_rank)
        times.append(g.processed_seconds(err, dt1))
        walls.append(dt1)
        if rank == 0 and not args.no_cpu_baseline:
            rdt1, rout, rerr = g.run_gmap("gmap_ref", case, cores)
            ref_times.append(g.processed_seconds(rerr, rdt1))
            ref_walls.append(rdt1)
    dt = min(times)
    stats = [l for l in err.splitlines() if "device batches" in l]
    gaps = sum(int(l.split(" device batches, ")[1].split(" gaps")[0]) for l in stats)
    batches = sum(int(l.split(" fibers, ")[1].split(" device batches")[0]) for l in stats)
    total_dt, total_n, total_gaps = dt, float(n), float(gaps)
    if dist is not None:
        import torch
        t = torch.tensor([dt, 0, 0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([n, gaps], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        total_dt, total_n, total_gaps = float(t[0]), float(u[0]), float(u[1])
    if rank == 0:
        line.update({"value": total_n / total_dt, "ms_per_step": 1e3 * total_dt, "gap_fills_per_s": total_gaps / total_dt,
                     "gap_fills": int(total_gaps), "device_batches_rank0": batches, "gpu_launches": batches,
                     "e2e": {"value": total_n / total_dt, "unit": "queries/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                             "includes": "stage 1-3 on the host, gap fills on the device, output (GMAP's stopwatch: after the index load)"}})
        line["pass_seconds"] = {"gmap_cuda": times, "gmap_ref": ref_times, "gmap_cuda_process_wall": walls, "gmap_ref_process_wall": ref_walls,
                                "clock": "GMAP's own stopwatch (Processed N queries in S seconds); process wall beside it"}
        if not args.no_cpu_baseline:
            rdt = min(ref_times)
            line["cpu_baseline"] = {"value": n / rdt, "unit": "queries/s", "cores": cores, "kind": "reference",
                                    "sample": "the same %d transcripts of rank 0, unmodified gmap -t %d" % (n, cores)}
            line["outputs_identical_to_reference"] = bool(filecmp.cmp(rout, out, shallow=False))
            tdt, tout, terr = g.run_gmap("gmap_ref_timed", case, cores) if os.path.exists(os.path.join(g.REFDIR, "gmap_ref_timed")) else (0, None, "")
            for l in terr.splitlines():
                if l.startswith("dynprog_timed:"):
                    secs = float(l.split(" solver calls, ")[1].split(" thread-seconds")[0])
                    line["reference_share_of_time_in_the_five_solvers"] = secs / (tdt * cores)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure(args, lib, workload, n, rank, world, local_rank, host_threads, barrier, allreduce, sampler, have_ref):
    """One workload through both measurements (kernels only, end to end) plus the parity sample of this rank's shard.
    Returns the JSON object (rank 0 prints it; the other ranks contribute through the reductions)."""
    L = lib.lib
    mix = workload == "genome"
    w, probs = make_workload(rank, n, workload, genome_mix=mix)
    ref = reference_for(w, workload) if have_ref else None
    hook = (ref.splice_prob if ref is not None else STANDIN_HOOK) if mix else None     # MaxEnt hook = the reference's own tables
    lib.setup(w.make_setup(splice_prob=hook))
    if mix:
        probs = checkers.arm_probability_mode(probs, lib)       # second call of stage3.c:5833: threshold from a first pass
    lib.load(probs)                                             # leaves the batch resident in HBM on the context's own stream
    kernel_stats = lib.stats()
    cells = int(kernel_stats.cells)
    assert cells == cells_numpy(probs, workload), "cell count mismatch"

    def step():
        lib.check(L.dpc_relaunch(lib.ctx), "dpc_relaunch")
        lib.check(L.dpc_sync(lib.ctx), "dpc_sync")
        return lib.kernel_ms()

    for _ in range(args.warmup):
        step()
    barrier()
    if sampler is not None:
        sampler.start()
    ms = [step() for _ in range(args.steps)]
    barrier()
    if sampler is not None:
        # the clocks belong to the device-timed region; the sampler (an nvidia-smi process per sample) must not compete
        # with the host threads of the end-to-end calls below, which use every core
        sampler.stop_flag.set()
        sampler.join()
    dev_ms = allreduce(float(sum(ms)), "MAX")
    total_cells = allreduce(float(cells), "SUM")
    total_fills = allreduce(float(n), "SUM")
    launches = kernel_stats.launches * args.steps

    # end to end through the C ABI with host buffers (results + Pair records); the caller-owned arrays are page-locked
    # once (like a long-lived pair pool), so the copy engine reads the problems and writes results / records in place
    lib.check(L.dpc_reset(lib.ctx), "dpc_reset")
    res, pairs, off = lib.solve(probs)                 # also sizes the output arrays, reused below
    npairs = len(pairs)
    pairs = np.zeros(npairs + 1024, dtype=api.PAIR_DT)
    r2 = np.zeros(n, dtype=api.RESULT_DT)
    pinned = [pairs, r2, probs, w.last_qbuf]
    for a in pinned:
        lib.register(a)
    for _ in range(3):
        assert lib.solve_into(probs, r2, pairs, off) == npairs
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        lib.solve_into(probs, r2, pairs, off)
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    st = lib.stats()
    # the same call without Pair records (scores, end points, intron boundaries, counts only): what a caller pays
    # that compares candidates by score first
    lib.solve_into(probs, r2, None, off)
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        lib.solve_into(probs, r2, None, off)
    e2e_results_s = (time.perf_counter() - t0) / args.e2e_steps
    for a in pinned:
        lib.unregister(a)
    barrier()
    e2e_s = allreduce(e2e_s, "MAX")
    e2e_results_s = allreduce(e2e_results_s, "MAX")
    same = True
    for f in api.RESULT_FIELDS:
        same = same and bool((r2[f] == res[f]).all())
    assert same, "dpc_solve is not repeatable"

    # parity: every rank compares a strided sample of ITS OWN shard with the compiled reference (all fields, Pair records)
    parity_ok, sample_n, stride = 0.0, 0, max(1, n // 50_000)
    if have_ref:
        sample = probs[::stride]
        sample_n = len(sample)
        bad = api.compare(*ref.solve(sample), *lib.solve(sample), rtol=1e-6)
        if bad:
            raise AssertionError("rank %d: results differ from the compiled reference: %s" % (rank, bad[:3]))
        parity_ok = 1.0
    parity_ranks = int(allreduce(parity_ok, "SUM"))

    hbm_peak, peak_src = peaks()
    alu_gops, mix_gops = lib.int_peak(local_rank)
    step_ms = dev_ms / args.steps
    mean_ms = float(np.mean(ms))
    gcups = total_cells / (step_ms * 1e-3) / 1e9
    algo_bytes = float(kernel_stats.fill_bytes)
    traffic, traffic_src = ncu_traffic(workload, n)
    line = {
        "metric": METRICS[workload], "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "fills_per_s": total_fills / (step_ms * 1e-3),
        "config": config_of(workload, n, probs["extraband"][0], cells),
        "e2e": {"value": total_cells / e2e_s / 1e9, "unit": UNIT, "fills_per_s": total_fills / e2e_s, "ms_per_step": 1e3 * e2e_s,
                "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.d2h_bytes),
                "host_threads_per_rank": host_threads,
                "pipeline_chunks": int(st.pipeline_chunks), "host_half_chunks": int(st.host_chunks),
                "results_only": {"value": total_cells / e2e_results_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * e2e_results_s,
                                 "note": "dpc_solve with pairs == NULL: scores, end points, intron boundaries, counts"},
                "includes": "dpc_solve from and into page-locked caller arrays: copy-in of problem records + query bytes, prepare / solve / "
                            "finish / expand kernels, copy-out of results and %d Pair records (expanded by host threads from the compact "
                            "device records while they keep up, by the device otherwise)" % npairs},
        "gpu_launches": int(launches),
        "parity_checked_ranks": parity_ranks,
        "parity_sample": ("every rank: %d problems of its own shard (stride %d), all result fields and Pair records equal to the compiled "
                          "reference" % (sample_n, stride)) if have_ref else "compiled reference not on this box",
        "roofline": {"bound": "int32_alu", "achieved": cells * OPS_PER_CELL / (mean_ms * 1e-3) / 1e12, "peak": alu_gops / 1e3,
                     "unit": "T int32 lane-op/s", "frac": cells * OPS_PER_CELL / (mean_ms * 1e-3) / 1e9 / alu_gops,
                     "ops_per_cell": OPS_PER_CELL,
                     "peak_source": "measured in this process (dpc_measure_int_peak: independent VIMNMX+LOP3 chains = the integer ALU pipe "
                                    "the fill saturates, 64 lanes per clock per SM)",
                     "peak_fill_mix": mix_gops / 1e3,
                     "frac_of_fill_mix": cells * OPS_PER_CELL / (mean_ms * 1e-3) / 1e9 / mix_gops,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "note": "achieved = in-band cells x 20 integer ops (SURVEY.md 8d; derivation in DESIGN.md section 3) / mean kernel "
                             "time per step; peak_fill_mix is the same device running the fill's add/max/select mix, where adds may "
                             "issue on the FMA pipe"},
        "roofline_hbm": {"bound": "hbm", "achieved": algo_bytes / (mean_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": algo_bytes / (mean_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": algo_bytes, "traffic": traffic,
                         "note": "matrices stay in registers and direction planes in shared memory: algorithmic HBM bytes are "
                                 "descriptors, sequences, result records and staged genome characters only"},
    }
    if sampler is not None:
        line["clocks"] = sampler.summary()
    if rank == 0 and have_ref and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(n, max(20000, 40000 * cores))
        _, secs = ref.solve_mt(probs[:sample], cores)
        scells = cells_numpy(probs[:sample], workload)
        line["cpu_baseline"] = {"value": scells / secs / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                                "fills_per_s": sample / secs,
                                "sample": "first %d of the %d problems, unmodified reference dynprog.c (-O3), %d threads" % (sample, n, cores)}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--problems", type=int, default=0, help="problems per GPU (default: the config's size)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true", help="headline workload only")
    ap.add_argument("--workload", default="single", choices=sorted(WORKLOADS), help="single = the headline config; genome / end = the other hot-path configs")
    ap.add_argument("--genome-bases", type=int, default=100_000_000, help="gmap workload: size of the synthetic genome database")
    ap.add_argument("--fibers", type=int, default=16, help="gmap workload: worker loops per worker thread (DPC_FIBERS)")
    ap.add_argument("--kernel-only", action="store_true", help="profiling aid: load the batch, run the timed kernel steps, print their times, exit")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.workload == "gmap":
        run_gmap_workload(args, rank, world, local_rank)
        return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()

    def allreduce(x, op):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    n = args.problems or DEFAULT_PROBLEMS[args.workload]
    lib = api.CudaLib()
    lib.init()
    boot = api.Workload(1_000_000, seed=1, nchr=1)        # a context needs a registered genome; measure() registers its own
    lib.setup(boot.make_setup())
    lib.open(local_rank)
    # the ranks of one box share its host cores: split them (driver + worker threads of dpc_solve)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, (os.cpu_count() or 1) // max(1, local_world))
    lib.check(lib.lib.dpc_set_threads(lib.ctx, min(64, host_threads)), "dpc_set_threads")

    if args.kernel_only:
        w, probs = make_workload(rank, n, args.workload, genome_mix=False)
        lib.setup(w.make_setup())
        lib.load(probs)
        t = []
        for _ in range(args.warmup + args.steps):
            lib.check(lib.lib.dpc_relaunch(lib.ctx), "dpc_relaunch")
            lib.check(lib.lib.dpc_sync(lib.ctx), "dpc_sync")
            t.append(lib.kernel_ms())
        print(json.dumps({"kernel_only_ms": t, "problems": n, "cells": int(lib.stats().cells)}), flush=True)
        lib.close()
        return

    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdynprog_ref.so"))
    sampler = ClockSampler(local_rank)
    line = measure(args, lib, args.workload, n, rank, world, local_rank, host_threads, barrier, allreduce, sampler, have_ref)
    if args.workload == "single" and not args.no_other_workloads:
        # BASELINE configs[2] and configs[3] in the same record: fewer timed steps, same measurements
        short = argparse.Namespace(**vars(args))
        short.steps, short.e2e_steps = max(3, args.steps // 3), 3
        others = {}
        for wl in ("genome", "end"):
            o = measure(short, lib, wl, DEFAULT_PROBLEMS[wl], rank, world, local_rank, host_threads, barrier, allreduce, None, have_ref)
            keep = ("metric", "value", "unit", "ms_per_step", "steps", "fills_per_s", "config", "e2e", "gpu_launches", "parity_checked_ranks",
                    "roofline", "cpu_baseline")
            others[wl] = {k: o[k] for k in keep if k in o}
        line["other_workloads"] = others
    lib.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
