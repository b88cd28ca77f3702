#!/usr/bin/env python
"""bench.py -- banded gap-fill DP throughput (GCUPS, gap-fills/s) of libdynprog_cuda on B200.

Workload (BASELINE.json configs[1]): batched Dynprog_single_gap, 1 M synthetic gap fills of 10-100 bp,
band (extraband_single) 30, widebandp, on ONE B200; random genome, queries = genomic windows with 5 %
substitutions, 3 % deletions, 3 % insertions.  Under torchrun each rank owns one GPU and its own
1 M problems (sharded by read, no collective on the data path): weak scaling.

A step = one pass of the hot path over the batch.
  value   GCUPS with the batch resident in HBM (kernels only, CUDA events on the library's stream).
  e2e     GCUPS through the C ABI with HOST buffers: dpc_solve = pack + H2D + kernels + D2H + result
          finalisation + Pair-record rebuild, wall clock.
  --impl reference   the reference's own dynprog.c (oracle/_ref, compiled unmodified) on all host cores.
"""
import argparse
import ctypes as C
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
from oracle import checkers  # noqa: E402   (the CPU reference: cpu_baseline and --impl reference only)

METRIC = "banded_dp_gcups_single_gap"
UNIT = "GCUPS"
N_PROBLEMS = 1_000_000
GENOME_BASES = 64_000_000
EXTRABAND = 30


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def int32_alu_gops():
    """Measured throughput of the integer ALU pipe (VIMNMX / LOP3), giga lane-operations per second."""
    path = os.path.join(ROOT, "profiles", "int32_peak_r1.json")
    if os.path.exists(path):
        return float(json.load(open(path))["max_xor_gops"])
    return 148 * 64 * 1.965          # the same figure from the data sheet: 64 lanes per clock per SM


def band_cells(probs):
    """In-band cells of every matrix, for single gaps: one matrix each."""
    L1 = probs["length1"].astype(np.int64)
    L2 = probs["length2"].astype(np.int64)
    eb = probs["extraband"].astype(np.int64)
    rband = np.where(L2 >= L1, L2 - L1 + eb, eb)
    lband = np.where(L2 >= L1, eb, L1 - L2 + eb)
    total = np.zeros(len(probs), dtype=np.int64)
    for c in range(1, int(L2.max()) + 1):
        lo = np.maximum(1, c - rband)
        hi = np.minimum(L1, c + lband)
        total += np.where(c <= L2, np.maximum(hi - lo + 1, 0), 0)
    return total


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


WORKLOADS = {
    "single": "Dynprog_single_gap, 1M synthetic 10-100 bp gap fills, band 30, on 1 B200 per rank (BASELINE configs[1])",
    "genome": "Dynprog_genome_gap, synthetic cDNA/genomic gaps across GT-AG/GC-AG/AT-AC introns 50 bp-20 kb, band 7, 10% long (BASELINE configs[2], finalp off)",
    "end": "Dynprog_end5_gap/end3_gap, synthetic 250-bp read ends, tails 1-40 + 11 peeled, band 3 (BASELINE configs[3])",
    "gmap": "whole-program GMAP (stage 1-3) on synthetic 2-kb spliced transcripts vs a synthetic genome database, "
            "gap fills collected from stage 3 into device batches (BASELINE configs[4], bounded sample)",
}


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


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path, all host threads (gmap -t N shape, gmap.c:2254-2276)."""
    if rank != 0:
        return
    ref = checkers.RefOracle()
    ref.init()
    cores = os.cpu_count() or 1
    sample = min(N_PROBLEMS, max(20000, 40000 * cores))
    w, probs = make_workload(0, sample)
    ref.setup(w.make_setup())
    cells = int(band_cells(probs).sum())
    for _ in range(args.warmup):
        ref.solve_mt(probs[: sample // 4], cores)
    secs = 0.0
    for _ in range(args.steps):
        _, s = ref.solve_mt(probs, cores)
        secs += s
    gcups = cells * args.steps / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "fills_per_s": sample * args.steps / secs,
        "config": {"workload": "Dynprog_single_gap, synthetic 10-100 bp gap fills, band 30 (BASELINE configs[1])",
                   "problems_per_step": sample, "extraband_single": EXTRABAND, "genome_bases": GENOME_BASES},
        "cpu_baseline": {"value": gcups, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d of the 1M single-gap problems per step, unmodified dynprog.c -O3, one Dynprog_T + Pairpool per thread" % sample},
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
    n = args.problems if args.problems != N_PROBLEMS else 16000
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--problems", type=int, default=N_PROBLEMS)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    w, probs = make_workload(rank, args.problems, args.workload)
    lib = api.CudaLib()
    lib.init()
    lib.setup(w.make_setup())
    lib.open(local_rank)
    L = lib.lib
    n = len(probs)
    # the ranks of one box share its host cores: split them (packing / Pair rebuild threads of dpc_solve)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_threads = max(1, (os.cpu_count() or 1) // max(1, local_world))
    lib.check(L.dpc_set_threads(lib.ctx, min(64, host_threads)), "dpc_set_threads")

    # results for the sanity checks (bulk call), then the whole batch resident in HBM on the context's own stream
    if args.kernel_only:
        lib.load(probs)
        t = []
        for _ in range(args.warmup + args.steps):
            lib.check(L.dpc_relaunch(lib.ctx), "dpc_relaunch")
            lib.check(L.dpc_sync(lib.ctx), "dpc_sync")
            t.append(lib.kernel_ms())
        print(json.dumps({"kernel_only_ms": t, "problems": n, "cells": int(lib.stats().cells)}), flush=True)
        lib.close()
        return
    res, _, _ = lib.solve(probs, want_pairs=False)
    lib.load(probs)
    stats = lib.stats()
    cells = int(stats.cells)
    if args.workload == "single":
        assert cells == int(band_cells(probs).sum()), "cell count mismatch"
        assert (res["null_list"] == 0).all()

    def step():
        lib.check(L.dpc_relaunch(lib.ctx), "dpc_relaunch")
        lib.check(L.dpc_sync(lib.ctx), "dpc_sync")
        return lib.kernel_ms()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ms = [step() for _ in range(args.steps)]
    barrier()
    dev_ms = allreduce(float(sum(ms)), "MAX")
    total_cells = allreduce(float(cells), "SUM")
    total_fills = allreduce(float(n), "SUM")
    launches = stats.launches * args.steps

    # end to end through the C ABI with host buffers (results + Pair records)
    kernel_stats = lib.stats()
    lib.check(L.dpc_reset(lib.ctx), "dpc_reset")
    r2, pairs, off = lib.solve(probs)             # also sizes the caller-owned output arrays, reused below
    npairs = len(pairs)
    pairs = np.zeros(npairs + 1024, dtype=api.PAIR_DT)
    # caller-owned output arrays, page-locked once (like a long-lived pair pool): the copy engine writes into them
    pinned = [pairs, r2, probs, w.last_qbuf]
    for a in pinned:
        lib.register(a)
    for _ in range(2):
        assert lib.solve_into(probs, r2, pairs, off) == npairs
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        lib.solve_into(probs, r2, pairs, off)
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    # the same call without Pair records (scores, end points, intron boundaries, counts only): what a caller pays
    # that compares candidates by score first, and what the Pair rebuild costs on the host
    lib.solve_into(probs, r2, None, off)
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        lib.solve_into(probs, r2, None, off)
    e2e_results_s = (time.perf_counter() - t0) / args.e2e_steps
    for a in pinned:
        lib.unregister(a)
    pairs = pairs[:npairs]
    barrier()
    sampler.stop_flag.set()
    sampler.join()
    e2e_s = allreduce(e2e_s, "MAX")
    e2e_results_s = allreduce(e2e_results_s, "MAX")
    st = lib.stats()
    assert (r2 == res).all()

    hbm_peak, peak_src, sm_max = peaks()
    step_ms = dev_ms / args.steps
    gcups = total_cells / (step_ms * 1e-3) / 1e9
    algo_bytes = float(kernel_stats.fill_bytes)
    line = {
        "metric": METRIC, "value": gcups, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "fills_per_s": total_fills / (step_ms * 1e-3),
        "config": {"workload": WORKLOADS[args.workload],
                   "problems_per_gpu": n, "extraband": int(probs["extraband"][0]), "genome_bases": GENOME_BASES,
                   "cells_per_gpu": cells, "l2": "inputs larger than L2 (descriptors + sequences + results = %.0f MB per step)" % (algo_bytes / 1e6)},
        "e2e": {"value": total_cells / e2e_s / 1e9, "unit": UNIT, "fills_per_s": total_fills / e2e_s, "ms_per_step": 1e3 * e2e_s,
                "h2d_bytes_per_step": int(st.h2d_bytes), "d2h_bytes_per_step": int(st.d2h_bytes),
                "host_threads_per_rank": host_threads,
                "results_only": {"value": total_cells / e2e_results_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * e2e_results_s,
                                 "note": "dpc_solve with pairs == NULL: everything but the Pair-record rebuild"},
                "includes": "dpc_solve: pack, H2D, kernels, D2H, result finalisation, Pair-record rebuild (%d records)" % len(pairs)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"bound": "hbm", "achieved": algo_bytes / (ms[len(ms) // 2] * 1e-3) / 1e9 if rank == 0 else None,
                     "peak": hbm_peak, "unit": "GB/s", "frac": algo_bytes / (float(np.mean(ms)) * 1e-3) / 1e9 / hbm_peak,
                     "traffic": 620.4e6 if (args.workload == "single" and n == 1_000_000) else None,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the step's two launches, "
                                       "profiles/r1_final_single_gap_ncu_full.csv (bytes per step)",
                     "peak_source": peak_src,
                     "note": "fused fill+traceback keeps matrices and direction nibbles in shared memory; algorithmic HBM bytes are "
                             "descriptors, sequences and result records only, so the kernel is integer-ALU/latency bound (see alu_roofline)"},
        "alu_roofline": {"achieved_gcups_per_gpu": cells / (float(np.mean(ms)) * 1e-3) / 1e9,
                         "peak_gcups": int32_alu_gops() / 25,
                         "note": "peak = measured integer-ALU-pipe throughput (tools/int_peak.cu on this pool's B200: 18.5 T "
                                 "VIMNMX+LOP3 lane-ops/s = 63.7 per clock per SM, profiles/int32_peak_r1.json) / 25 ALU ops per cell (DESIGN.md)"},
    }
    line["roofline"]["achieved"] = algo_bytes / (float(np.mean(ms)) * 1e-3) / 1e9
    line["alu_roofline"]["frac"] = line["alu_roofline"]["achieved_gcups_per_gpu"] / line["alu_roofline"]["peak_gcups"]

    if args.workload != "single":
        line["metric"] = "banded_dp_gcups_%s_gap" % args.workload
    if rank == 0 and not args.no_cpu_baseline and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libdynprog_ref.so")):
        ref = checkers.RefOracle()
        ref.init()
        ref.setup(w.make_setup())
        cores = os.cpu_count() or 1
        sample = min(n, max(20000, 40000 * cores))
        _, secs = ref.solve_mt(probs[:sample], cores)
        scells = int(band_cells(probs[:sample]).sum()) if args.workload == "single" else int(round(cells * sample / n))
        line["cpu_baseline"] = {"value": scells / secs / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                                "fills_per_s": sample / secs,
                                "sample": "first %d of the %d problems, unmodified reference dynprog.c (-O3), %d threads" % (sample, n, cores)}
    lib.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
