"""Exploration aid (not part of the bench contract): gmap_ref vs gmap_cuda over thread / fiber counts."""
import os, sys, time, filecmp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmap_gsnap_b200 import gmap_e2e as g

bases = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
combos = sys.argv[3] if len(sys.argv) > 3 else "16x64,16x256,8x64"
cores = os.cpu_count()
case = g.prepare("/tmp/gmap_case", bases, 4, n)
print("case", case, flush=True)
for t in (cores // 2, cores):
    dt, ref_out, err = g.run_gmap("gmap_ref", case, t)
    print("gmap_ref -t %d: %.2f s  %s  user/sys/minflt %s" % (t, dt, err.strip().splitlines()[-1], g.run_gmap.last_cpu), flush=True)
for c in combos.split(","):
    t, f = (int(x) for x in c.split(":")[0].split("x"))
    for k in ("MALLOC_MMAP_MAX_", "MALLOC_TRIM_THRESHOLD_", "MALLOC_TOP_PAD_", "DPC_SYNC", "DPC_MEMO", "DPC_MALLOC_TUNING"):
        os.environ.pop(k, None)
    for opt in c.split(":")[1:]:
        if opt == "memo0":
            os.environ["DPC_MEMO"] = "0"
        elif opt == "notune":
            os.environ["DPC_MALLOC_TUNING"] = "0"
        elif opt == "nommap":
            os.environ.update(MALLOC_MMAP_MAX_="0", MALLOC_TRIM_THRESHOLD_="17179869184", MALLOC_TOP_PAD_="268435456")
        else:
            os.environ["DPC_SYNC"] = opt
    dt, out, err = g.run_gmap("gmap_cuda", case, t, fibers=f, out="/tmp/gmap_case/cuda_%s.out" % c)
    same = filecmp.cmp(ref_out, out, shallow=False)
    lines = err.strip().splitlines()
    stats = [l for l in lines if "device batches" in l]
    print("gmap_cuda " + c + " -t %d fibers %d: %.2f s  identical=%s  %s | %s" % (t, f, dt, same, lines[-1], str(g.run_gmap.last_cpu) + " " + " ## ".join(stats[:1])), flush=True)
