"""N > 1 path on CPU: two processes (gloo), each solving its shard, host-side gather on rank 0 == one-process
result.  The solver here is the CPU emulation of the device routines (tests/emul); on the GPU box bench.py runs the
same sharding with one CudaLib context per rank."""
import os
import subprocess
import sys

import numpy as np

from gmap_gsnap_b200 import api, shard
from oracle import checkers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import torch.distributed as dist
from gmap_gsnap_b200 import api, shard
from oracle import checkers
from util import mixed_problems
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w = api.Workload(1_000_000, seed=5, nchr=2)
probs = mixed_problems(w, 150, 33)            # same seeds on every rank: the same problem array everywhere
em = checkers.EmulLib(); em.init()
em.setup(w.make_setup(splice_prob=api.PROB_FN(lambda which, pos, chroffset, user: ((pos * 2654435761 + which * 97) %% 1000003) / 1000003.0)))
out = shard.solve_sharded(em, probs, rank, world, dist)
if rank == 0:
    np.savez(%(out)r, res=out[0], pairs=out[1], off=out[2])
dist.barrier()
dist.destroy_process_group()
'''


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo_equal_one_process(tmp_path):
    out = str(tmp_path / "gathered.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "out": out})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)], env=env, timeout=600)
    z = np.load(out)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import mixed_problems
    w = api.Workload(1_000_000, seed=5, nchr=2)
    probs = mixed_problems(w, 150, 33)
    em = checkers.EmulLib()
    em.init()
    em.setup(w.make_setup(splice_prob=api.PROB_FN(lambda which, pos, chroffset, user: ((pos * 2654435761 + which * 97) % 1000003) / 1000003.0)))
    want = em.solve(probs)
    assert not api.compare(*want, z["res"], z["pairs"], z["off"])
