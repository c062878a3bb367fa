"""CPU tests of the N>1 path (gloo, world_size 2): batch sharding and the all-reduce of batch-reduced
gradients.  The compute on each rank is the oracle (no GPU here); the exchange is the same call pattern the
GPU path makes with NCCL (idoc_b200/dist.py)."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from oracle import lqr_np as O
from idoc_b200 import dist as D
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(123)                     # same seed on every rank: the full batch
B, n, m, T = 7, 4, 2, 6                              # ragged split: 4 + 3
p = O.random_params(rng, n, m, T, batch=(B,))
so = O.direct(p)
w = O.State(*[rng.standard_normal(a.shape) for a in so])
full = O.vjp_exact(p, so, w)
start, count = D.shard_bounds(B, rank, world)
ps = D.shard_params(p, rank, world)
ws = D.shard_params(w, rank, world)
assert ps.x0.shape[0] == count
ss = O.direct(ps)
np.testing.assert_allclose(ss.X, so.X[start:start + count], rtol=0, atol=1e-13)
g = O.vjp_exact(ps, ss, ws)
local = [a.sum(axis=0) for a in g.lqr]              # what idoc_lqr_vjp(reduce_batch=1) leaves on this rank
summed = D.allreduce_sum_host(local)
for got, want in zip(summed, full.lqr):
    np.testing.assert_allclose(got, want.sum(axis=0), rtol=1e-11, atol=1e-12)
# every problem is owned by exactly one rank
owned = D.allreduce_sum_host([np.bincount(np.arange(start, start + count), minlength=B).astype(np.float64)])[0]
assert (owned == 1).all()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds():
    from idoc_b200 import dist as D
    for B in (1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            tot, nxt = 0, 0
            for r in range(world):
                s, c = D.shard_bounds(B, r, world)
                assert s == nxt and c >= 0
                nxt = s + c
                tot += c
            assert tot == B


def test_two_rank_reduced_gradients(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
