"""Multi-GPU plumbing of the hot path.

Problems are independent, so the batch is split evenly and every rank (one process per GPU) solves its own
shard with no data-path communication.  The one exchange step exists only when the caller wants gradients
summed over the batch (`DevicePlan(reduce_batch=True)`): each rank reduces its shard on the device and a single
NCCL all-reduce over NVLink combines the per-rank buffers (`GradAllReducer`).  Rendezvous (distributing the
128-byte ncclUniqueId) is left to the host program; `torch.distributed` (any backend) or a shared file work.
"""
import ctypes as C
import os
import time

import numpy as np

from . import _lib as L


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous even split of the batch axis: (start, count) of `rank`'s shard (first `batch % world` ranks
    get one extra problem)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_params(params, rank: int, world: int):
    """Slice every leaf of a (batched) pytree of NumPy arrays along the leading axis."""
    def rec(t):
        if isinstance(t, tuple):
            cols = [rec(c) for c in t]
            return type(t)(*cols) if hasattr(t, "_fields") else tuple(cols)
        a = np.asarray(t)
        s, c = shard_bounds(a.shape[0], rank, world)
        return a[s:s + c]
    return rec(params)


def exchange_unique_id(rank: int, world: int, path: str = None):
    """Create the ncclUniqueId on rank 0 and hand it to every rank: through torch.distributed when a process
    group is initialised, else through the file `path` (shared file system)."""
    buf = (C.c_char * 128)()
    try:
        import torch.distributed as dist
        have_pg = dist.is_available() and dist.is_initialized()
    except Exception:
        have_pg = False
    if have_pg:
        obj = [None]
        if rank == 0:
            L.check(L.lib.idoc_nccl_unique_id(buf), "nccl_unique_id")
            obj = [bytes(buf)]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    if path is None:
        raise L.IdocError("no torch.distributed process group and no rendezvous file given")
    # File rendezvous.  A stale file from an earlier run must not be mistaken for this run's id: the file name carries the
    # launcher's run id / port when there is one, rank 0 removes any old file before publishing (atomically, tmp + rename),
    # and the other ranks only accept a 128-byte file written within the last `max_age` seconds.
    run = os.environ.get("TORCHELASTIC_RUN_ID") or os.environ.get("MASTER_PORT")
    if run:
        path = f"{path}.{run}"
    max_age = 120.0
    if rank == 0:
        try:
            os.unlink(path)
        except FileNotFoundError:
            pass
        L.check(L.lib.idoc_nccl_unique_id(buf), "nccl_unique_id")
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(bytes(buf))
        os.replace(tmp, path)
        return bytes(buf)
    for _ in range(600):
        try:
            st = os.stat(path)
            if st.st_size == 128 and time.time() - st.st_mtime < max_age:
                return open(path, "rb").read()
        except FileNotFoundError:
            pass
        time.sleep(0.1)
    raise L.IdocError("timed out waiting for the ncclUniqueId file")


class GradAllReducer:
    """NCCL communicator over the ranks of one box + in-place sum all-reduce of DeviceArrays."""

    def __init__(self, rank: int, world: int, unique_id: bytes):
        self.rank, self.world = rank, world
        comm = C.c_void_p()
        idbuf = (C.c_char * 128).from_buffer_copy(unique_id)
        L.check(L.lib.idoc_nccl_init(idbuf, rank, world, C.byref(comm)), "nccl_init")
        self.comm = comm

    def allreduce_(self, arrays, stream=None):
        """arrays: iterable of DeviceArray (summed in place across ranks, on `stream`)."""
        for a in arrays:
            count = int(np.prod(a.shape, dtype=np.int64))
            L.check(L.lib.idoc_nccl_allreduce_sum(self.comm, stream.ptr if stream else None, L.dtype_code(a.dtype),
                                                  a.ptr, count), "nccl_allreduce")

    def allreduce_plan_grads(self, plan, stream=None):
        """All-reduce the batch-reduced gradient leaves of a DevicePlan(reduce_batch=True) (x0 stays per problem)."""
        if not plan.reduce_batch:
            raise L.IdocError("plan was not created with reduce_batch=True")
        self.allreduce_([plan.grad_buffer], stream)  # ONE ncclAllReduce over the contiguous buffer of all ten leaves

    def close(self):
        if self.comm is not None:
            L.lib.idoc_nccl_destroy(self.comm)
            self.comm = None


def allreduce_sum_host(arrays):
    """Sum host (NumPy) arrays across the ranks of an initialised torch.distributed process group (any backend:
    gloo on CPU).  Used for host-resident reductions and by the CPU tests of the sharding logic."""
    import torch
    import torch.distributed as dist
    out = []
    for a in arrays:
        t = torch.from_numpy(np.ascontiguousarray(a).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        out.append(t.numpy())
    return out


def verify_allreduced_gradients(reducer: "GradAllReducer", n: int, m: int, T: int, batch_per_rank: int, dtype, seed: int = 77,
                                stream=None):
    """Check the path's one exchange step end to end on the GPUs of an initialised NCCL communicator.

    Every rank generates its own `batch_per_rank` problems, runs fwd + implicit VJP twice — once with per-problem gradients
    (reduce_batch = 0), once batch-summed (vectors-only sweeps + the batch contraction of csrc/grad_reduce.cuh) followed by
    ONE ncclAllReduce over the contiguous gradient buffer — and compares  sum over ranks and problems of the per-problem
    gradients  (summed on the host in float64, combined across ranks with torch.distributed / gloo) with the all-reduced
    device buffer.  Returns the worst relative error over the ten leaves (relative to the largest summand), the same on
    every rank.  Needs an initialised torch.distributed process group when world > 1."""
    from . import lqr as _lqr
    rank, world = reducer.rank, reducer.world
    B = int(batch_per_rank)
    prob = _lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=seed + rank, stream=stream)
    per = _lqr.DevicePlan(B, T, n, m, dtype, vjp=True, reduce_batch=False)
    red = _lqr.DevicePlan(B, T, n, m, dtype, vjp=True, reduce_batch=True)
    per.fwd(prob, stream=stream)
    per.vjp(prob, per.X, per.U, None, stream=stream)
    red.fwd(prob, stream=stream)
    red.vjp(prob, red.X, red.U, None, stream=stream)
    reducer.allreduce_plan_grads(red, stream=stream)
    if stream is not None:
        stream.sync()
    else:
        L.sync()
    want = {k: per.grads[k].numpy().astype(np.float64).sum(axis=0) for k in _lqr.LEAVES}
    scale = {k: float(np.abs(per.grads[k].numpy()).max()) for k in _lqr.LEAVES}
    if world > 1:
        import torch
        import torch.distributed as dist
        flat = torch.from_numpy(np.concatenate([want[k].ravel() for k in _lqr.LEAVES]))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        sc = torch.tensor([scale[k] for k in _lqr.LEAVES], dtype=torch.float64)
        dist.all_reduce(sc, op=dist.ReduceOp.MAX)
        o = 0
        for i, k in enumerate(_lqr.LEAVES):
            cnt = want[k].size
            want[k] = flat[o:o + cnt].numpy().reshape(want[k].shape)
            scale[k] = float(sc[i])
            o += cnt
    worst = 0.0
    for k in _lqr.LEAVES:
        got = red.grads[k].numpy().astype(np.float64)
        worst = max(worst, float(np.abs(got - want[k]).max() / max(scale[k], 1e-300)))
    return worst
