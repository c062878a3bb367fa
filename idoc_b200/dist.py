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
    if rank == 0:
        L.check(L.lib.idoc_nccl_unique_id(buf), "nccl_unique_id")
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(bytes(buf))
        os.replace(tmp, path)
        return bytes(buf)
    for _ in range(600):
        if os.path.exists(path) and os.path.getsize(path) == 128:
            return open(path, "rb").read()
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
        self.allreduce_([plan.grads[k] for k in plan.grads if k != "x0"], stream)

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
