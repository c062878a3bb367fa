"""Host-resident batched solve: the call a user with NumPy (pinned) buffers makes.

`HostPipeline` streams a batch that lives in page-locked host memory through the device in chunks:
H2D copy of chunk c+1, the fwd + implicit-VJP kernels of chunk c and the D2H copy of chunk c-1 run on
three CUDA streams, so PCIe (full duplex) and the SMs overlap.  This is the path bench.py's `e2e`
number times (host->device and device->host copies inside the timed region).
"""
import numpy as np

from . import _lib as L
from . import lqr

GRAD_KEYS = lqr.LEAVES + ("x0",)


def pinned_like_problem(B, T, n, m, dtype, reduce_batch=False):
    """Pinned host buffers for the 11 problem leaves, the solution and the 11 gradient leaves (batch-summed gradients:
    leaf shapes without the batch axis, x0 stays per problem)."""
    shp = lqr.leaf_shapes(B, T, n, m)
    prob = {k: L.PinnedArray(s, dtype) for k, s in shp.items()}
    sol = {"X": L.PinnedArray((B, T, n), dtype), "U": L.PinnedArray((B, T, m), dtype), "Nu": L.PinnedArray((B, T, n), dtype)}
    gshp = {k: (s[1:] if (reduce_batch and k != "x0") else s) for k, s in shp.items()}
    grads = {k: L.PinnedArray(s, dtype) for k, s in gshp.items()}
    return prob, sol, grads


class HostPipeline:
    """chunk: problems per pipelined chunk.  ramp: the first chunk is split into pieces of chunk/8, chunk/4 and
    5 chunk/8 problems (each with its own device slot), so that the device->host stream -- the busier PCIe direction,
    gradients + solution against the problem going in -- starts after the transfer of chunk/8 problems instead of a
    whole chunk.  Measured on B200 / PCIe 5 x16 hosts: no gain (the pipeline already runs at 93-95 % of the duplex
    copy bandwidth of the host, tools/pcie_probe.py), hence off by default."""

    def __init__(self, B, T, n, m, dtype, chunk=8192, nslots=2, ramp=False, reduce_batch=False):
        self.B, self.T, self.n, self.m, self.dtype = B, T, n, m, np.dtype(dtype)
        # batch-summed gradients (the training use case): every chunk reduces on the device and sends ONE summed buffer
        # (the ten leaves, contiguous) to a pinned staging row; finish() adds the rows on the host (22 880 scalars per chunk
        # at n=8, m=4, T=100 against 228 T per PROBLEM in the per-problem mode)
        self.reduce_batch = bool(reduce_batch)
        self.stage = None
        self._pending = None
        self.chunk = min(chunk, B)
        if B % self.chunk:
            raise ValueError("batch must be a multiple of the chunk size")
        self.nslots = nslots
        cs = self.chunk
        head = [cs // 8, cs // 4, cs - cs // 8 - cs // 4] if (ramp and cs % 8 == 0 and B > cs) else [cs]
        self.sizes = head + [cs] * (B // cs - 1)
        self.slots = {}
        for size in sorted(set(self.sizes)):
            self.slots[size] = []
            for _ in range(nslots if size == cs else 1):
                p = lqr.DeviceProblem(size, T, n, m, dtype)
                plan = lqr.DevicePlan(size, T, n, m, dtype, vjp=True, reduce_batch=self.reduce_batch)
                self.slots[size].append([p, plan, L.Event(), L.Event(), L.Event(), False])
        if self.reduce_batch:
            any_plan = next(iter(self.slots.values()))[0][1]
            self.stage = L.PinnedArray((len(self.sizes),) + any_plan.grad_buffer.shape, dtype)
            self.leaf_views = {k: (int((any_plan.grads[k].ptr - any_plan.grad_buffer.ptr) // self.dtype.itemsize), any_plan.grads[k].shape)
                               for k in lqr.LEAVES}
        self.s_in, self.s_comp, self.s_out = L.Stream(), L.Stream(), L.Stream()
        self.ev0, self.ev1 = L.Event(), L.Event()
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def solve_and_vjp(self, prob, sol, grads, wait=True):
        """prob/sol/grads: dicts of PinnedArray.  Cotangent = (X, U, 0): the loss 0.5|X|^2 + 0.5|U|^2 of the
        reference's tests (tests/test_lqr.py:52-58).  Returns elapsed device milliseconds; with wait=False the call only
        enqueues (returns None): consecutive batches then overlap -- the copies of the next batch's first chunks run
        while the last chunks of this one compute and drain (order on the three streams keeps every slot safe)."""
        B = self.B
        self.h2d_bytes = self.d2h_bytes = 0
        self.ev0.record(self.s_in)
        b0 = 0
        self._chunk_index = 0
        uses = {size: 0 for size in self.slots}
        for nb in self.sizes:
            ring = self.slots[nb]
            slot = ring[uses[nb] % len(ring)]
            uses[nb] += 1
            p, plan, e_in, e_comp, e_out, used = slot
            if used:
                self.s_in.wait(e_out)  # slot's previous results have left the device
            slot[5] = True
            for k in GRAD_KEYS:
                src = prob[k]
                row = src.nbytes // B
                L.check(L.lib.idoc_memcpy_h2d(p.arr[k].ptr, src.ptr + b0 * row, nb * row, self.s_in.ptr), "h2d")
                self.h2d_bytes += nb * row
            e_in.record(self.s_in)
            self.s_comp.wait(e_in)
            plan.fwd(p, stream=self.s_comp)
            plan.vjp(p, plan.X, plan.U, None, stream=self.s_comp)
            e_comp.record(self.s_comp)
            self.s_out.wait(e_comp)
            for k, darr in (("X", plan.X), ("U", plan.U), ("Nu", plan.Nu)):
                row = sol[k].nbytes // B
                L.check(L.lib.idoc_memcpy_d2h(sol[k].ptr + b0 * row, darr.ptr, nb * row, self.s_out.ptr), "d2h")
                self.d2h_bytes += nb * row
            if self.reduce_batch:
                row = plan.grad_buffer.nbytes
                L.check(L.lib.idoc_memcpy_d2h(self.stage.ptr + self._chunk_index * row, plan.grad_buffer.ptr, row, self.s_out.ptr), "d2h")
                self.d2h_bytes += row
                rowx = grads["x0"].nbytes // B
                L.check(L.lib.idoc_memcpy_d2h(grads["x0"].ptr + b0 * rowx, plan.grads["x0"].ptr, nb * rowx, self.s_out.ptr), "d2h")
                self.d2h_bytes += nb * rowx
                self._chunk_index += 1
            else:
                for k in GRAD_KEYS:
                    row = grads[k].nbytes // B
                    L.check(L.lib.idoc_memcpy_d2h(grads[k].ptr + b0 * row, plan.grads[k].ptr, nb * row, self.s_out.ptr), "d2h")
                    self.d2h_bytes += nb * row
            e_out.record(self.s_out)
            b0 += nb
        self.ev1.record(self.s_out)
        self._pending = grads
        if not wait:
            return None
        ms = self.ev0.elapsed_ms(self.ev1)
        self.finish()
        return ms

    def finish(self):
        """Batch-summed mode: wait for the last enqueued batch and add its per-chunk sums into the host gradient leaves."""
        if not self.reduce_batch or self._pending is None:
            return
        self.ev1.sync()
        tot = self.stage.a.sum(axis=0)
        for k, (off, shape) in self.leaf_views.items():
            self._pending[k].a[...] = tot[off:off + int(np.prod(shape))].reshape(shape)
        self._pending = None


class NativeHostPipeline:
    """The same chunk pipeline behind ONE C-ABI call per batch (idoc_lqr_host_create / idoc_lqr_host_solve_vjp /
    idoc_lqr_host_finish, csrc/host_pipeline.cu): what a non-Python host binding (an XLA custom call on a CPU-resident
    buffer, cgo, JNI ...) uses.  prob / sol / grads are dicts of PinnedArray as for HostPipeline."""

    def __init__(self, B, T, n, m, dtype, chunk=2048, nslots=3, reduce_batch=False):
        import ctypes as C
        self.B, self.T, self.n, self.m, self.dtype = int(B), T, n, m, np.dtype(dtype)
        self.reduce_batch = bool(reduce_batch)
        h = C.c_void_p()
        L.check(L.lib.idoc_lqr_host_create(C.byref(h), L.dtype_code(dtype), min(int(chunk), self.B), T, n, m, nslots,
                                           int(self.reduce_batch)), "lqr_host_create")
        self.h = h
        s_in, s_out = C.c_void_p(), C.c_void_p()
        L.check(L.lib.idoc_lqr_host_streams(self.h, C.byref(s_in), C.byref(s_out)), "lqr_host_streams")

        class _S:  # stream handles for event timing across batches (not owned)
            def __init__(self, ptr):
                self.ptr = ptr
        self.s_in, self.s_out = _S(s_in.value), _S(s_out.value)
        self.h2d_bytes = self.d2h_bytes = 0

    def solve_and_vjp(self, prob, sol, grads, wait=True):
        import ctypes as C
        ms = C.c_float()
        L.check(L.lib.idoc_lqr_host_solve_vjp(self.h, self.B, *[prob[k].ptr for k in lqr.LEAVES], prob["x0"].ptr,
                                              sol["X"].ptr, sol["U"].ptr, sol["Nu"].ptr, *[grads[k].ptr for k in lqr.LEAVES],
                                              grads["x0"].ptr, int(bool(wait)), C.byref(ms)), "lqr_host_solve_vjp")
        a, b = C.c_uint64(), C.c_uint64()
        L.check(L.lib.idoc_lqr_host_bytes(self.h, C.byref(a), C.byref(b)), "lqr_host_bytes")
        self.h2d_bytes, self.d2h_bytes = int(a.value), int(b.value)
        return float(ms.value) if wait else None

    def finish(self):
        L.check(L.lib.idoc_lqr_host_finish(self.h), "lqr_host_finish")

    def close(self):
        if self.h is not None:
            L.lib.idoc_lqr_host_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
