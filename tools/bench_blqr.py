#!/usr/bin/env python
"""Timing of the shared-control batch LQR (idoc/blqr.py) on the device: the block-diagonal + low-rank solver
(idoc_blqr_fwd + idoc_blqr_vjp, csrc/blqr.cuh) over the number of systems, next to the dense lift on the LQR kernels
(n' = batch_size * n <= 128: what round 1 shipped).  One JSON line per size."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib


def problem(rng, T, bs, n, m):
    spd = lambda shape, k: (lambda W: W @ np.swapaxes(W, -1, -2) / k + np.eye(k))(rng.standard_normal(shape + (k, k)))
    A = 0.5 * np.linalg.qr(rng.standard_normal((T, bs, n, n)))[0]
    l = ib.blqr.BLQR(Q=spd((T, bs), n), q=0.3 * rng.standard_normal((T, bs, n)), Qf=spd((bs,), n), qf=0.3 * rng.standard_normal((bs, n)),
                     M=0.05 * rng.standard_normal((T, bs, n, m)), R=bs * spd((T,), m), r=rng.standard_normal((T, m)), A=A,
                     B=rng.standard_normal((T, bs, n, m)), d=0.3 * rng.standard_normal((T, bs, n)))
    return ib.blqr.Params(rng.standard_normal((bs, n)), l)


def main():
    rng = np.random.default_rng(0)
    n, m, T = 3, 2, 100
    for bs in (30, 256, 1024, 4096):
        p = problem(rng, T, bs, n, m)
        dev = ib.blqr._Dev(p, True)
        ct = ib.typs.State(np.ones((T, bs, n)), np.ones((T, m)), None)
        dev.fwd()
        dev.vjp(ct)
        L.sync()
        e0, e1 = L.Event(), L.Event()
        reps = 3
        e0.record()
        for _ in range(reps):
            a = dev.arr
            L.check(L.lib.idoc_blqr_fwd(None, dev.code, 1, T, bs, n, m, *[a[k].ptr for k in ib.blqr.FIELDS], dev.x0.ptr, dev.X.ptr,
                                        dev.U.ptr, dev.Nu.ptr, dev.K.ptr, dev.k.ptr, dev.dCdc.ptr, dev.status.ptr, dev.ws.ptr,
                                        dev.ws_bytes), "blqr_fwd")
        e1.record()
        e1.sync()
        fwd_ms = e0.elapsed_ms(e1) / reps
        t0 = time.perf_counter()
        dev.vjp(ct)
        L.sync()
        vjp_wall_ms = (time.perf_counter() - t0) * 1e3
        rec = dict(batch_size=bs, n=n, m=m, T=T, lifted_state=bs * n, fwd_ms=fwd_ms, vjp_wall_ms_incl_copies=vjp_wall_ms)
        if bs * n <= 128:  # the dense lift on the LQR kernels (round 1)
            lp = ib.blqr.lift(p)
            dp = ib.lqr.DeviceProblem.from_params(lp)
            plan = ib.lqr.DevicePlan(1, T, bs * n, m, np.float64, vjp=False)
            plan.fwd(dp)
            L.sync()
            e0.record()
            for _ in range(reps):
                plan.fwd(dp)
            e1.record()
            e1.sync()
            rec["dense_lift_fwd_ms"] = e0.elapsed_ms(e1) / reps
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
