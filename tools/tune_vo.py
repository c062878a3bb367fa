#!/usr/bin/env python
"""Sweep of the adjoint-forward sweep's variants in vectors-only mode (batch-summed gradients; library built with
IDOC_BUILD_TUNABLE=1): per (NBUF, BM) the mean time of lqr_adj_fwd and grad_reduce at BASELINE config 2.

    python tools/tune_vo.py [--dtype f64] [--batch 65536]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib
VARIANTS = [(2, 228), (2, 224), (2, 100), (2, 99), (2, 97), (2, 96), (2, 98), (1, 100), (1, 228), (2, 353), (2, 352), (2, 356), (2, 225), (2, 226)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dtype = np.float64 if args.dtype == "f64" else np.float32
    B, T, n, m = args.batch, 100, 8, 4
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=1)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, reduce_batch=True)
    plan.fwd(prob)
    ref = None
    for nb, bm in [(None, None)] + VARIANTS:
        if nb is None:
            os.environ.pop("IDOC_NBUF_F", None)
            os.environ.pop("IDOC_BM_AFWD", None)
        else:
            os.environ["IDOC_NBUF_F"], os.environ["IDOC_BM_AFWD"] = str(nb), str(bm)
        plan.vjp(prob, plan.X, plan.U, None)
        L.sync()
        L.profile_begin()
        for _ in range(args.reps):
            plan.vjp(prob, plan.X, plan.U, None)
        L.sync()
        k = L.profile_end()
        g = plan.grad_buffer.numpy()
        if ref is None:
            ref = g
        print(json.dumps(dict(nbuf=nb, bm=bm, dtype=args.dtype, adj_fwd_ms=float(np.mean(k["lqr_adj_fwd"])),
                              adj_bwd_ms=float(np.mean(k["lqr_adj_bwd"])), grad_reduce_ms=float(np.sum(k["grad_reduce"]) / args.reps),
                              same=bool(np.array_equal(g, ref)))), flush=True)


if __name__ == "__main__":
    main()
