#!/usr/bin/env python
"""Times fwd + VJP of every shape in csrc/mid_shapes.def through the mid-size kernels and through the shape-generic
kernels (IDOC_NO_MID=1) on the same device-resident problems.  One JSON line per (shape, dtype).

    python tools/mid_vs_generic.py [--T 100] [--mb 1500]
"""
import argparse
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import idoc_b200 as ib  # noqa: E402

L = ib._lib


def shapes():
    for line in open(os.path.join(ROOT, "idoc_b200", "csrc", "mid_shapes.def")):
        mm = re.match(r"\s*IDOC_MID_SHAPE\(([^)]*)\)", line)
        if mm:
            yield tuple(int(x) for x in mm.group(1).split(","))[:2]


def run(n, m, T, B, dtype, steps=3, warmup=2):
    stream = L.Stream()
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=7, stream=stream)
    plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True, reduce_batch=True)
    out = {}
    for no_mid in ("0", "1"):
        os.environ["IDOC_NO_MID"] = no_mid

        def step():
            plan.fwd(prob, stream=stream)
            plan.vjp(prob, plan.X, plan.U, None, stream=stream)

        for _ in range(warmup):
            step()
        stream.sync()
        e0, e1 = L.Event(), L.Event()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        stream.sync()
        out[no_mid] = e0.elapsed_ms(e1) / steps
    os.environ["IDOC_NO_MID"] = "0"
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=100)
    ap.add_argument("--mb", type=float, default=1500.0, help="problem data per run (MB) -> batch size")
    a = ap.parse_args()
    for n, m in shapes():
        for dtype in (np.float32, np.float64):
            per = (2 * n * n + 2 * n * m + m * m + 2 * n + m) * a.T * np.dtype(dtype).itemsize
            B = max(256, int(a.mb * 1e6 / per) // 256 * 256)
            r = run(n, m, a.T, B, dtype)
            print(json.dumps(dict(n=n, m=m, T=a.T, batch=B, dtype=np.dtype(dtype).name, mid_ms=round(r["0"], 3),
                                  generic_ms=round(r["1"], 3), speedup=round(r["1"] / r["0"], 2),
                                  mid_solves_per_s=round(B / r["0"] * 1e3))), flush=True)
