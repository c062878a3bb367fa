#!/usr/bin/env python
"""Short driver for ncu captures of the mid-size Riccati sweep (fp32): FFMA and MMA variants back to back.
    python tools/prof_mid.py [n m T B ti]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib
n, m, T, B, ti = (int(x) for x in (sys.argv[1:6] + ["32", "16", "40", "2048", "0"][len(sys.argv) - 1:]))
if ti:
    src = ib.lqr.DeviceProblem.generate(B, 1, n, m, np.float32, seed=5)
    plan = ib.tilqr.DevicePlan(B, T, n, m, np.float32, vjp=False)
    for k in ("Q", "R", "A", "B", "x0", "Qf"):
        L.check(L.lib.idoc_memcpy_d2d(plan.arr[k].ptr, src.arr[k].ptr, plan.arr[k].nbytes, None), "d2d")
    step = plan.fwd
else:
    prob = ib.lqr.DeviceProblem.generate(B, T, n, m, np.float32, seed=5)
    plan = ib.lqr.DevicePlan(B, T, n, m, np.float32, vjp=False)
    step = lambda: plan.fwd(prob)
for mma in ("0", "1", "0", "1"):
    os.environ["IDOC_MID_MMA"] = mma
    step()
    L.sync()
print("ok")
