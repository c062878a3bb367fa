#!/usr/bin/env python
"""Occupancy sensitivity of each sweep kernel: pads the dynamic shared memory of one kernel at a time (tunable build)
so that fewer warps are resident per SM, and reports the kernel time."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib
dtype = np.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else np.float32
B, T, n, m = 65536, 100, 8, 4
stream = L.Stream()
prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=1234, stream=stream)
plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=True)
names = {0: "lqr_riccati", 1: "lqr_rollout", 3: "lqr_adj_bwd", 4: "lqr_adj_fwd"}


def timed():
    for _ in range(2):
        plan.fwd(prob, stream=stream)
        plan.vjp(prob, plan.X, plan.U, None, stream=stream)
    stream.sync()
    L.profile_begin()
    for _ in range(3):
        plan.fwd(prob, stream=stream)
        plan.vjp(prob, plan.X, plan.U, None, stream=stream)
    stream.sync()
    return {k: float(np.mean(v)) for k, v in L.profile_end().items()}


pads = [int(x) for x in os.environ.get("IDOC_PROBE_PADS", "0,3000,6000,9000,14000,20000,30000").split(",")]
kids = [int(x) for x in os.environ.get("IDOC_PROBE_KERNELS", "0,1,3,4").split(",")]
for kid, nm in names.items():
    if kid not in kids:
        continue
    os.environ["IDOC_PAD_KERNEL"] = str(kid)
    for pad in pads:
        os.environ["IDOC_SMEM_PAD"] = str(pad)
        print(f"{nm} pad={pad:6d}: {timed()[nm]:.3f} ms", flush=True)
