#!/usr/bin/env python
"""A/B of the mid-size fp32 Riccati sweep: FFMA register tiles (IDOC_MID_MMA=0) vs 3xTF32 mma.sync (IDOC_MID_MMA=1), per
shape, time-varying and time-invariant, device-resident, CUDA events per kernel.  One JSON line per case."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idoc_b200 as ib  # noqa: E402

L = ib._lib


def run(n, m, T, B, ti, dtype=np.float32):
    res = {}
    if ti:
        src = ib.lqr.DeviceProblem.generate(B, 1, n, m, dtype, seed=5)
        plan = ib.tilqr.DevicePlan(B, T, n, m, dtype, vjp=False)
        for k in ("Q", "R", "A", "B", "x0"):
            L.check(L.lib.idoc_memcpy_d2d(plan.arr[k].ptr, src.arr[k].ptr, plan.arr[k].nbytes, None), "d2d")
        L.check(L.lib.idoc_memcpy_d2d(plan.arr["Qf"].ptr, src.arr["Qf"].ptr, plan.arr["Qf"].nbytes, None), "d2d")
        step = plan.fwd
        getx = lambda: plan.X.numpy()
    else:
        prob = ib.lqr.DeviceProblem.generate(B, T, n, m, dtype, seed=5)
        plan = ib.lqr.DevicePlan(B, T, n, m, dtype, vjp=False)
        step = lambda: plan.fwd(prob)
        getx = lambda: plan.X.numpy()
    for mma in ("0", "1", "g1"):  # FFMA, tensor-core sweep (default problems per warp), one problem per warp
        os.environ["IDOC_MID_MMA"] = "0" if mma == "0" else "1"
        os.environ.pop("IDOC_MID_MMA_G", None)
        os.environ.pop("IDOC_MID_MMA_NW", None)
        if mma == "g1":  # the other geometry: one problem per warp at n = 16, two warps per problem at n = 32
            os.environ["IDOC_MID_MMA_G"] = "1"
            os.environ["IDOC_MID_MMA_NW"] = "2"
        step()
        L.sync()
        L.profile_begin()
        for _ in range(3):
            step()
        L.sync()
        k = L.profile_end()
        res[mma] = (float(np.mean(k["lqr_riccati"])), getx())
    d = float(np.max(np.abs(res["1"][1] - res["0"][1])) / np.max(np.abs(res["0"][1])))
    print(json.dumps(dict(dtype=np.dtype(dtype).name, n=n, m=m, T=T, B=B, ti=ti, riccati_ms_ffma=res["0"][0], riccati_ms_mma=res["1"][0],
                          riccati_ms_mma_other_geometry=res["g1"][0] if "g1" in res else None,
                          speedup=res["0"][0] / res["1"][0], max_rel_diff_X=d)), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "f64":
        run(32, 16, 200, 2048, False, np.float64)
        run(32, 16, 200, 2048, True, np.float64)
        run(32, 8, 200, 2048, False, np.float64)
        run(16, 8, 200, 32768, True, np.float64)
        run(16, 8, 200, 8192, False, np.float64)
        run(16, 16, 200, 8192, False, np.float64)
        sys.exit(0)
    run(32, 16, 200, 4096, False)
    run(32, 16, 200, 4096, True)
    run(32, 8, 200, 4096, False)
    run(16, 8, 200, 65536, True)
    run(16, 8, 200, 16384, False)
    run(16, 16, 200, 16384, False)
