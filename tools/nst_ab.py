#!/usr/bin/env python
"""A/B of the number of cp.async prefetch stages (IDOC_MID_NST) of the mid-size vector sweeps (rollout, adjoint backward,
adjoint forward; csrc/mid.cuh): config 3 (tilqr n=16 m=8 T=200) and the config-5 shape (n=32 m=16 T=200), per-kernel CUDA-event
times.  One JSON line per case."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs as bc  # noqa: E402

if __name__ == "__main__":
    for leg, dtype, B, stages in ((bc.leg_tilqr, np.float32, 131072, (2, 3, 4)), (bc.leg_tilqr, np.float64, 65536, (2, 3, 4)),
                                  (bc.leg_lqr32, np.float32, 4096, (2, 3)), (bc.leg_lqr32, np.float64, 2048, (2, 3))):
        for nst in stages:
            os.environ["IDOC_MID_NST"] = str(nst)
            r = leg(dtype, B, 3, 1)
            print(json.dumps(dict(leg=r["leg"], dtype=r["dtype"], batch=B, stages=nst, ms_per_step=round(r["ms_per_step"], 3),
                                  solves_per_s=round(r["value"]), kernels=dict(bc.KERNELS))), flush=True)
