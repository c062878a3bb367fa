#!/usr/bin/env python
"""Host <-> device copy bandwidth with N ranks copying AT THE SAME TIME (torchrun, one rank per GPU): the ceiling of bench.py's
`e2e` leg at N GPUs.  Every rank moves `--mb` MiB pinned host -> device and device -> pinned host concurrently (two streams),
all ranks start behind a barrier; prints per-rank and aggregate GB/s (rank 0, one JSON line).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_ranks.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
    import idoc_b200 as ib
    L = ib._lib
    L.check(L.lib.idoc_set_device(local), "set_device")
    nbytes = args.mb << 20
    h_in, h_out = L.PinnedArray((nbytes,), np.uint8), L.PinnedArray((nbytes,), np.uint8)
    d_in, d_out = L.DeviceArray((nbytes,), np.uint8), L.DeviceArray((nbytes,), np.uint8, zero=True)
    h_in.a[:] = 1
    s0, s1 = L.Stream(), L.Stream()
    res = {}
    for mode in ("h2d", "d2h", "duplex"):
        best = 0.0
        for _ in range(args.reps):
            L.sync()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            if mode in ("h2d", "duplex"):
                L.check(L.lib.idoc_memcpy_h2d(d_in.ptr, h_in.ptr, nbytes, s0.ptr), "h2d")
            if mode in ("d2h", "duplex"):
                L.check(L.lib.idoc_memcpy_d2h(h_out.ptr, d_out.ptr, nbytes, s1.ptr), "d2h")
            s0.sync()
            s1.sync()
            dt = time.perf_counter() - t0
            if dist is not None:  # the slowest rank defines the aggregate
                import torch
                t = torch.tensor([dt], dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t[0])
            best = max(best, (2 if mode == "duplex" else 1) * nbytes / dt / 1e9)
        res[mode] = best
    if rank == 0:
        print(json.dumps(dict(ranks=world, mb_per_direction=args.mb,
                              per_rank_GBs={k: v for k, v in res.items()},
                              aggregate_GBs={k: v * world for k, v in res.items()},
                              note="duplex = H2D + D2H bytes of one rank per second with both directions busy; every rank timed to "
                                   "the slowest one")), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
