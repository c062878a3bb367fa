"""Worker of tests/test_gpu_dist.py: one rank per GPU (torchrun), NCCL all-reduce of batch-summed gradients verified against
the per-problem gradients of every rank (idoc_b200.dist.verify_allreduced_gradients).  Exit code 0 = within tolerance."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    import idoc_b200 as ib
    from idoc_b200 import dist as D
    ib._lib.check(ib._lib.lib.idoc_set_device(local), "set_device")
    red = D.GradAllReducer(rank, world, D.exchange_unique_id(rank, world))
    worst = {}
    for (n, m, T, B) in ((8, 4, 20, 300), (32, 16, 8, 24), (5, 3, 6, 17)):
        for dt, tol in ((np.float64, 1e-10), (np.float32, 1e-4)):
            e = D.verify_allreduced_gradients(red, n, m, T, B, dt)
            worst[(n, m, T, B, np.dtype(dt).name)] = e
            assert e < tol, (n, m, T, B, dt, e)
    if rank == 0:
        print("allreduce verified:", worst, flush=True)
    red.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
