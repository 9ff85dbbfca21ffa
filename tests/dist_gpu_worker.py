"""NCCL worker for tests/test_gpu_dist.py: every rank renders its interleaved rows on its own GPU, one gather to rank 0,
and rank 0 checks the frame against a single-GPU render of the whole image (bit-identical)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "rust-raytracer_b200"))
import rtb200 as R      # noqa: E402
from rtb200 import dist as RD, scenes   # noqa: E402


def main():
    rank, world, local = RD.init("nccl")
    sc = scenes.cover_scene(160, 90, 8)
    ok = True
    for band in (1, 8):
        rdr = RD.DistributedRenderer(sc, band_rows=band)
        st = rdr.render()
        torch.cuda.synchronize()
        rays = torch.tensor([float(st["rays"])], device="cuda", dtype=torch.float64)
        dist.all_reduce(rays)
        if rank == 0:
            full, st_full = R.render_rgb8(sc, R.make_options(device=local))
            ok = ok and np.array_equal(rdr.frame.cpu().numpy(), full) and int(rays.item()) == st_full["rays"]
        # frame loop without host waits: gathers overlap the next frame on a side stream, shards are double-buffered
        for _ in range(5):
            rdr.render_async()
        st2 = rdr.wait()
        torch.cuda.synchronize()
        if rank == 0:
            ok = ok and np.array_equal(rdr.frame.cpu().numpy(), full) and st2["frames"] == 5
        rdr.release()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_GPU_OK" if ok else "DIST_GPU_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
