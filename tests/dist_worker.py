"""World-size-N gloo worker for tests/test_dist_gloo.py: exercises the multi-process host path (row-band sharding +
framebuffer gather + de-interleave) on CPU. The shards are produced by the CPU oracle (checker), the gather logic
under test is rtb200.dist.gather_frame."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "rust-raytracer_b200")); sys.path.insert(0, os.path.join(REPO, "oracle"))
import oracle_py as O   # noqa: E402
import rtb200 as R      # noqa: E402
from rtb200 import dist as RD, scenes   # noqa: E402


def main():
    rank, world, _ = RD.init("gloo")
    sc = scenes.cover_scene(40, 30, 2)
    h, w = sc.c.height, sc.c.width
    full_lin, full_img, _ = O.render(sc, threads=2)
    ok = True
    for band in (1, 4, 16):
        rows = R.shard_row_indices(h, rank, world, band)
        assert len(rows) == R.shard_rows(h, rank, world, band)
        # each rank renders ONLY its rows with the oracle (row-local RNG keys make shards independent)
        mine = np.zeros((len(rows), w, 3), np.uint8)
        for k, y in enumerate(rows):
            _, img, _ = O.render(sc, linear=False, rgb8=True, y0=int(y), y1=int(y) + 1, threads=1)
            mine[k] = img[y]
        rows_max = RD.padded_rows(h, world, band)
        shard = torch.zeros((rows_max, w, 3), dtype=torch.uint8)
        shard[: len(rows)] = torch.from_numpy(mine)
        frame = RD.gather_frame(shard, h, world, band, rank)
        if rank == 0:
            ok = ok and np.array_equal(frame.numpy(), full_img)
        else:
            assert frame is None
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST_OK" if ok else "DIST_FAIL", flush=True)
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
