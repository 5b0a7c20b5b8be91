"""Multi-GPU host logic on CPU: partition functions and the band gather over torch.distributed (gloo, world_size 2)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rt_octree_b200 import sharding as SH


def test_shard_frames_partition():
    for n in (1, 7, 200, 201):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = SH.shard_frames(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [SH.shard_frames(n, r, world)[1] - SH.shard_frames(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_tile_bands_and_halo():
    for H in (800, 1080, 2160, 37):
        for world in (1, 2, 4, 8):
            bands = SH.tile_bands(H, world)
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
            for b in bands:
                y0, y1 = SH.render_rows_for_band(b, H, denoise=True)
                assert y0 == max(0, b[0] - 6) and y1 == min(H, b[1] + 6)
                assert SH.render_rows_for_band(b, H, denoise=False) == b


def _worker(rank, world, port, H, W, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(H * W * 4, dtype=torch.float32).reshape(H, W, 4)
    bands = SH.tile_bands(H, world)
    b = bands[rank]
    res = SH.gather_bands(full[b[0]:b[1]].clone() + 0.0, bands, rank, world)
    if rank == 0:
        out.put(bool(torch.equal(res, full)))
    else:
        assert res is None
    # frame sharding needs no collective; a max-reduction of per-rank times is all bench.py does
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.destroy_process_group()


@pytest.mark.parametrize("H", [9, 16])
def test_gather_bands_gloo_world2(H):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400) + H
    procs = [ctx.Process(target=_worker, args=(r, 2, port, H, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
