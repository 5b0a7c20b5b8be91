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
                # the halo follows the net: 2 rows for the two 3x3 convolutions + one per filter level
                for L in (1, 4, 6):
                    assert SH.render_rows_for_band(b, H, True, levels=L) == (max(0, b[0] - 2 - L), min(H, b[1] + 2 + L))
    assert SH.denoise_halo(4) == SH.DENOISE_HALO == 6 and SH.denoise_halo(6) == 8


def test_rebalance_bands():
    """Band balancer of the tile split: endpoints fixed, boundaries strictly increasing with a minimum band height, and the
    bands move towards equal cost (iterating on a fixed cost profile converges to it)."""
    H, n = 2160, 8
    rows = np.arange(H)
    density = 1.0 + 3.0 * np.exp(-((rows - 1000.0) / 300.0) ** 2)        # heavy in the middle, like a frame through the object
    bounds = [H * g // n for g in range(n + 1)]
    spread0 = None
    for it in range(12):
        ms = [float(density[bounds[g]:bounds[g + 1]].sum()) for g in range(n)]
        if spread0 is None:
            spread0 = max(ms) / min(ms)
        bounds = SH.rebalance_bands(bounds, ms, H)
        assert bounds[0] == 0 and bounds[-1] == H and all(bounds[g + 1] - bounds[g] >= 16 for g in range(n))
    ms = [float(density[bounds[g]:bounds[g + 1]].sum()) for g in range(n)]
    assert spread0 > 2.5 and max(ms) / min(ms) < 1.1
    # degenerate inputs: zero times, two bands, tiny image
    assert SH.rebalance_bands([0, 50, 100], [0.0, 0.0], 100) == [0, 50, 100]
    assert SH.rebalance_bands([0, 20, 40], [1.0, 100.0], 40) == [0, 24, 40]      # min_rows keeps both bands >= 16 rows


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


def _peer_worker(rank, world, port, out, denoise):
    """One rank of the peer-store tile split; both ranks share cuda:0 (the IPC mapping works between processes on one
    device exactly as between devices), the completion barrier goes through gloo."""
    import numpy as np

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rt_octree_b200 import capi, synthetic as S

    torch.cuda.set_device(0)
    capi.set_device(0)
    tree = S.make_tree(depth=7, shell=1.0, halo=0.1, seed=0)
    W, H = 200, 151
    fx = S.blender_focal(W)
    poses = S.poses_to_c2w12(S.make_poses(8))
    t = capi.N3Tree(tree)
    net = capi.Denoiser(S.make_guidance_weights(0))
    cam = capi.Camera(W, H, fx, fx)
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = 6, denoise
    ts = SH.PeerTileSplit(capi, dist, rank, world, W, H, net.levels)
    ok = True
    for f in (2, 5):
        ts.render(t, net, cam, opt, poses[f], f)
        dist.barrier()
        if rank == 0:
            got, got8 = ts.ctx.read_image().copy(), ts.ctx.read_image_rgba8().copy()
            c1 = capi.RenderContext(W, H)
            cam.transform = poses[f]
            c1.rng_set_frame(f)
            capi.launch_renderer(t, cam, opt, c1)
            if denoise:
                net.denoise(cam, c1)
            ok = ok and bool(np.array_equal(c1.read_image(), got)) and bool(np.array_equal(c1.read_image_rgba8(), got8))
            ok = ok and got[..., :3].min() < 0.9
            c1.close()
        dist.barrier()
    if rank == 0:
        out.put(bool(ok))
    ts.close()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("denoise", [True, False])
def test_peer_store_tile_split_two_processes(denoise):
    """Tile split with peer-direct stores between two PROCESSES (CUDA IPC mapping of rank 0's image; here both on cuda:0,
    on a multi-GPU box tools/tile_split_check.py runs it over NVLink): the frame assembled in rank 0's buffers by the two
    bands' own filter epilogues equals the single-context frame bit for bit, float image and RGBA8 copy."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90) + (1 if denoise else 0)
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q, denoise)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
