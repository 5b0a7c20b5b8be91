#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): single-frame TILE SPLIT over N GPUs (rt_octree_b200/sharding.py)
reproduces the single-GPU frame bit for bit, and reports the single-frame latency of both exchange schemes: the filter
epilogue's peer-direct stores into rank 0's image (PeerTileSplit, the product path) and one NCCL gather of the bands.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/tile_split_check.py --width 3840 --height 2160
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rt_octree_b200 import capi, sharding as SH, synthetic as S  # noqa: E402


def image_view(ctx):
    """torch view [H, W, 4] of the context's device image (no copy)."""
    n = ctx.height * ctx.width * 4

    class _Arr:   # __cuda_array_interface__ wrapper around the raw device pointer
        __cuda_array_interface__ = {"shape": (ctx.height, ctx.width, 4), "typestr": "<f4", "data": (ctx.image_ptr, False), "version": 2}

    return torch.as_tensor(_Arr(), device="cuda")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--frames", type=int, default=8)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    capi.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tree = S.make_tree(depth=a.depth, shell=1.0, halo=0.1, seed=0)
    poses = S.poses_to_c2w12(S.make_poses(a.frames))
    W, H = a.width, a.height
    fx = float(np.float32(S.blender_focal(W)))
    t = capi.N3Tree(tree)
    net = capi.Denoiser(S.make_guidance_weights(0))
    cam = capi.Camera(W, H, fx, fx)
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = 6, True
    ctx = capi.RenderContext(W, H)
    img = image_view(ctx)
    ok, lat = True, []
    for f in range(a.frames):
        cam.transform = poses[f]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        full = SH.render_frame_tile_split(capi, t, net, cam, opt, ctx, f, rank, world, img)
        torch.cuda.synchronize()
        lat.append(time.perf_counter() - t0)
        if rank == 0:
            # single-GPU reference frame on a second context
            ctx1 = capi.RenderContext(W, H)
            ctx1.rng_set_frame(f)
            capi.launch_renderer(t, cam, opt, ctx1)
            net.denoise(cam, ctx1)
            ref = torch.from_numpy(ctx1.read_image())
            same = bool(torch.equal(full.cpu(), ref))
            ok &= same
            ctx1.close()
    peer = None
    if world > 1:
        peer = SH.bench_tile_split(capi, torch, dist, tree, S.make_guidance_weights(0), poses, rank, world, local, frames=a.frames,
                                   width=W, height=H)
        ok &= bool(peer["bit_identical_to_single_gpu"]) if rank == 0 else True
    if rank == 0:
        print(json.dumps({"tile_split": {"n_gpus": world, "width": W, "height": H, "frames": a.frames, "bit_identical_to_single_gpu": ok,
                                         "gather_latency_ms_median": float(np.median(lat[1:]) * 1e3),
                                         "gather_bytes_per_rank": int(H // world * W * 16), "peer_store": peer}}))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
