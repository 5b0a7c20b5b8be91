#!/usr/bin/env python
"""Target program for ncu (GPU box): renders a few frames of one BASELINE config serially, nothing else.
    python tools/profile_run.py {bench|spp1|tt} [frames]
bench = lego depth 9, 800x800, SPP 6 + denoise (config 3); spp1 = same tree, SPP 1, denoiser off (config 2);
tt = T&T-shaped depth-10 tree, 1920x1080, SPP 6 + denoise (config 4)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rt_octree_b200 import capi, synthetic as S  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "bench"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
capi.set_device(0)
if cfg == "tt":
    tree = bench.load_tree(0, None, bench.TT_TREE_KW, "tt")
    poses = S.poses_to_c2w12(S.make_poses(bench.N_POSES, **bench.TT_POSES_KW))
    W, H, fx, spp, den = bench.TT_W, bench.TT_H, bench.TT_FX, 6, True
else:
    tree = bench.load_tree()
    poses, fx = bench.workload_poses()
    W, H = bench.W, bench.H
    spp, den = (1, False) if cfg == "spp1" else (6, True)
t = capi.N3Tree(tree)
net = capi.Denoiser(S.make_guidance_weights(0)) if den else None
cam = capi.Camera(W, H, fx, fx)
opt = capi.RenderOptions()
opt.spp, opt.denoise = spp, den
ctx = capi.RenderContext(W, H)
ctx.image_rgba8_ptr   # the producing kernel writes the RGBA8 copy too, as in the product's e2e path
for f in range(frames):
    cam.transform = poses[f % len(poses)]
    ctx.rng_set_frame(f, bench.WARMUP_RNG)
    capi.launch_renderer(t, cam, opt, ctx)
    if net:
        net.denoise(cam, ctx)
capi.synchronize()
print("rendered", frames, "frames of", cfg, "checksum", float(np.abs(ctx.read_image()).sum()))
