#!/usr/bin/env python
"""GPU box: latency mode (render_kernel_split) against throughput mode on the bench frame: bit-identical buffers + timing.
Run under `timeout`: a protocol bug in the hand-off queue would show up as a hang."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rt_octree_b200 import capi, synthetic as S  # noqa: E402

capi.set_device(0)
small = len(sys.argv) > 1 and sys.argv[1] == "small"
if small:
    tree = S.make_tree(depth=7, shell=1.0, halo=0.1, seed=0)
    W, H = 320, 240
else:
    tree = bench.load_tree()
    W, H = bench.W, bench.H
poses, _ = bench.workload_poses()
fx = float(np.float32(S.blender_focal(W)))
t = capi.N3Tree(tree)
cam = capi.Camera(W, H, fx, fx)
for spp, den in ((6, True), (1, False), (8, True)):
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = spp, den
    a, b = capi.RenderContext(W, H), capi.RenderContext(W, H)
    b.set_mode(True)
    for f in (17, 60, 117, 3):
        cam.transform = poses[f]
        for c in (a, b):
            c.rng_set_frame(f)
            capi.launch_renderer(t, cam, opt, c)
        capi.synchronize()
        ok = np.array_equal(a.read_aux(), b.read_aux()) and (den or np.array_equal(a.read_image(), b.read_image()))
        print("spp", spp, "frame", f, "identical" if ok else "DIFFERENT", flush=True)
        assert ok
    # timing, 100 frames each, serial
    for name, c in (("throughput", a), ("latency", b), ("latency-nodonate", b)):
        os.environ["RTO_SPLIT_NODONATE"] = "1" if name.endswith("nodonate") else "0"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            e0.record()
            for f in range(100):
                cam.transform = poses[f]
                c.rng_set_frame(f)
                capi.launch_renderer(t, cam, opt, c)
            e1.record()
            torch.cuda.synchronize()
        print("spp", spp, name, "mode: %.4f ms / frame (render only, back to back)" % (e0.elapsed_time(e1) / 100), flush=True)
print("OK")
