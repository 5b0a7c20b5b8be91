#!/usr/bin/env python
"""Diagnostic (GPU box): per-warp-tile timeline of ONE render launch of the bench frame, from a library built with
tools/build_variant.sh tilelog "-DRTO_TILE_LOG".  Writes gpurun_out/tile_log_<tag>.npz: rows = [tile, sm, t0, t1, max_steps,
sum_steps, loop_ns, hits] per 8x4 warp tile (ns from %globaltimer), plus the kernel's own start.
    RTO_LIB=build/var_tilelog/librtoctree_b200.so python tools/tile_log.py [tag] [streams_busy]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rt_octree_b200 import capi, synthetic as S  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "serial"
busy = int(sys.argv[2]) if len(sys.argv) > 2 else 0
capi.set_device(0)
L = capi.load()
tree = bench.load_tree()
poses, fx = bench.workload_poses()
t = capi.N3Tree(tree)
net = capi.Denoiser(S.make_guidance_weights(0))
cam = capi.Camera(bench.W, bench.H, fx, fx)
opt = capi.RenderOptions()
opt.spp, opt.denoise = 6, True
ctxs = [capi.RenderContext(bench.W, bench.H) for _ in range(1 + busy)]
streams = [torch.cuda.Stream() for _ in range(1 + busy)]
TW, TH = int(os.environ.get("TILE_W", 8)), int(os.environ.get("TILE_H", 4))
n_tiles = ((bench.W + 2 * TW - 1) // (2 * TW)) * ((bench.H + 2 * TH - 1) // (2 * TH)) * 4
log = torch.zeros((n_tiles, 8), dtype=torch.int64, device="cuda")
for f in range(20):     # warm-up without logging
    for k in range(1 + busy):
        cam.transform = poses[(f + k) % 200]
        ctxs[k].rng_set_frame(f + k)
        capi.launch_renderer(t, cam, opt, ctxs[k], stream=streams[k].cuda_stream)
        net.denoise(cam, ctxs[k], stream=streams[k].cuda_stream)
torch.cuda.synchronize()
out = {}
for f in (17, 60, 117):
    log.zero_()
    # other streams keep the GPU busy with neighbouring frames (the pipelined regime) when busy > 0
    for k in range(1, 1 + busy):
        for j in range(3):
            cam.transform = poses[(f + k + j) % 200]
            ctxs[k].rng_set_frame(f + k + j)
            capi.launch_renderer(t, cam, opt, ctxs[k], stream=streams[k].cuda_stream)
            net.denoise(cam, ctxs[k], stream=streams[k].cuda_stream)
    assert L.rto_debug_set_tile_log(C.c_void_p(log.data_ptr())) == 0
    cam.transform = poses[f]
    ctxs[0].rng_set_frame(f)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    capi.launch_renderer(t, cam, opt, ctxs[0], stream=streams[0].cuda_stream)
    e1.record(streams[0])
    torch.cuda.synchronize()
    assert L.rto_debug_set_tile_log(C.c_void_p(0)) == 0
    out["f%d" % f] = log.cpu().numpy().copy()
    out["ms%d" % f] = np.float64(e0.elapsed_time(e1))
    print(tag, "frame", f, "kernel ms", e0.elapsed_time(e1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tile_log_%s.npz" % tag), **out)
