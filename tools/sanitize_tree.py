#!/usr/bin/env python
"""Small driver for compute-sanitizer: the device tree loader (dense + quantised, structure check, grid build, plane
read-back) and the marching kernel in its three forms (fused index, v9 loop, word plane).  Run by tools/gpu_sanitize.sh."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rt_octree_b200 import capi, synthetic as S  # noqa: E402


def main():
    tree = S.make_tree(depth=6, shell=1.0, halo=0.1, seed=2)
    cap = tree["child"].shape[0]
    rs = np.random.default_rng(0)
    z = {k: tree[k] for k in ("data_dim", "data_format", "invradius3", "offset", "child")}
    z["quant_colors"] = rs.normal(size=(8, 65536, 3)).astype(np.float16)
    z["quant_map"] = rs.integers(0, 65536, size=(8, cap, 2, 2, 2)).astype(np.uint16)
    z["sigma"] = np.ascontiguousarray(tree["data"][..., -1])
    z["data_retained"] = rs.normal(size=(1, cap, 2, 2, 2, 3)).astype(np.float16)
    W, H = 96, 64
    cam = capi.Camera(W, H, S.blender_focal(W))
    cam.transform = S.poses_to_c2w12(S.make_poses(8))[2]
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = 6, False
    for src in (tree, z):
        t = capi.N3Tree(src)
        for name in capi.N3Tree.PLANES:
            t.read_plane(name)
        ctx = capi.RenderContext(W, H)
        out = []
        # production (fused-index marcher over the march table + byte plane), the v9 loop, the word plane
        for env in ({}, {"RTO_FUSED_INDEX": "0"}, {"RTO_GRID8": "0"}):
            os.environ.update(env)
            ctx.rng_set_frame(2)
            capi.launch_renderer(t, cam, opt, ctx)
            out.append(ctx.read_aux().copy())
            for k in env:
                del os.environ[k]
        assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2]) and out[0][3].max() == 1.0
        ctx.close()
        t.close()
    print("sanitize_tree OK")


if __name__ == "__main__":
    main()
