#!/usr/bin/env python
"""Driver of tools/coherence_sim.cpp: distinct cache lines / sectors per warp-level load of the marching loop for alternative
grid layouts and warp-tile shapes, on the bench workload (CPU only, no GPU needed).

    python tools/coherence_sim.py [--depth 9] [--poses 0,50,100] [--stride 4]

Calibration: for the shipped layout ncu reports 6.2 lines / 10.7 sectors per byte-brick load and 3.1 / 3.1 per table load
(profiles/r01_render_v8_ncu_full.txt, source page)."""
import argparse
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rt_octree_b200 import synthetic as S  # noqa: E402


def build():
    so = os.path.join(ROOT, "build", "libcoherence_sim.so")
    src = os.path.join(ROOT, "tools", "coherence_sim.cpp")
    deps = [src, os.path.join(ROOT, "rt_octree_b200", "csrc", "rto_ray.cuh"), os.path.join(ROOT, "rt_octree_b200", "csrc", "rto_grid_host.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-mf16c", "-fPIC", "-shared", src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.coherence_layout_name.restype = C.c_char_p
    return lib


def frame_rng(frame, warmup=100, seed=20230418):
    """ctx.rng when the headless driver renders pose `frame`: pcg32(seed) advanced by (warmup + frame) * 2^32
    (pcg32.h:53-59,145-166; stated here in plain Python so that the tool does not touch oracle/)."""
    M, MASK = 6364136223846793005, (1 << 64) - 1
    inc = 3                                         # (initseq = 1) << 1 | 1
    state = (0 * M + inc) & MASK                    # first next() on state 0
    state = (state + seed) & MASK
    state = (state * M + inc) & MASK
    delta = ((warmup + frame) << 32) & MASK
    cur_mult, cur_plus, acc_mult, acc_plus = M, inc, 1, 0
    while delta > 0:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & MASK
            acc_plus = (acc_plus * cur_mult + cur_plus) & MASK
        cur_plus = ((cur_mult + 1) * cur_plus) & MASK
        cur_mult = (cur_mult * cur_mult) & MASK
        delta >>= 1
    return (acc_mult * state + acc_plus) & MASK, inc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=9)
    ap.add_argument("--poses", default="0,50,100,150")
    ap.add_argument("--stride", type=int, default=4, help="simulate every stride-th warp tile")
    ap.add_argument("--size", type=int, default=800)
    a = ap.parse_args()
    lib = build()
    tree = S.make_tree(depth=a.depth, shell=1.0, halo=0.25, seed=0)
    child = np.ascontiguousarray(tree["child"].reshape(-1), np.int32)
    data = np.ascontiguousarray(tree["data"].reshape(-1)).view(np.uint16)
    poses = S.poses_to_c2w12(S.make_poses(200))
    W = H = a.size
    fx = S.blender_focal(W)
    off = np.ascontiguousarray(tree["offset"], np.float32)
    sc = np.ascontiguousarray(tree["invradius3"], np.float32)
    shapes = [(8, 4), (4, 8), (16, 2), (32, 1)]
    rows = {}
    for tw, th in shapes:
        acc, steps_all = None, 0.0
        for pi in [int(p) for p in a.poses.split(",")]:
            out = np.zeros(16 * 6, np.float64)
            nl = C.c_int(0)
            steps = C.c_double(0)
            st, inc = frame_rng(pi)
            pose = np.ascontiguousarray(poses[pi], np.float32)
            K = lib.coherence_sim(C.c_void_p(child.ctypes.data), C.c_void_p(data.ctypes.data), int(tree["data_dim"]),
                                  C.c_int64(child.size // 8), a.depth, C.c_void_p(pose.ctypes.data), C.c_void_p(off.ctypes.data),
                                  C.c_void_p(sc.ctypes.data), C.c_float(fx), C.c_float(fx), W, H, C.c_uint64(st), C.c_uint64(inc),
                                  tw, th, a.stride, C.c_void_p(out.ctypes.data), C.byref(nl), C.byref(steps))
            assert K > 0
            o = out[: nl.value * 6].reshape(nl.value, 6)
            acc = o if acc is None else acc + o
            steps_all += steps.value
        rows[(tw, th)] = (acc, steps_all)
    print("bench tree depth %d, %dx%d, poses %s, every %d-th tile; per WARP-LEVEL load: lines (128 B) / sectors (32 B)" % (
        a.depth, W, H, a.poses, a.stride))
    for (tw, th), (acc, steps) in rows.items():
        print("\nwarp tile %dx%d  (%.1f M ray steps, %.2f M warp iterations, lane efficiency %.1f / 32)" % (
            tw, th, steps / 1e6, acc[0, 0] / 1e6, steps / acc[0, 0]))
        print("  %-34s %14s %14s %10s" % ("layout", "table ln / sec", "brick ln / sec", "brick/iter"))
        for i in range(acc.shape[0]):
            t0, tl, ts, b0, bl, bs = acc[i]
            print("  %-34s %6.2f / %5.2f %7.2f / %5.2f %10.2f" % (lib.coherence_layout_name(i).decode(), tl / t0, ts / t0, bl / b0, bs / b0, b0 / t0))


if __name__ == "__main__":
    main()
