#!/usr/bin/env python
"""GPU box: the PRODUCT's own throughput — volrend_headless on the bench workload (tree.npz + transforms json on disk), serial
protocol vs --pipe N (frame graphs, pinned RGBA8 ring), optionally frame-sharded over several GPUs — next to the same loop
through the C ABI from Python (bench.Rig.e2e).  Prints one JSON line; used by tests/test_cli.py and the multi-GPU runs.
    python tools/cli_bench.py [--gpus N] [--pipe 8] [--frames 600]"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from rt_octree_b200 import synthetic as S  # noqa: E402

CLI = os.path.join(ROOT, "rt_octree_b200", "bin", "volrend_headless")


def workload_files(d):
    os.makedirs(d, exist_ok=True)
    npz, pj, oj, ts = [os.path.join(d, n) for n in ("tree.npz", "transforms_test.json", "opt.json", "ts_latest.ts.npz")]
    if not os.path.exists(npz):
        t = bench.load_tree()
        t = dict(t, child=np.asarray(t["child"]), data=np.asarray(t["data"]))
        S.write_tree_npz(npz + ".tmp.npz", t)
        os.replace(npz + ".tmp.npz", npz)
    S.write_blender_json(pj, S.make_poses(bench.N_POSES))
    S.write_opt_json(oj, spp=bench.SPP, denoise=True)
    np.savez(ts, **S.make_guidance_weights(0))
    return npz, pj, oj, ts[:-4]


def run_cli(files, extra, frames):
    npz, pj, oj, ts = files
    reps = max(1, frames // bench.N_POSES)
    # the CLI renders the pose list once; a longer timed loop = the list repeated (one json with `reps` copies)
    pj_rep = pj
    if reps > 1:
        pj_rep = pj[:-5] + "_x%d.json" % reps
        if not os.path.exists(pj_rep):
            S.write_blender_json(pj_rep, np.concatenate([S.make_poses(bench.N_POSES)] * reps))
    t0 = time.perf_counter()
    r = subprocess.run([CLI, npz, pj_rep, "--options", oj, "--ts_module", ts, "--warmup", "50", *extra], capture_output=True, text=True, timeout=900)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-1000:])
    out = {"process_wall_s": wall}
    for key, pat in (("fps", r"^FPS:\s+([0-9.]+)"), ("wall_fps", r"wall-clock FPS ([0-9.]+)"), ("aggregate_wall_fps", r"aggregate wall FPS: ([0-9.]+)")):
        m = re.search(pat, r.stdout, re.M)
        if m:
            out[key] = float(m.group(1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--pipe", type=int, default=8)
    ap.add_argument("--frames", type=int, default=600)
    ap.add_argument("--no-api", action="store_true")
    a = ap.parse_args()
    files = workload_files(os.path.join(bench.CACHE, "cli"))
    res = {"frames": a.frames, "pipe": a.pipe}
    res["cli_serial"] = run_cli(files, [], min(a.frames, 400))
    res["cli_pipe_rgba8"] = run_cli(files, ["--pipe", str(a.pipe), "--readback", "rgba8"], a.frames)
    res["cli_pipe_no_graph_rgba8"] = run_cli(files, ["--pipe", str(a.pipe), "--readback", "rgba8", "--no_graph"], a.frames)
    if a.gpus > 1:
        res["cli_pipe_rgba8_%dgpu" % a.gpus] = run_cli(files, ["--pipe", str(a.pipe), "--readback", "rgba8", "--num_gpus", str(a.gpus)], a.frames * a.gpus)
    if not a.no_api:
        import torch

        from rt_octree_b200 import capi

        capi.set_device(0)
        tree = bench.load_tree()
        poses, fx = bench.workload_poses()
        rig = bench.Rig(capi, torch, tree, S.make_guidance_weights(0), bench.W, bench.H, fx, bench.SPP, True, poses, a.pipe)
        e = rig.e2e(list(range(bench.N_POSES)), a.pipe, "rgba8", 10, 0.5, lambda: None, graph=True)
        res["api_e2e_rgba8"] = {"fps": e["frames"] / e["seconds"]}
        res["cli_over_api"] = res["cli_pipe_rgba8"]["fps"] / res["api_e2e_rgba8"]["fps"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
