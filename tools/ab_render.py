#!/usr/bin/env python
"""A/B of render-kernel variants on the bench workload (GPU box): serial stage time (Timer events) and pipelined throughput.
    python tools/ab_render.py name=ENV1=v,ENV2=v name2=... [--lib name=path]
Each variant runs in a fresh subprocess (environment read at library load); variants are interleaved twice."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, os, json, time
sys.path.insert(0, %r)
import numpy as np, torch
import bench
from rt_octree_b200 import capi, synthetic as S
capi.set_device(0)
tree = bench.load_tree(); poses, fx = bench.workload_poses()
spp = int(os.environ.get("AB_SPP", "6")); den = spp != 1
rig = bench.Rig(capi, torch, tree, S.make_guidance_weights(0), bench.W, bench.H, fx, spp, den, poses, 8)
frames = list(range(200))
sp = rig.serial_protocol(frames, 0.6)
pl = rig.pipelined(frames, 4, 10, 0.6, lambda: None, graph=os.environ.get("AB_GRAPH", "1") == "1")
print(json.dumps({"render_ms": sp["render_ms"], "net_ms": sp["net_ms"], "filter_ms": sp["filter_ms"],
                  "serial_fps": 1e3 / (sp["render_ms"] + sp["net_ms"] + sp["filter_ms"]), "serial_wall_fps": sp["wall_fps"],
                  "pipe_fps": 1e3 * pl["reps"] * 200 / pl["ms_total"]}))
''' % ROOT

variants = []
for a in sys.argv[1:]:
    name, _, envs = a.partition("=")
    env = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
    variants.append((name, env))
res = {n: [] for n, _ in variants}
for rep in range(2):
    for name, env in variants:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, "-c", WORKER], env=e, capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            print(name, "FAILED", r.stderr[-800:])
            continue
        d = json.loads(r.stdout.strip().splitlines()[-1])
        res[name].append(d)
        print(name, json.dumps(d), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ab_render.json"), "w"), indent=1)
