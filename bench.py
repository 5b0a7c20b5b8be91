#!/usr/bin/env python
"""bench.py — headline benchmark of the RT-Octree hot path on B200 (BASELINE.json: FPS @800x800 SPP6 + denoise).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation on the host cores

One "step" = one frame of the hot path: ray generation + octree traversal + SH shade + aux write (1 kernel) and the
GuidanceNet + kernel-filter denoiser, on a synthetic lego-shaped PlenOctree (depth 9, data_dim 28) at 800x800, SPP 6,
a different test pose every step.  Frames shard across ranks with no collective (weak scaling: K frames per rank).
Every measurement repeats its K-frame block `reps` times so that the timed region lasts >= --min-seconds (0.5 s).
Prints ONE JSON line (rank 0).  See DESIGN.md §6 for every field.

Headline numbers of the line:
  value                      frames/s with --pipe frames in flight (device-timed, max over ranks): whole-job throughput
  value_reference_protocol   frames/s by the reference's own definition: 1000 / (render + net + filter ms), one
                             stream, cudaEvents per stage, host sync per frame — compare THIS with reference_cuda.fps
  e2e                        through the public API with host buffers: pose from host memory, RGBA8 image into pinned memory
  configs                    the other BASELINE configs measured the same way (SPP 1 / no denoise; T&T-shaped depth-10 tree at
                             1920x1080; the --write_buffer copy; the 4K tile split when WORLD_SIZE > 1)
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOAD = "lego-synthetic depth9 800x800 spp6 denoise"
W = H = 800
SPP = 6
N_POSES = 200
WARMUP_RNG = 100          # main_headless.cpp:469-479: 100 warm-up advances precede pose 0
TREE_KW = dict(depth=9, shell=1.0, halo=0.25, seed=0)
# BASELINE config 4: Tanks-and-Temples-shaped tree (anisotropic, depth 10) at the tt loader's 1920x1080 (main_headless.cpp:274-275)
TT_TREE_KW = dict(depth=10, shell=0.03, halo=0.02, seed=1, invradius3=(0.30, 0.42, 0.36), offset=(0.5, 0.52, 0.48))
TT_W, TT_H, TT_FX = 1920, 1080, 1166.0
TT_POSES_KW = dict(radius=3.2, elevation_deg=20.0)
CACHE = os.environ.get("RTO_CACHE", "/tmp/rto_cache")


# ------------------------------------------------------------------------------------------------ workload
def load_tree(rank=0, barrier=None, kw=None, tag="lego"):
    """Synthetic tree, generated once per box and cached as .npy under /tmp."""
    from rt_octree_b200 import synthetic as S

    kw = TREE_KW if kw is None else kw
    name = "%s_d%d_s%g_h%g_r%d" % (tag, kw["depth"], kw["shell"], kw["halo"], kw["seed"])
    fc, fd = os.path.join(CACHE, name + "_child.npy"), os.path.join(CACHE, name + "_data.npy")
    if rank == 0 and not (os.path.exists(fc) and os.path.exists(fd)):
        os.makedirs(CACHE, exist_ok=True)
        t = S.make_tree(**kw)
        np.save(fc + ".tmp.npy", t["child"])
        np.save(fd + ".tmp.npy", t["data"])
        os.replace(fc + ".tmp.npy", fc)
        os.replace(fd + ".tmp.npy", fd)
    if barrier:
        barrier()
    tree = {"data_dim": np.int64(28), "data_format": np.array("SH9"),
            "invradius3": np.asarray(kw.get("invradius3", (0.375,) * 3), np.float32),
            "offset": np.asarray(kw.get("offset", (0.5,) * 3), np.float32),
            "child": np.load(fc, mmap_mode="r"), "data": np.load(fd, mmap_mode="r")}
    return tree


def workload_poses():
    from rt_octree_b200 import synthetic as S

    return S.poses_to_c2w12(S.make_poses(N_POSES)), float(np.float32(S.blender_focal(W)))


def base_config(world, K):
    """The `config` object, identical for both arms (the reference arm measures THIS workload)."""
    return {"workload": WORKLOAD, "poses": N_POSES, "frames_per_rank": K, "parallelism": "frame-sharded x%d" % world,
            "tree": dict(TREE_KW), "width": W, "height": H, "spp": SPP, "denoise": True,
            "l2": "inputs larger than L2 (tree 1.1 GB of HBM planes), a different pose every step, no flush; cold-L2 variant in value_l2_flushed"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.p, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        return time.perf_counter()

    def stop(self, t_from=None, t_to=None):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.p.terminate()
        sm, smax, reasons = [], [], set()
        for ts, r in self.rows:
            if t_from is not None and not (t_from <= ts <= t_to):
                continue
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_frame_seconds(tree, poses, fx, weights, frames, nthreads, first_frame=0):
    """The reference's CPU implementation of the path, `frames` full frames of the workload:
    traversal/SH/composite = the reference's own trace_ray host-compiled (oracle/_ref/libref_cpu.so, OpenMP) when it
    was built, else the C port (oracle/rt_oracle.c, scalar); GuidanceNet = PyTorch CPU forward (deployed graph);
    filter = oracle C restatement of filtering.cu `applying` (OpenMP)."""
    import torch

    from oracle import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_ts_module as M

    kind = "reference" if O.ref_cpu_lib() is not None else "port"
    torch.set_num_threads(nthreads)
    net = M.DeployedGuidanceNet(weights).eval().float()   # CPU: fp32 math on the fp16 weights
    net.forward = _fp32_forward.__get__(net)
    t_render = t_net = t_filter = 0.0
    for f in range(first_frame, first_frame + frames):
        rng = O.frame_rng(f, WARMUP_RNG)
        t0 = time.perf_counter()
        if kind == "reference":
            aux = O.ref_cpu_render(tree, poses[f % len(poses)], W, H, fx, fx, SPP, rng, nthreads=nthreads)
        else:
            aux = O.render(tree, poses[f % len(poses)], W, H, fx, fx, SPP, rng, trace=False)["aux"]
        t1 = time.perf_counter()
        with torch.no_grad():
            wm, gm = net(torch.from_numpy(aux)[None])
        t2 = time.perf_counter()
        img_in = np.ones((H, W, 4), np.float32)
        img_in[..., :3] = np.transpose(aux[:3], (1, 2, 0))
        O.filtering(wm[0].numpy(), gm[0].numpy(), img_in)
        t3 = time.perf_counter()
        t_render += t1 - t0; t_net += t2 - t1; t_filter += t3 - t2
    per = (t_render + t_net + t_filter) / frames
    return per, kind, {"render_s": t_render / frames, "net_s": t_net / frames, "filter_s": t_filter / frames}


def _fp32_forward(self, aux_buffer):
    import torch.nn.functional as F

    x = F.relu6(F.conv2d(aux_buffer, self.w1.float(), self.b1.float(), padding="same"))
    x = F.relu6(F.conv2d(x, self.w2.float(), self.b2.float(), padding="same"))
    return F.softmax(x[:, :self.levels], dim=1), x[:, self.levels:]


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation on the host cores, same metric/config; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rt_octree_b200 import synthetic as S

    tree = load_tree()
    poses, fx = workload_poses()
    weights = S.make_guidance_weights(0)
    threads = os.cpu_count() or 1
    # every step = one full frame (bounded: --steps is clamped so the run ends within minutes)
    per0, kind, _ = cpu_frame_seconds(tree, poses, fx, weights, 1, threads)
    budget = 150.0
    steps = max(1, min(args.steps, int(budget / max(per0, 1e-3))))
    warm = max(0, min(args.warmup, int(20.0 / max(per0, 1e-3))))
    if warm:
        cpu_frame_seconds(tree, poses, fx, weights, warm, threads)
    per, kind, br = cpu_frame_seconds(tree, poses, fx, weights, steps, threads)
    fps = 1.0 / per
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": "fps_800x800_spp6_denoise", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(world, args.steps),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                             "sample": "%d full 800x800 SPP6 frames (trace_ray host build + torch CPU GuidanceNet + filter)" % steps,
                             "breakdown_s": br},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ CUDA arm
def algorithmic_bytes(capi, tree_h, ctx, cam, opt, poses, frames, w, h):
    """per ray sum_steps(4*d_s + 2) + 54*n_leaf + 32 (aux) [+16 image when denoise is off], with
    d_s = child look-ups of the REFERENCE's root-restart query — exact counters from the PRODUCTION marcher with the
    traversal record switched on (untimed)."""
    import torch

    n = w * h
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    dsum = torch.zeros(n, dtype=torch.int32, device="cuda")
    hits = torch.zeros(n, dtype=torch.int32, device="cuda")
    loads = torch.zeros(n, dtype=torch.int32, device="cuda")
    tr = capi.TracePOD()
    tr.steps, tr.depth_sum, tr.n_hits, tr.n_loads = steps.data_ptr(), dsum.data_ptr(), hits.data_ptr(), loads.data_ptr()
    tr.marcher = 1
    tot = {"steps": 0, "depth_sum": 0, "hits": 0, "loads": 0}
    for f in frames:
        cam.transform = poses[f % len(poses)]
        ctx.rng_set_frame(f, WARMUP_RNG)
        capi.launch_renderer(tree_h, cam, opt, ctx, trace=tr)
        torch.cuda.synchronize()
        tot["steps"] += int(steps.sum()); tot["depth_sum"] += int(dsum.sum())
        tot["hits"] += int(hits.sum()); tot["loads"] += int(loads.sum())
    k = len(frames)
    per_frame = (4 * tot["depth_sum"] + 2 * tot["steps"] + 54 * tot["hits"]) / k + 32 * n + (0 if opt.denoise else 16 * n)
    return per_frame, {a: b / k for a, b in tot.items()}


def reference_cuda_fps(tree, poses, fx, weights, frames=40):
    """The reference's own CUDA renderer rebuilt for sm_100a (oracle/_ref/ref_driver), same box, same inputs,
    its own Timer protocol (100 warm-up frames, mean per-stage cudaEvent ms)."""
    drv = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if not os.path.exists(drv):
        return {"unavailable": "oracle/_ref/ref_driver not built"}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import make_ts_module as M
        from rt_octree_b200 import synthetic as S

        d = os.path.join(CACHE, "refcuda")
        os.makedirs(d, exist_ok=True)
        npz = os.path.join(d, "tree.npz")
        if not os.path.exists(npz):
            t = dict(tree)
            t["child"], t["data"] = np.asarray(tree["child"]), np.asarray(tree["data"])
            S.write_tree_npz(npz, t)
        np.ascontiguousarray(poses, np.float32).tofile(os.path.join(d, "poses.bin"))
        ts = M.make_ts(weights, os.path.join(d, "ts.ts"), device="cuda")
        r = subprocess.run([drv, npz, os.path.join(d, "poses.bin"), ts, str(W), str(H), repr(fx), repr(fx), str(SPP), "1", "-",
                            str(frames), "100"], capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            return {"unavailable": "ref_driver exit %d: %s" % (r.returncode, r.stderr[-300:])}
        out = {}
        for line in r.stdout.splitlines():
            for key in ("render", "torch", "filter", "all", "FPS"):
                if line.startswith(key + ":"):
                    out[key.lower() + ("_ms" if key != "FPS" else "")] = float(line.split(":")[1].split()[0])
        out["frames"] = frames
        out["how"] = "unmodified reference kernels (oracle/_ref/ref_driver), Timer::report protocol"
        return out
    except Exception as e:  # never fail the bench because of the side baseline
        return {"unavailable": repr(e)[:300]}


def render_kernel_name(info):
    """The instantiation launch_spp picks for this tree (rto_render_kernel.cuh): GRID = 10 + K is the fused-index marcher."""
    off = lambda k: os.environ.get(k, "1")[:1] == "0"
    if not info.grid_level:
        g = 0
    elif off("RTO_GRID8"):
        g = 1
    elif off("RTO_DEFER_HITS") or off("RTO_LEAF_PLANES"):
        g = 2
    elif off("RTO_FUSED_INDEX"):
        g = 3
    else:
        g = 10 + int(info.grid_level)
    return "render_kernel<%d,false,%d>" % (SPP, g)


class Rig:
    """One workload on this rank: tree + net + a ring of (context, stream) slots, and the three measurements."""

    def __init__(self, capi, torch, tree, weights, w, h, fx, spp, denoise, poses, n_slots):
        self.capi, self.torch = capi, torch
        self.w, self.h, self.poses = w, h, poses
        t0 = time.perf_counter()
        self.tree = capi.N3Tree(tree)
        self.load_s = time.perf_counter() - t0
        self.net = capi.Denoiser(weights) if denoise else None
        self.cam = capi.Camera(w, h, fx, fx)
        self.opt = capi.RenderOptions()
        self.opt.spp, self.opt.denoise = spp, denoise
        self.ctxs = [capi.RenderContext(w, h) for _ in range(n_slots)]
        self.streams = [torch.cuda.Stream() for _ in range(n_slots)]
        self.frames_cache = {}

    def close(self):
        for fr in self.frames_cache.values():
            for f in fr["frames"]:
                f.close()
        self.frames_cache = {}
        for c in self.ctxs:
            c.close()
        if self.net:
            self.net.close()
        self.tree.close()

    def frame(self, slot, f):
        c, sp = self.ctxs[slot], self.streams[slot].cuda_stream
        self.cam.transform = self.poses[f % len(self.poses)]
        c.rng_set_frame(f, WARMUP_RNG)
        self.capi.launch_renderer(self.tree, self.cam, self.opt, c, stream=sp)
        if self.net:
            self.net.denoise(self.cam, c, stream=sp)

    # ---- device-resident throughput: CUDA events around reps x K frames issued round-robin on n_pipe slots; every frame is
    #      one rto_frame graph launch (render -> net -> filter, no read-back) unless graph=False (three separate launches)
    def pipelined(self, my_frames, n_pipe, warm, min_s, barrier, graph=True):
        torch, K = self.torch, len(my_frames)
        if graph:
            key = ("none", n_pipe, True)
            if key not in self.frames_cache:
                self.frames_cache[key] = {"bufs": [], "frames": [self.capi.Frame(self.ctxs[k], self.tree, self.net, self.opt, self.cam.fx, self.cam.fy)
                                                                  for k in range(n_pipe)]}
            graphs = self.frames_cache[key]["frames"]
            plain_frame = self.frame

            pose_c = [self.capi.Frame.pose_array(p) for p in self.poses]

            def gframe(slot, f):
                graphs[slot].launch_indexed(pose_c[f % len(pose_c)], f, WARMUP_RNG, self.streams[slot].cuda_stream)

            self.frame = gframe
            try:
                return self.pipelined(my_frames, n_pipe, warm, min_s, barrier, graph=False)
            finally:
                self.frame = plain_frame

        def block(reps):
            e0 = torch.cuda.Event(enable_timing=True)
            ends = [torch.cuda.Event(enable_timing=True) for _ in range(n_pipe)]
            torch.cuda.synchronize()
            barrier()
            e0.record(self.streams[0])
            for j in range(1, n_pipe):
                self.streams[j].wait_event(e0)
            i = 0
            for _ in range(reps):
                for f in my_frames:
                    self.frame(i % n_pipe, f)
                    i += 1
            for j in range(n_pipe):
                ends[j].record(self.streams[j])
            torch.cuda.synchronize()
            return max(e0.elapsed_time(e) for e in ends)

        for i in range(warm):
            self.frame(i % n_pipe, my_frames[i % K])
        est = block(1)
        reps = max(1, min(4000, int(math.ceil(min_s * 1e3 / max(est, 1e-3)))))
        l0 = self.capi.launch_count()
        ms = block(reps)
        return {"ms_total": ms, "reps": reps, "launches": self.capi.launch_count() - l0}

    # ---- the reference's Timer protocol: one stream, stage events, host sync per frame (main_headless.cpp:481-506)
    def serial_protocol(self, my_frames, min_s):
        ctx, K = self.ctxs[0], len(my_frames)
        ctx.timer_enable(True)
        for f in my_frames[:3]:
            self.frame(0, f)
            ctx.timer_record(bool(self.net))
        ctx.timer_reset()
        n, t0 = 0, time.perf_counter()
        while True:
            for f in my_frames:
                self.frame(0, f)
                ctx.timer_record(bool(self.net))
            n += K
            wall = time.perf_counter() - t0
            if wall >= min_s or n >= 4000 * K:
                break
        ms, cnt = ctx.timer_report()
        ctx.timer_enable(False)
        return {"render_ms": ms[0], "net_ms": ms[1], "filter_ms": ms[2], "frames": cnt, "wall_fps": n / wall}

    # ---- end to end through the public API with host buffers: the pose comes from host memory, the result lands in pinned
    #      host memory every frame.  Ring of n_slots (context, stream, pinned buffer): the host blocks on the oldest slot only.
    def e2e(self, my_frames, n_slots, readback, warm, min_s, barrier, graph=True, sequence=True):
        """sequence (with graph frames): the K-frame block is ONE library call, rto_frame_sequence — the host loop
        volrend_headless --pipe runs; sequence=False issues one rto_frame_launch_indexed + stream wait per frame from Python."""
        capi, torch, K = self.capi, self.torch, len(my_frames)
        shape, dt = {"rgba8": ((self.h, self.w, 4), np.uint8), "float": ((self.h, self.w, 4), np.float32),
                     "aux": ((8, self.h, self.w), np.float32)}[readback]
        key = (readback, n_slots, graph)
        if key not in self.frames_cache:
            bufs = [capi.PinnedBuffer(shape, dt) for _ in range(n_slots)]
            frames = []
            if graph:
                for k in range(n_slots):
                    frames.append(capi.Frame(self.ctxs[k], self.tree, self.net, self.opt, self.cam.fx, self.cam.fy,
                                             **{{"rgba8": "rgba8", "float": "image", "aux": "aux"}[readback]: bufs[k]}))
            self.frames_cache[key] = {"bufs": bufs, "frames": frames}
        bufs, frames = self.frames_cache[key]["bufs"], self.frames_cache[key]["frames"]
        host_poses = np.ascontiguousarray(self.poses)            # pageable host memory, read per step
        pose_c = [capi.Frame.pose_array(p) for p in host_poses]  # ... as the ctypes arrays the frame launch takes

        def one(i, f):
            k = i % n_slots
            c, st = self.ctxs[k], self.streams[k]
            st.synchronize()                                     # slot k is free again (its previous copy has landed)
            if graph:
                frames[k].launch_indexed(pose_c[f % len(pose_c)], f, WARMUP_RNG, st.cuda_stream)   # rng for pose f + ONE graph launch
                return
            c.rng_set_frame(f, WARMUP_RNG)
            self.cam.transform = host_poses[f % len(host_poses)]
            capi.launch_renderer(self.tree, self.cam, self.opt, c, stream=st.cuda_stream)
            if self.net:
                self.net.denoise(self.cam, c, stream=st.cuda_stream)
            if readback == "rgba8":
                c.read_image_rgba8(bufs[k].array, stream=st.cuda_stream, sync=False)
            elif readback == "float":
                c.read_image(bufs[k].array, stream=st.cuda_stream, sync=False)
            else:
                c.read_aux(bufs[k].array, stream=st.cuda_stream, sync=False)

        seq = None
        if graph and sequence and all(b - a == 1 for a, b in zip(my_frames, my_frames[1:])):
            seq = capi.FrameSequence(frames, [self.streams[k].cuda_stream for k in range(n_slots)], host_poses, WARMUP_RNG)

        def block(reps):
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            i = 0
            for _ in range(reps):
                if seq is not None:
                    seq.run(my_frames[0], K)     # K frames: pose + rng per frame, slot waits, graph launches, all inside the library
                    i += K
                    continue
                for f in my_frames:
                    one(i, f)
                    i += 1
            torch.cuda.synchronize()
            return time.perf_counter() - t0, i

        for i in range(max(warm, n_slots)):
            one(i, my_frames[i % K])
        est, _ = block(1)
        reps = max(1, min(4000, int(math.ceil(min_s / max(est, 1e-6)))))
        sec, n = block(reps)
        last = bufs[(my_frames[-1] if seq is not None else n - 1) % n_slots].array
        chk = int(last.sum(dtype=np.int64)) if readback == "rgba8" else float(last.sum(dtype=np.float64))
        return {"seconds": sec, "frames": n, "reps": reps, "checksum": chk, "bytes": int(np.prod(shape)) * np.dtype(dt).itemsize}


def run_cuda_arm(args):
    import torch
    import torch.distributed as dist

    from rt_octree_b200 import capi, synthetic as S

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    capi.set_device(local)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_group = dist.new_group(backend="gloo")   # host-side rendezvous that neither spins a CPU core nor parks a kernel on the GPUs

    def barrier():
        if world > 1:
            dist.barrier()

    def host_barrier():
        if world > 1:
            dist.barrier(group=host_group)

    def reduce_max(vals):
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    tree = load_tree(rank, barrier)
    poses, fx = workload_poses()
    weights = S.make_guidance_weights(0)
    K, Wm = args.steps, max(args.warmup, 3)
    min_s = args.min_seconds
    # frame sharding: rank r renders global frames r*K .. r*K+K-1 (weak scaling); rng is a pure function of the frame
    my_frames = [rank * K + i for i in range(K)]
    n_pipe = 1 if args.serial else args.pipe
    NSLOT = max(args.pipe, 4) + 4   # ring of the read-back loops: the host blocks on the oldest slot every frame, so a deeper
                                    # ring keeps four frames queued on the GPU (measured 4/6/8 slots: 4948/5344/5374 frames/s)
    rig = Rig(capi, torch, tree, weights, W, H, fx, SPP, True, poses, NSLOT)
    info = rig.tree.info

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    t_clk0 = clocks.mark()
    pl = rig.pipelined(my_frames, n_pipe, Wm, min_s, barrier, graph=not args.no_graph)
    t_clk1 = clocks.mark()
    clk = clocks.stop(t_clk0, t_clk1) if rank == 0 else None   # the sampler covers the timed region of `value` and is gone before
    barrier()                                                   # the host-side (e2e) measurements: nvidia-smi polling takes driver locks
    sp = rig.serial_protocol(my_frames, min_s)

    # ---- cold-L2 variant: flush L2 (256 MB write) before every frame, per-frame events
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cold = []
    with torch.cuda.stream(rig.streams[0]):
        for f in my_frames[: min(K, 30)]:
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rig.frame(0, f)
            b.record()
            cold.append((a, b))
    torch.cuda.synchronize()
    cold_ms = float(np.mean([a.elapsed_time(b) for a, b in cold]))
    del flush

    e8 = rig.e2e(my_frames, NSLOT, "rgba8", Wm, min_s, barrier, graph=not args.no_graph)
    e8py = rig.e2e(my_frames, NSLOT, "rgba8", Wm, min_s, barrier, graph=not args.no_graph, sequence=False)
    ef = rig.e2e(my_frames, max(args.pipe, 3), "float", Wm, min_s, barrier, graph=not args.no_graph)

    ms_per_frame, e2e8_ms, e2ef_ms, cold_ms, render_ms, net_ms, filter_ms, e2e8py_ms = reduce_max(
        [pl["ms_total"] / (pl["reps"] * K), 1e3 * e8["seconds"] / e8["frames"], 1e3 * ef["seconds"] / ef["frames"], cold_ms,
         sp["render_ms"], sp["net_ms"], sp["filter_ms"], 1e3 * e8py["seconds"] / e8py["frames"]])

    extras = {}
    if world == 1 and not args.no_extras:
        # ---- the --write_buffer path of config 3: the 20.48 MB guidance buffer copied to the host every frame
        ea = rig.e2e(my_frames, max(args.pipe, 3), "aux", Wm, min_s, barrier, graph=not args.no_graph)
        extras["e2e_aux_write_buffer"] = {
            "value": ea["frames"] / ea["seconds"], "unit": "frames/s", "d2h_bytes_per_step": ea["bytes"], "reps": ea["reps"],
            "checksum": ea["checksum"], "note": "same frames, aux [8][H][W] fp32 read back (main_headless.cpp:512-523): PCIe-bound"}
    bytes_frame = counters = None
    if rank == 0:
        bytes_frame, counters = algorithmic_bytes(capi, rig.tree, rig.ctxs[0], rig.cam, rig.opt, poses, my_frames[: min(K, 8)], W, H)

    if world == 1 and not args.no_extras:
        # ---- BASELINE config 2: same octree, SPP 1, denoiser off (the image comes straight out of the render kernel)
        rig2 = Rig(capi, torch, tree, weights, W, H, fx, 1, False, poses, NSLOT)
        p2 = rig2.pipelined(my_frames, n_pipe, Wm, min_s, barrier, graph=not args.no_graph)
        s2 = rig2.serial_protocol(my_frames, min_s)
        x2 = rig2.e2e(my_frames, NSLOT, "rgba8", Wm, min_s, barrier, graph=not args.no_graph)
        b2, c2 = algorithmic_bytes(capi, rig2.tree, rig2.ctxs[0], rig2.cam, rig2.opt, poses, my_frames[: min(K, 8)], W, H)
        extras["config2_spp1_no_denoise"] = {
            "workload": "lego-synthetic depth9 800x800 spp1, denoiser off", "value": 1e3 * p2["reps"] * K / p2["ms_total"],
            "unit": "frames/s", "streams": n_pipe, "reps": p2["reps"], "value_reference_protocol": 1e3 / s2["render_ms"],
            "render_ms": s2["render_ms"], "serial_wall_fps": s2["wall_fps"], "e2e": x2["frames"] / x2["seconds"],
            "e2e_d2h_bytes_per_step": x2["bytes"], "algorithmic_bytes_per_launch": b2, "per_frame": c2,
            "roofline_achieved_gbs": b2 / (s2["render_ms"] * 1e-3) / 1e9}
        rig2.close()
        del rig2
    rig_tt = None
    if world == 1 and not args.no_extras and not args.no_tt:
        # ---- BASELINE config 4: Tanks-and-Temples-shaped depth-10 tree at 1920x1080, SPP 6 + denoise (one GPU's share of
        #      the frame-sharded job; frame sharding adds no per-frame work)
        tt_tree = load_tree(rank, None, TT_TREE_KW, "tt")
        tt_poses = S.poses_to_c2w12(S.make_poses(N_POSES, **TT_POSES_KW))
        Ktt = min(K, 50)
        ttf = my_frames[:Ktt]
        rig_tt = Rig(capi, torch, tt_tree, weights, TT_W, TT_H, TT_FX, SPP, True, tt_poses, max(args.pipe, 4))
        p4 = rig_tt.pipelined(ttf, n_pipe, Wm, min_s, barrier, graph=not args.no_graph)
        s4 = rig_tt.serial_protocol(ttf, min_s)
        x4 = rig_tt.e2e(ttf, max(args.pipe, 4), "rgba8", Wm, min_s, barrier, graph=not args.no_graph)
        b4, c4 = algorithmic_bytes(capi, rig_tt.tree, rig_tt.ctxs[0], rig_tt.cam, rig_tt.opt, tt_poses, ttf[: min(Ktt, 4)], TT_W, TT_H)
        i4 = rig_tt.tree.info
        extras["config4_tt_1080p"] = {
            "workload": "T&T-shaped synthetic depth10 (anisotropic) 1920x1080 spp6 denoise", "tree": dict(TT_TREE_KW, nodes=int(i4.capacity), leaves=int(i4.n_leaves)),
            "value": 1e3 * p4["reps"] * Ktt / p4["ms_total"], "unit": "frames/s", "streams": n_pipe, "reps": p4["reps"], "frames_per_rep": Ktt,
            "value_reference_protocol": 1e3 / (s4["render_ms"] + s4["net_ms"] + s4["filter_ms"]),
            "stage_ms": {"render": s4["render_ms"], "net": s4["net_ms"], "filter": s4["filter_ms"]}, "serial_wall_fps": s4["wall_fps"],
            "e2e": x4["frames"] / x4["seconds"], "e2e_d2h_bytes_per_step": x4["bytes"], "msamples_per_s": 1e3 * p4["reps"] * Ktt / p4["ms_total"] * TT_W * TT_H * SPP / 1e6,
            "algorithmic_bytes_per_launch": b4, "per_frame": c4, "roofline_achieved_gbs": b4 / (s4["render_ms"] * 1e-3) / 1e9}
        rig_tt.close()
        del rig_tt
    if world > 1 and not args.no_extras:
        # ---- BASELINE config 5: single-frame latency, 3840x2160 SPP 6 + denoise split into row bands over the ranks, every
        #      rank's filter epilogue storing its band straight into rank 0's image over NVLink (rt_octree_b200/sharding.py)
        try:
            from rt_octree_b200 import sharding as SH

            extras["config5_4k_tile_split"] = SH.bench_tile_split(capi, torch, dist, tree, weights, poses, rank, world, local,
                                                                   frames=30, warmup_rng=WARMUP_RNG)
        except Exception as e:  # a side measurement must not take the headline down
            extras["config5_4k_tile_split"] = {"unavailable": repr(e)[:300]}

    if not args.no_extras and not args.no_cli:
        # ---- the PRODUCT's own end to end: volrend_headless --pipe 8 --readback rgba8 on the same workload read from disk
        #      (tree.npz, transforms json), frame-sharded over all `world` GPUs by the C++ driver itself (one host thread per
        #      GPU).  Rank 0 runs it while the other ranks wait in a HOST-side (gloo) barrier: an NCCL barrier would park a
        #      polling kernel on every GPU and a spinning thread on every core the C++ driver needs.
        torch.cuda.synchronize()
        host_barrier()
        if rank == 0:
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import cli_bench

                files = cli_bench.workload_files(os.path.join(CACHE, "cli"))
                flags = ["--pipe", "8", "--readback", "rgba8"] + (["--num_gpus", str(world)] if world > 1 else [])
                r = cli_bench.run_cli(files, flags, 600 * world)
                extras["e2e_cli_pipe8"] = {"value": r.get("aggregate_wall_fps", r.get("fps")), "unit": "frames/s", "n_gpus": world,
                                           "command": "volrend_headless tree.npz transforms_test.json --options opt.json --ts_module w "
                                                      + " ".join(flags), "d2h_bytes_per_step": W * H * 4,
                                           "note": "wall clock of the C++ driver's timed loop (slowest shard), RGBA8 frames into pinned host memory"}
                if world == 1:
                    r1 = cli_bench.run_cli(files, [], 400)
                    extras["cli_serial_protocol"] = {"fps_stage_sum": r1.get("fps"), "fps_wall": r1.get("wall_fps"),
                                                     "command": "volrend_headless ... (no --pipe: the reference's one-stream protocol)"}
            except Exception as e:
                extras["e2e_cli_pipe8"] = {"unavailable": repr(e)[:300]}
        host_barrier()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = traffic_src = None   # DRAM bytes per launch of the render kernel: from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "render_traffic.json")))
            traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source", "profiles/render_traffic.json")
        except Exception:
            pass
        achieved = bytes_frame / (render_ms * 1e-3) / 1e9
        fps = world * 1e3 / ms_per_frame
        fps_protocol = world * 1e3 / (render_ms + net_ms + filter_ms)
        cfg = base_config(world, K)    # identical to the reference arm's `config` (same workload, same keys)
        run_info = {"streams": n_pipe, "tree_built": dict(nodes=int(info.capacity), leaves=int(info.n_leaves), max_depth=int(info.max_depth),
                                                          node_bytes=int(info.node_bytes), payload_bytes=int(info.payload_bytes),
                                                          grid_bytes=int(info.grid_bytes)),
                    "tree_load_s": rig.load_s, "min_seconds_per_measurement": min_s}
        line = {
            "metric": "fps_800x800_spp6_denoise", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_per_frame, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 traversal/shade, f16 GuidanceNet", "data": "synthetic",
            "reps": pl["reps"], "timed_region_s": pl["ms_total"] * 1e-3,
            "config": cfg,
            "run_info": run_info,
            "value_protocol": "%d frames in flight on %d (context, stream) pairs, %s, device-timed over reps x steps frames"
                              % (n_pipe, n_pipe, "three launches per frame" if args.no_graph else "one rto_frame graph launch per frame"),
            "value_reference_protocol": fps_protocol,
            "reference_protocol": {"definition": "1000 / (render + net + filter ms): one stream, cudaEvents per stage, host sync per frame "
                                                 "(Timer::report, render_context.hpp:190-206)", "frames": sp["frames"],
                                   "stage_ms": {"render": render_ms, "net": net_ms, "filter": filter_ms},
                                   "wall_fps_incl_host_gaps": world * sp["wall_fps"]},
            "msamples_per_s": fps * W * H * SPP / 1e6,
            "value_l2_flushed": world * 1e3 / cold_ms,
            "stage_ms": {"render": render_ms, "denoise": net_ms + filter_ms},
            "e2e": {"value": world * 1e3 / e2e8_ms, "unit": "frames/s", "h2d_bytes_per_step": 48 + 28,
                    "d2h_bytes_per_step": W * H * 4, "checksum": e8["checksum"], "frame_slots": NSLOT, "reps": e8["reps"],
                    "timed_region_s": e8["seconds"], "one_graph_launch_per_frame": not args.no_graph,
                    "api": "rto_frame_sequence: one library call per %d-frame block (pose + rng per frame, slot waits and graph launches "
                           "inside the library: the host loop of volrend_headless --pipe)" % K if not args.no_graph else "separate launches per frame",
                    "value_per_frame_calls": world * 1e3 / e2e8py_ms,   # the same ring driven frame by frame from Python (rto_frame_launch_indexed)
                    "readback": "RGBA8 written by the filter epilogue (rto_frame / rto_context_read_image_rgba8), the bytes volrend_headless -o "
                                "writes to the PNG; the reference converts the same values on the host (main_headless.cpp:524-541)"},
            "e2e_f32": {"value": world * 1e3 / e2ef_ms, "unit": "frames/s", "d2h_bytes_per_step": W * H * 16,
                        "checksum": ef["checksum"], "reps": ef["reps"],
                        "note": "same loop, float4 image read back (rto_context_read_image, the reference CLI's 10.24 MB copy): "
                                "bound by the PCIe link"},
            "gpu_launches": int(pl["launches"]),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": render_kernel_name(info), "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": bytes_frame, "per_frame": counters,
                         "kernel_ms": render_ms,
                         "note": "algorithmic bytes = reference-equivalent traffic (4*depth+2 per step, 54 per collided leaf, 32 aux per ray); "
                                 "the kernel is latency/issue-bound, its physical DRAM traffic is `traffic` (see profiles/ and DESIGN.md §4.1)"},
            "configs": extras,
        }
        if world == 1 and not args.no_baselines:
            threads = os.cpu_count() or 1
            per, kind, br = cpu_frame_seconds(tree, poses, fx, weights, args.cpu_frames, threads)
            line["cpu_baseline"] = {"value": 1.0 / per, "unit": "frames/s", "cores": threads, "kind": kind,
                                    "sample": "%d full 800x800 SPP6 frames" % args.cpu_frames, "breakdown_s": br}
            rc = reference_cuda_fps(tree, poses, fx, weights)
            if "fps" in rc:
                rc["speedup_same_protocol"] = fps_protocol / rc["fps"]
                rc["speedup_pipelined_value"] = fps / rc["fps"]
            line["reference_cuda"] = rc
        print(json.dumps(line))
    rig.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=3, help="frames timed for the cpu_baseline sample")
    ap.add_argument("--pipe", type=int, default=4, help="frames in flight (contexts/streams) of the device-timed loop")
    ap.add_argument("--serial", action="store_true", help="one stream, frames strictly back to back, for `value` too")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="minimum duration of every timed region (the K-frame block is repeated)")
    ap.add_argument("--no-graph", action="store_true", help="issue separate launches instead of one rto_frame graph launch per frame")
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu_baseline / reference_cuda side measurements")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs (SPP 1, T&T 1080p, write_buffer, tile split)")
    ap.add_argument("--no-tt", action="store_true", help="skip BASELINE config 4 (saves the depth-10 tree generation)")
    ap.add_argument("--no-cli", action="store_true", help="skip the volrend_headless --pipe measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
