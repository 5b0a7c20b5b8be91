#!/usr/bin/env python
"""bench.py — headline benchmark of the RT-Octree hot path on B200 (BASELINE.json: FPS @800x800 SPP6 + denoise).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation on the host cores

One "step" = one frame of the hot path: ray generation + octree traversal + SH shade + aux write (1 kernel) and the
GuidanceNet + kernel-filter denoiser, on a synthetic lego-shaped PlenOctree (depth 9, data_dim 28) at 800x800, SPP 6,
a different test pose every step.  Frames shard across ranks with no collective (weak scaling: K frames per rank).
Prints ONE JSON line (rank 0).  See DESIGN.md §6 for every field.
"""
from __future__ import annotations

import argparse
import json
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
CACHE = os.environ.get("RTO_CACHE", "/tmp/rto_cache")


# ------------------------------------------------------------------------------------------------ workload
def load_tree(rank=0, barrier=None):
    """Synthetic lego-shaped tree, generated once per box and cached as .npy under /tmp."""
    from rt_octree_b200 import synthetic as S

    tag = "lego_d%d_s%g_h%g_r%d" % (TREE_KW["depth"], TREE_KW["shell"], TREE_KW["halo"], TREE_KW["seed"])
    fc, fd = os.path.join(CACHE, tag + "_child.npy"), os.path.join(CACHE, tag + "_data.npy")
    if rank == 0 and not (os.path.exists(fc) and os.path.exists(fd)):
        os.makedirs(CACHE, exist_ok=True)
        t = S.make_tree(**TREE_KW)
        np.save(fc + ".tmp.npy", t["child"])
        np.save(fd + ".tmp.npy", t["data"])
        os.replace(fc + ".tmp.npy", fc)
        os.replace(fd + ".tmp.npy", fd)
    if barrier:
        barrier()
    tree = {"data_dim": np.int64(28), "data_format": np.array("SH9"), "invradius3": np.full(3, 0.375, np.float32),
            "offset": np.full(3, 0.5, np.float32), "child": np.load(fc, mmap_mode="r"), "data": np.load(fd, mmap_mode="r")}
    return tree


def workload_poses():
    from rt_octree_b200 import synthetic as S

    return S.poses_to_c2w12(S.make_poses(N_POSES)), float(np.float32(S.blender_focal(W)))


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
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.p.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
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
def cpu_frame_seconds(tree, poses, fx, weights, frames, nthreads, want_breakdown=False):
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
    for f in range(frames):
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
    warm = min(args.warmup, 1)
    per, kind, br = cpu_frame_seconds(tree, poses, fx, weights, steps, threads)
    fps = 1.0 / per
    line = {"impl": "reference", "metric": "fps_800x800_spp6_denoise", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "poses": N_POSES, "tree": TREE_KW},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                             "sample": "%d full 800x800 SPP6 frames (trace_ray host build + torch CPU GuidanceNet + filter)" % steps,
                             "breakdown_s": br},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ CUDA arm
def algorithmic_bytes(capi, tree_h, ctx, cam, opt, poses, frames):
    """per ray sum_steps(4*d_s + 2) + 54*n_leaf + 32 (aux) [+16 image when denoise is off], with
    d_s = child look-ups of the REFERENCE's root-restart query — exact counters from the trace kernel (untimed)."""
    import torch

    n = W * H
    steps = torch.zeros(n, dtype=torch.int32, device="cuda")
    dsum = torch.zeros(n, dtype=torch.int32, device="cuda")
    hits = torch.zeros(n, dtype=torch.int32, device="cuda")
    loads = torch.zeros(n, dtype=torch.int32, device="cuda")
    tr = capi.TracePOD()
    tr.steps, tr.depth_sum, tr.n_hits, tr.n_loads = steps.data_ptr(), dsum.data_ptr(), hits.data_ptr(), loads.data_ptr()
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    tree = load_tree(rank, barrier)
    poses, fx = workload_poses()
    weights = S.make_guidance_weights(0)
    t0 = time.perf_counter()
    tree_h = capi.N3Tree(tree)
    load_s = time.perf_counter() - t0
    info = tree_h.info
    net = capi.Denoiser(weights)
    cam = capi.Camera(W, H, fx, fx)
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = SPP, True
    K, Wm = args.steps, max(args.warmup, 3)
    # frame sharding: rank r renders global frames r*K .. r*K+K-1 (weak scaling); rng is a pure function of the frame
    my_frames = [rank * K + i for i in range(K)]
    NBUF = max(3, args.pipe)   # frame slots of the fp32 read-back loop (the device-timed loop uses the first --pipe)
    NBUF8 = NBUF + 4           # frame slots of the RGBA8 read-back loop: the host blocks on the oldest slot every frame, so
                               # a deeper ring keeps four frames queued on the GPU (measured 4/6/8 slots: 4948/5344/5374)
    ctxs = [capi.RenderContext(W, H) for _ in range(NBUF8)]
    streams = [torch.cuda.Stream() for _ in range(NBUF8)]
    ctx = ctxs[0]
    s0 = streams[0].cuda_stream

    def frame(c, f, stream_ptr):
        cam.transform = poses[f % len(poses)]
        c.rng_set_frame(f, WARMUP_RNG)
        capi.launch_renderer(tree_h, cam, opt, c, stream=stream_ptr)
        net.denoise(cam, c, stream=stream_ptr)

    # ---- device-resident throughput: K frames, CUDA events, max over ranks.  Frames are independent, so they are
    #      issued alternately on two (context, stream) pairs: the long tail of one frame's render kernel (a few heavy
    #      warps) overlaps the next frame's start.  --serial uses one stream (the reference's protocol).
    n_pipe = 1 if args.serial else args.pipe
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()     # before the warm-up so that nvidia-smi is already streaming when the timed region starts
    for i in range(Wm):
        frame(ctxs[i % n_pipe], my_frames[i % K], streams[i % n_pipe].cuda_stream)
    torch.cuda.synchronize()
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(n_pipe)]
    launches0 = capi.launch_count()
    torch.cuda.synchronize()
    e0.record(streams[0])
    for j in range(1, n_pipe):
        streams[j].wait_event(e0)
    for i, f in enumerate(my_frames):
        frame(ctxs[i % n_pipe], f, streams[i % n_pipe].cuda_stream)
    for j in range(n_pipe):
        ends[j].record(streams[j])
    torch.cuda.synchronize()
    launches = capi.launch_count() - launches0
    ms_total = max(e0.elapsed_time(e) for e in ends)
    barrier()

    # ---- per-kernel stage times (Timer: cudaEvents around each launch), same frames
    ctx.timer_enable(True)
    ctx.timer_reset()
    for f in my_frames[: min(K, 100)]:
        frame(ctx, f, s0)
        ctx.timer_record(True)
    stage_ms, _ = ctx.timer_report()
    ctx.timer_enable(False)

    # ---- cold-L2 variant: flush L2 (256 MB write) before every frame, per-frame events
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cold = []
    with torch.cuda.stream(streams[0]):
        for f in my_frames[: min(K, 30)]:
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            frame(ctx, f, s0)
            b.record()
            cold.append((a, b))
    torch.cuda.synchronize()
    cold_ms = float(np.mean([a.elapsed_time(b) for a, b in cold]))
    del flush

    # ---- end to end through the public API with host buffers: pose from host memory, final image read back into
    #      pinned host memory every frame; a ring of contexts/streams so frame f's D2H overlaps the next frames' kernels.
    #      Two read-backs of the same frames: the float4 image (what the reference CLI copies, 10.24 MB: PCIe-bound,
    #      reported as e2e_f32) and RGBA8 converted on the device (what `volrend_headless -o` copies for the PNG, 2.56 MB:
    #      the headline e2e).
    pinned = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(NBUF)]
    host_poses = np.ascontiguousarray(poses)            # pageable host memory, read per step

    def e2e_frame(i, f):
        c, st = ctxs[i % NBUF], streams[i % NBUF]
        st.synchronize()                                 # buffer i&1 is free again (its previous D2H finished)
        cam.transform = host_poses[f % len(poses)]
        c.rng_set_frame(f, WARMUP_RNG)
        capi.launch_renderer(tree_h, cam, opt, c, stream=st.cuda_stream)
        net.denoise(cam, c, stream=st.cuda_stream)
        c.read_image(pinned[i % NBUF].numpy(), stream=st.cuda_stream, sync=False)

    for i in range(Wm):
        e2e_frame(i, my_frames[i % K])
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i, f in enumerate(my_frames):
        e2e_frame(i, f)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # same loop with the RGBA8 read-back (the bytes the CLI writes to PNG; a quarter of the D2H traffic)
    pinned8 = [torch.empty((H, W, 4), dtype=torch.uint8).pin_memory() for _ in range(NBUF8)]

    def e2e8_frame(i, f):
        c, st = ctxs[i % NBUF8], streams[i % NBUF8]
        st.synchronize()
        cam.transform = host_poses[f % len(poses)]
        c.rng_set_frame(f, WARMUP_RNG)
        capi.launch_renderer(tree_h, cam, opt, c, stream=st.cuda_stream)
        net.denoise(cam, c, stream=st.cuda_stream)
        c.read_image_rgba8(pinned8[i % NBUF8].numpy(), stream=st.cuda_stream, sync=False)

    for i in range(max(Wm, NBUF8)):
        e2e8_frame(i, my_frames[i % K])
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i, f in enumerate(my_frames):
        e2e8_frame(i, f)
    torch.cuda.synchronize()
    e2e8_s = time.perf_counter() - t0
    # clocks / throttle reasons were sampled (nvidia-smi, 100 ms period) from the start of the device-timed loop to here:
    # the timed region itself can be shorter than one sampling period, the loops after it keep the GPU under the same load
    clk = clocks.stop() if rank == 0 else None
    checksum = float(pinned[(K - 1) % NBUF].sum())
    checksum8 = int(pinned8[(K - 1) % NBUF8].sum(dtype=torch.int64))

    # ---- reduce over ranks: max time
    tt = torch.tensor([ms_total, e2e_s * 1e3, cold_ms, stage_ms[0], stage_ms[1] + stage_ms[2], e2e8_s * 1e3], device="cuda",
                      dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, cold_ms, render_ms, denoise_ms, e2e8_ms = [float(v) for v in tt.cpu()]
    if rank == 0:
        bytes_frame, counters = algorithmic_bytes(capi, tree_h, ctx, cam, opt, poses, my_frames[: min(K, 8)])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        traffic = None   # DRAM bytes per launch of the render kernel from the committed ncu --set full capture
        try:
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", "render_traffic.json")))["dram_bytes_per_launch"])
        except Exception:
            pass
        achieved = bytes_frame / (render_ms * 1e-3) / 1e9
        fps = world * K / (ms_total * 1e-3)
        line = {
            "metric": "fps_800x800_spp6_denoise", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 traversal/shade, f16 GuidanceNet", "data": "synthetic",
            "config": {"workload": WORKLOAD, "poses": N_POSES, "frames_per_rank": K, "parallelism": "frame-sharded x%d" % world,
                       "streams": n_pipe,
                       "tree": dict(TREE_KW, nodes=int(info.capacity), leaves=int(info.n_leaves), max_depth=int(info.max_depth),
                                    node_bytes=int(info.node_bytes), payload_bytes=int(info.payload_bytes)),
                       "l2": "inputs larger than L2 (tree %.2f GB), a different pose every step, no flush; cold-L2 variant in value_l2_flushed"
                             % ((info.node_bytes + info.payload_bytes) / 1e9),
                       "tree_load_s": load_s},
            "msamples_per_s": fps * W * H * SPP / 1e6,
            "value_l2_flushed": world * 1e3 / cold_ms,
            "stage_ms": {"render": render_ms, "denoise": denoise_ms},
            "e2e": {"value": world * K / (e2e8_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 48 + 28,
                    "d2h_bytes_per_step": W * H * 4, "checksum": checksum8, "frame_slots": NBUF8,
                    "readback": "RGBA8 converted on the device (rto_context_read_image_rgba8), the bytes volrend_headless -o "
                                "writes to the PNG; the reference converts the same values on the host (main_headless.cpp:524-541)"},
            "e2e_f32": {"value": world * K / (e2e_ms * 1e-3), "unit": "frames/s", "d2h_bytes_per_step": W * H * 16,
                        "checksum": checksum, "frame_slots": NBUF,
                        "note": "same loop, float4 image read back (rto_context_read_image, the reference CLI's 10.24 MB copy): "
                                "bound by the PCIe link"},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "render_kernel<6>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": bytes_frame, "per_frame": counters,
                         "note": "algorithmic bytes = reference-equivalent traffic (4*depth+2 per step, 54 per collided leaf, 32 aux per ray)"},
        }
        if world == 1 and not args.no_baselines:
            threads = os.cpu_count() or 1
            per, kind, br = cpu_frame_seconds(tree, poses, fx, weights, args.cpu_frames, threads)
            line["cpu_baseline"] = {"value": 1.0 / per, "unit": "frames/s", "cores": threads, "kind": kind,
                                    "sample": "%d full 800x800 SPP6 frames" % args.cpu_frames, "breakdown_s": br}
            line["reference_cuda"] = reference_cuda_fps(tree, poses, fx, weights)
        print(json.dumps(line))
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
    ap.add_argument("--serial", action="store_true", help="one stream, frames strictly back to back (reference protocol)")
    ap.add_argument("--no-baselines", action="store_true", help="skip the cpu_baseline / reference_cuda side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
