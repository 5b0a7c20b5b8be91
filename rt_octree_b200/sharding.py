"""Multi-GPU host logic.  The reference is single-GPU; two ways to use N GPUs are provided:

* frame sharding (throughput): the pose list is split contiguously over ranks; the octree and GuidanceNet are
  replicated in each GPU's HBM; NO collective is needed because frames are independent and the RNG is a pure function
  of the global frame index (rto_context_rng_set_frame).
* tile split (single-frame latency): the image is cut into row bands; each rank renders its band plus a halo of
  2 + L rows (2 for the two 3x3 convolutions + L for the largest filter support, L = the net's levels; 6 for the shipped
  net), denoises its band locally, and the kernel that PRODUCES the final image (the filter epilogue) stores the band
  straight into rank 0's image through a peer mapping (rto_context_set_image_target: CUDA IPC handle between processes,
  cudaDeviceEnablePeerAccess inside one process) — peer-direct stores over NVLink 5 / NVSwitch fused with the last kernel.
  The only communication left is ONE tiny stream-ordered all-reduce that tells rank 0 every band has landed.  No halo
  exchange, no reduction of pixel data, no gather.  (`gather_bands` — one NCCL gather of the bands — is kept as the
  comparison path and for backends without peer access.)

One process per GPU (torchrun); torch.distributed is the plumbing (nccl on GPUs, gloo in the CPU tests).  The same split
inside ONE process (one host thread per GPU) is `volrend_headless --tile_split`.
"""
from __future__ import annotations

import time
from typing import List, Tuple

CONV_HALO = 2      # rows: conv1 + conv2, 3x3 'same'
DENOISE_HALO = 6   # rows for the shipped net: CONV_HALO + 4 filter levels (see denoise_halo)


def denoise_halo(levels: int = 4) -> int:
    """Rows of rendered halo a band needs so that it can be denoised without talking to its neighbours: the filter at row y
    reads guidance rows y-L..y+L (support of the largest level), each of which needs aux rows +-2 (two 3x3 convolutions)."""
    return CONV_HALO + int(levels)


def shard_frames(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the pose list for `rank` (same split as volrend_headless --num_gpus)."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def tile_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Row bands [y0, y1) per rank; band sizes differ by at most one row."""
    return [(height * r // world, height * (r + 1) // world) for r in range(world)]


def render_rows_for_band(band: Tuple[int, int], height: int, denoise: bool, levels: int = 4) -> Tuple[int, int]:
    """Rows a rank must RENDER so that it can denoise `band` without talking to its neighbours."""
    h = denoise_halo(levels) if denoise else 0
    return max(0, band[0] - h), min(height, band[1] + h)


def rebalance_bands(bounds: List[int], band_ms: List[float], height: int, min_rows: int = 16) -> List[int]:
    """Band balancer of the tile split (same rule as volrend_headless --tile_split): cost per row constant inside a band,
    cumulative cost cut into equal parts, boundaries moved half way to those cuts.  bounds has len(band_ms) + 1 entries."""
    n = len(band_ms)
    cum = [0.0]
    for t in band_ms:
        cum.append(cum[-1] + max(1e-6, float(t)))
    out = list(bounds)
    for k in range(1, n):
        target = cum[n] * k / n
        g = 0
        while g + 1 < n and cum[g + 1] < target:
            g += 1
        frac = (target - cum[g]) / (cum[g + 1] - cum[g])
        y = bounds[g] + frac * (bounds[g + 1] - bounds[g])
        out[k] = int(round(0.5 * bounds[k] + 0.5 * y))
    for k in range(1, n):
        out[k] = max(out[k], out[k - 1] + min_rows)
    for k in range(n - 1, 0, -1):
        out[k] = min(out[k], out[k + 1] - min_rows)
    out[0], out[n] = 0, height
    return out


def gather_bands(band_tensor, bands: List[Tuple[int, int]], rank: int, world: int, dst: int = 0, group=None):
    """Gather the per-rank [rows_r, W, 4] final-image bands to `dst` and return the assembled [H, W, 4] frame there
    (None elsewhere).  Bands may differ in height by one row, so they are padded to the tallest band for the collective."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return band_tensor
    max_rows = max(b[1] - b[0] for b in bands)
    W = band_tensor.shape[1]
    send = band_tensor
    if band_tensor.shape[0] < max_rows:
        send = torch.zeros((max_rows, W, 4), dtype=band_tensor.dtype, device=band_tensor.device)
        send[: band_tensor.shape[0]] = band_tensor
    send = send.contiguous()
    if rank == dst:
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, recv, dst=dst, group=group)
        return torch.cat([recv[r][: bands[r][1] - bands[r][0]] for r in range(world)], dim=0)
    dist.gather(send, None, dst=dst, group=group)
    return None


def render_frame_tile_split(capi, tree, net, cam, opt, ctx, frame: int, rank: int, world: int, image_tensor, warmup: int = 100,
                            stream: int = 0):
    """Comparison path: render band+halo, denoise the band into this rank's OWN image, then one gather on rank 0.
    `image_tensor` is a torch view [H, W, 4] of ctx's image buffer (caller wraps rto_context_image)."""
    bands = tile_bands(cam.height, world)
    band = bands[rank]
    y0, y1 = render_rows_for_band(band, cam.height, opt.denoise, net.levels if net is not None else 4)
    ctx.rng_set_frame(frame, warmup)
    capi.launch_renderer(tree, cam, opt, ctx, stream=stream, rect=(0, y0, cam.width, y1))
    if opt.denoise:
        net.denoise(cam, ctx, stream=stream, rows=band)
    capi.synchronize(stream)
    return gather_bands(image_tensor[band[0]:band[1]], bands, rank, world)


class PeerTileSplit:
    """Single-frame tile split with peer-direct stores (one process per GPU).

    Rank 0 owns the destination image (its context's float4 image and RGBA8 copy); at construction it exports both
    allocations as CUDA IPC handles, the other ranks open them and point their contexts' image target there.  `render`
    then launches, on every rank, render(band + halo) -> GuidanceNet(band + L) -> filter(band), the filter storing the band
    into rank 0's memory, followed by one 4-byte all-reduce on the same stream (NCCL: stream-ordered, no host sync) — or,
    with a host-side backend (gloo), a stream synchronise + barrier."""

    def __init__(self, capi, dist, rank: int, world: int, width: int, height: int, levels: int = 4, rgba8: bool = True,
                 balance: bool = True):
        import torch

        self.capi, self.dist, self.rank, self.world = capi, dist, rank, world
        self.levels = levels
        self.height = height
        self.ctx = capi.RenderContext(width, height)
        self.bands = tile_bands(height, world)
        self.balance = balance and world > 1
        self._opened = []
        self.nccl = world > 1 and dist.get_backend() == "nccl"
        self.token = torch.zeros(1, dtype=torch.int32, device="cuda") if self.nccl else None
        if world > 1:
            handles = [None, None]
            if rank == 0:
                handles = [capi.ipc_export(self.ctx.image_ptr), capi.ipc_export(self.ctx.image_rgba8_ptr) if rgba8 else None]
            dist.broadcast_object_list(handles, src=0)
            if rank != 0:
                img = capi.ipc_open(handles[0])
                img8 = capi.ipc_open(handles[1]) if handles[1] is not None else None
                self._opened = [p for p in (img, img8) if p]
                self.ctx.set_image_target(img, img8)
        elif rgba8:
            self.ctx.image_rgba8_ptr   # allocate: the filter then writes the RGBA8 copy as well

    def render(self, tree, net, cam, opt, c2w12, frame: int, warmup: int = 100, stream: int = 0, band_done=None):
        """Enqueue this rank's share of the frame on `stream`; returns after the completion barrier has been ENQUEUED (nccl)
        or has completed (host backends).  On rank 0 the full frame is then in ctx's image / RGBA8 copy, in stream order."""
        capi, ctx = self.capi, self.ctx
        band = self.bands[self.rank]
        y0, y1 = render_rows_for_band(band, cam.height, opt.denoise, self.levels)
        cam.transform = c2w12
        ctx.rng_set_frame(frame, warmup)
        capi.launch_renderer(tree, cam, opt, ctx, stream=stream, rect=(0, y0, cam.width, y1))   # denoise off: y0, y1 == the band
        if opt.denoise:
            net.denoise(cam, ctx, stream=stream, rows=band)
        if band_done is not None:
            band_done.record()                         # torch event on the current stream: this rank's band is done here
        if self.world > 1:
            if self.nccl:
                self.dist.all_reduce(self.token)       # on torch's current stream: callers pass that stream as `stream`
            else:
                capi.synchronize(stream)
                self.dist.barrier()
        if self.rank == 0:
            ctx.mark_image_written(True)

    def report_band_time(self, ms: float):
        """Call after a frame has completed (outside any timed window) with this rank's device time for its band: every
        rank learns all band times (one small all-gather) and moves the boundaries for the next frame (rebalance_bands)."""
        if not self.balance:
            return
        import torch

        dev = "cuda" if self.nccl else "cpu"
        mine = torch.tensor([float(ms)], dtype=torch.float32, device=dev)
        allt = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allt, mine)
        bounds = [b[0] for b in self.bands] + [self.height]
        nb = rebalance_bands(bounds, [float(t.item()) for t in allt], self.height)
        self.bands = [(nb[r], nb[r + 1]) for r in range(self.world)]

    def close(self):
        if self.rank != 0:
            self.ctx.set_image_target(None, None)
        for p in self._opened:
            self.capi.ipc_close(p)
        self._opened = []
        self.ctx.close()


def bench_tile_split(capi, torch, dist, tree, weights, poses, rank, world, local, frames=30, warmup_rng=100, width=3840, height=2160,
                     spp=6, check=True):
    """BASELINE config 5: single-frame latency of a 3840x2160 SPP 6 + denoise frame split into row bands over `world` GPUs.
    Per frame: barrier, t0, every rank renders/denoises its band (peer-direct stores into rank 0), completion all-reduce,
    RGBA8 copy of the assembled frame to rank 0's pinned host memory, t1.  Returns the dict bench.py prints (rank 0) and
    verifies the assembled frame against a single-GPU render of the same frame, bit for bit."""
    import numpy as np

    from rt_octree_b200 import synthetic as S

    fx = float(np.float32(S.blender_focal(width)))
    t = capi.N3Tree(tree)
    net = capi.Denoiser(weights)
    cam = capi.Camera(width, height, fx, fx)
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = spp, True
    ts = PeerTileSplit(capi, dist, rank, world, width, height, net.levels)
    host8 = capi.PinnedBuffer((height, width, 4), np.uint8) if rank == 0 else None
    st = torch.cuda.current_stream()
    sp = st.cuda_stream
    lat, dev_ms = [], []
    for f in range(frames + 3):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        eb = torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ts.render(t, net, cam, opt, poses[f % len(poses)], f, warmup_rng, stream=sp, band_done=eb)
        e1.record(st)
        if rank == 0:
            ts.ctx.read_image_rgba8(host8.array, stream=sp, sync=False)
        st.synchronize()
        if f >= 3:
            lat.append(time.perf_counter() - t0)
            dev_ms.append(e0.elapsed_time(e1))
        ts.report_band_time(e0.elapsed_time(eb))       # outside the timed window: next frame's band boundaries
    identical = None
    if check:
        f = frames + 2
        if rank == 0:
            got = ts.ctx.read_image().copy()
            got8 = host8.array.copy()
            c1 = capi.RenderContext(width, height)
            c1.image_rgba8_ptr
            cam.transform = poses[f % len(poses)]
            c1.rng_set_frame(f, warmup_rng)
            capi.launch_renderer(t, cam, opt, c1)
            net.denoise(cam, c1)
            identical = bool(np.array_equal(c1.read_image(), got) and np.array_equal(c1.read_image_rgba8(), got8))
            c1.close()
    vals = torch.tensor([float(np.median(lat)), float(np.median(dev_ms))], device="cuda", dtype=torch.float64)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    # single-GPU latency of the same frame on rank 0, same protocol (for the speed-up)
    single_ms = None
    if rank == 0:
        c1 = capi.RenderContext(width, height)
        c1.image_rgba8_ptr
        one = []
        for f in range(8):
            cam.transform = poses[f % len(poses)]
            c1.rng_set_frame(f, warmup_rng)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            capi.launch_renderer(t, cam, opt, c1, stream=sp)
            net.denoise(cam, c1, stream=sp)
            c1.read_image_rgba8(host8.array, stream=sp, sync=False)
            st.synchronize()
            one.append(time.perf_counter() - t0)
        single_ms = float(np.median(one[2:]) * 1e3)
        c1.close()
    dist.barrier()
    ts.close()
    net.close()
    t.close()
    out = {"workload": "lego-synthetic depth9 %dx%d spp%d denoise, row bands over %d GPUs" % (width, height, spp, world),
           "metric": "single-frame latency, barrier -> assembled RGBA8 frame in rank 0's pinned host memory",
           "latency_ms": float(vals[0]) * 1e3, "device_ms_render_to_assembled": float(vals[1]), "frames": frames,
           "single_gpu_latency_ms": single_ms, "bit_identical_to_single_gpu": identical,
           "exchange": "filter epilogue stores each band into rank 0's image through a CUDA-IPC peer mapping (NVLink); one 4-byte "
                       "stream-ordered all-reduce signals completion; no gather", "halo_rows": denoise_halo(4),
           "bands": "balanced by the previous frame's band times" if ts.balance else "equal heights",
           "last_band_boundaries": [b[0] for b in ts.bands] + [height],
           "d2h_bytes_per_frame": width * height * 4}
    return out
