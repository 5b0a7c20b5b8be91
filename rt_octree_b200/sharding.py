"""Multi-GPU host logic.  The reference is single-GPU; two ways to use N GPUs are provided:

* frame sharding (throughput): the pose list is split contiguously over ranks; the octree and GuidanceNet are
  replicated in each GPU's HBM; NO collective is needed because frames are independent and the RNG is a pure function
  of the global frame index (rto_context_rng_set_frame).
* tile split (single-frame latency): the image is cut into row bands; each rank renders its band plus a halo of
  6 rows (2 for the two 3x3 convolutions + 4 for the 9x9 filter level), denoises its band locally, and ONE gather of the
  final RGBA bands over NCCL (NVLink 5 / NVSwitch) assembles the frame on rank 0.  No halo exchange, no reduction.

One process per GPU (torchrun); torch.distributed is the plumbing (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

DENOISE_HALO = 6   # rows: 2 (conv1+conv2, 3x3 'same') + 4 (largest filter support, levels = 4)


def shard_frames(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the pose list for `rank` (same split as volrend_headless --num_gpus)."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def tile_bands(height: int, world: int) -> List[Tuple[int, int]]:
    """Row bands [y0, y1) per rank; band sizes differ by at most one row."""
    return [(height * r // world, height * (r + 1) // world) for r in range(world)]


def render_rows_for_band(band: Tuple[int, int], height: int, denoise: bool) -> Tuple[int, int]:
    """Rows a rank must RENDER so that it can denoise `band` without talking to its neighbours."""
    h = DENOISE_HALO if denoise else 0
    return max(0, band[0] - h), min(height, band[1] + h)


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
    """Single-frame latency mode on this rank: render band+halo, denoise the band, gather on rank 0.
    `image_tensor` is a torch view [H, W, 4] of ctx's image buffer (caller wraps rto_context_image)."""
    bands = tile_bands(cam.height, world)
    band = bands[rank]
    y0, y1 = render_rows_for_band(band, cam.height, opt.denoise)
    ctx.rng_set_frame(frame, warmup)
    capi.launch_renderer(tree, cam, opt, ctx, stream=stream, rect=(0, y0, cam.width, y1))
    if opt.denoise:
        net.denoise(cam, ctx, stream=stream, rows=band)
    capi.synchronize(stream)
    return gather_bands(image_tensor[band[0]:band[1]], bands, rank, world)
