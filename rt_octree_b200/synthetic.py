"""Deterministic synthetic inputs for the RT-Octree hot path.

Nothing here is part of the render path: these are *input generators* in the reference's own file
formats, so that the same bytes can be fed to the reference binary, the oracle and the CUDA library.

* ``make_tree``      -- a PlenOctree in the svox/``tree.npz`` key set the reference loader reads
                        (renderer/src/n3tree.cpp:228-362): ``data_dim`` int64, ``data_format`` '<U' string,
                        ``invradius3`` f32[3], ``offset`` f32[3], ``child`` int32 [cap,2,2,2] (RELATIVE node
                        offsets, 0 = leaf: include/volrend/internal/n3tree_query.hpp:36-46) and ``data`` fp16
                        [cap,2,2,2,data_dim] with the layout [R:basis][G:basis][B:basis][sigma]
                        (include/volrend/cuda/rt_core.cuh:251,288-315).
* ``make_poses``     -- blender ``transforms_test.json`` content (renderer/main_headless.cpp:255-272).
* ``write_*``        -- writers for tree.npz / transforms json / tt pose directory / opt.json.
"""
from __future__ import annotations

import json
import os

import numpy as np

CAMERA_ANGLE_X = 0.6911112070083618  # nerf_synthetic value; fx = 0.5*W/tan(0.5*angle) = 1111.11 at W=800
BLENDER_RADIUS = 4.0311

# child index inside a node: index = (ix*N + iy)*N + iz  (n3tree_query.hpp:27-33)
_OFFS = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=np.int64)


def _sd_box(p, c, h):
    q = np.abs(p - np.asarray(c, np.float32)) - np.asarray(h, np.float32)
    outside = np.sqrt(np.sum(np.maximum(q, 0.0) ** 2, axis=-1))
    inside = np.minimum(np.max(q, axis=-1), 0.0)
    return outside + inside


def _sd_cyl(p, c, r, hh, axis):
    """Capped cylinder along `axis` centred at c, radius r, half-height hh."""
    d = p - np.asarray(c, np.float32)
    ax = d[..., axis]
    oth = [i for i in range(3) if i != axis]
    rad = np.sqrt(d[..., oth[0]] ** 2 + d[..., oth[1]] ** 2)
    q0 = rad - r
    q1 = np.abs(ax) - hh
    outside = np.sqrt(np.maximum(q0, 0.0) ** 2 + np.maximum(q1, 0.0) ** 2)
    inside = np.minimum(np.maximum(q0, q1), 0.0)
    return outside + inside


def bulldozer_sdf(p):
    """Signed distance (world units, object inside about [-0.65,0.65]^3) of a lego-bulldozer-like solid:
    a union of boxes (base plate, chassis, cabin, blade, arms, tracks), cylinders (wheels, exhaust) and a
    periodic field of studs on the plate and cabin roof."""
    p = p.astype(np.float32)
    d = _sd_box(p, (0.0, 0.0, -0.30), (0.62, 0.40, 0.03))             # base plate
    d = np.minimum(d, _sd_box(p, (0.0, 0.0, -0.12), (0.40, 0.24, 0.15)))   # chassis
    d = np.minimum(d, _sd_box(p, (-0.12, 0.0, 0.18), (0.20, 0.20, 0.16)))  # cabin
    d = np.minimum(d, _sd_box(p, (0.56, 0.0, -0.10), (0.025, 0.44, 0.16)))  # blade
    d = np.minimum(d, _sd_box(p, (0.42, 0.30, -0.10), (0.15, 0.02, 0.03)))  # arms
    d = np.minimum(d, _sd_box(p, (0.42, -0.30, -0.10), (0.15, 0.02, 0.03)))
    d = np.minimum(d, _sd_box(p, (0.0, 0.33, -0.20), (0.46, 0.06, 0.09)))   # tracks
    d = np.minimum(d, _sd_box(p, (0.0, -0.33, -0.20), (0.46, 0.06, 0.09)))
    for x in (-0.36, -0.12, 0.12, 0.36):                                     # wheels
        d = np.minimum(d, _sd_cyl(p, (x, 0.0, -0.20), 0.10, 0.42, 1))
    d = np.minimum(d, _sd_cyl(p, (0.22, 0.12, 0.22), 0.03, 0.20, 2))         # exhaust
    d = np.minimum(d, _sd_cyl(p, (-0.45, 0.0, 0.05), 0.015, 0.33, 2))        # antenna
    # periodic studs (pitch 0.08) on the base plate top (z=-0.27) and cabin roof (z=0.34)
    pitch = 0.08
    q = p.copy()
    q[..., 0] = (np.mod(p[..., 0] + 0.5 * pitch, pitch) - 0.5 * pitch)
    q[..., 1] = (np.mod(p[..., 1] + 0.5 * pitch, pitch) - 0.5 * pitch)
    studs_plate = _sd_cyl(q, (0.0, 0.0, -0.255), 0.024, 0.017, 2)
    studs_plate = np.maximum(studs_plate, _sd_box(p, (0.0, 0.0, -0.255), (0.60, 0.38, 0.05)))
    studs_roof = _sd_cyl(q, (0.0, 0.0, 0.355), 0.024, 0.017, 2)
    studs_roof = np.maximum(studs_roof, _sd_box(p, (-0.12, 0.0, 0.355), (0.19, 0.19, 0.05)))
    d = np.minimum(d, np.minimum(studs_plate, studs_roof))
    return d


def make_tree(depth: int = 9, shell: float = 0.05, halo: float = 0.0, seed: int = 0, basis_dim: int = 9,
              invradius3=(0.375, 0.375, 0.375), offset=(0.5, 0.5, 0.5), sdf=bulldozer_sdf,
              sigma_range=(5.0, 300.0), chunk: int = 1 << 21):
    """Build a PlenOctree refined to `depth` child look-ups around the solid's surface shell.

    Returns a dict with the npz key set.  Node order is breadth first, so every relative child offset is
    positive (as in svox).  Leaves inside the solid carry sigma ~ LogUniform(sigma_range) and SH coefficients
    (DC ~ N(0,1), higher bands ~ N(0,0.3)); leaves outside carry sigma = 0 (and random SH, which the
    renderer must never read).  Data of interior (non-leaf) entries is zero.
    """
    rng = np.random.default_rng(seed)
    inv = np.asarray(invradius3, np.float32)
    off = np.asarray(offset, np.float32)
    data_dim = 3 * basis_dim + 1

    level_coords = [np.zeros((1, 3), np.int64)]  # integer coords (at level l) of internal nodes of level l
    child_rel = []   # per level (n,8) int64: index of child node within next level, or -1 for a leaf
    inside = []      # per level (n,8) bool: leaf centre inside the solid
    for l in range(depth):
        c = level_coords[l]
        n = c.shape[0]
        cc = (2 * c[:, None, :] + _OFFS[None]).reshape(-1, 3)        # (n*8,3) coords at level l+1
        h = np.float32(1.0 / (1 << (l + 1)))
        d = np.empty(cc.shape[0], np.float32)
        for s in range(0, cc.shape[0], chunk):
            ctr = (cc[s:s + chunk].astype(np.float32) + 0.5) * h       # tree coords in [0,1]
            world = (ctr - off) / inv
            d[s:s + chunk] = sdf(world)
        half_diag = np.float32(0.8660254) * h / float(inv.min())       # cell half diagonal in world units
        # refine cells that touch the band  -shell <= d <= halo  (plus the cell's own extent)
        touches = (d < halo + half_diag) & (d > -(shell + half_diag))
        refine = touches & (l + 1 < depth)
        idx = np.full(cc.shape[0], -1, np.int64)
        idx[refine] = np.arange(int(refine.sum()))
        child_rel.append(idx.reshape(n, 8))
        inside.append((d < 0).reshape(n, 8))
        if l + 1 < depth:
            level_coords.append(cc[refine])

    counts = [c.shape[0] for c in level_coords]
    starts = np.concatenate([[0], np.cumsum(counts)])
    cap = int(starts[-1])
    child = np.zeros((cap, 8), np.int32)
    data = np.zeros((cap, 8, data_dim), np.float16)
    depth_of_node = np.zeros(cap, np.int32)
    for l in range(depth):
        n = counts[l]
        s0 = int(starts[l])
        rel = child_rel[l]
        is_node = rel >= 0
        node_ids = np.arange(s0, s0 + n, dtype=np.int64)[:, None]
        absolute = int(starts[l + 1]) + rel if l + 1 < depth else rel
        child[s0:s0 + n] = np.where(is_node, absolute - node_ids, 0).astype(np.int32)
        depth_of_node[s0:s0 + n] = l
        leaf = ~is_node
        nl = int(leaf.sum())
        vals = np.empty((nl, data_dim), np.float32)
        sh = rng.standard_normal((nl, 3, basis_dim), dtype=np.float32)
        sh[:, :, 1:] *= np.float32(0.3)
        vals[:, :data_dim - 1] = sh.reshape(nl, 3 * basis_dim)
        u = rng.random(nl, dtype=np.float32)
        sig = np.exp(np.log(sigma_range[0]) + u * (np.log(sigma_range[1]) - np.log(sigma_range[0])))
        vals[:, data_dim - 1] = np.where(inside[l][leaf], sig, 0.0)
        blk = data[s0:s0 + n]
        blk[leaf] = vals.astype(np.float16)
    return {
        "data_dim": np.int64(data_dim),
        "data_format": np.array("SH%d" % basis_dim),
        "invradius3": inv.copy(),
        "offset": off.copy(),
        "child": child.reshape(cap, 2, 2, 2),
        "data": data.reshape(cap, 2, 2, 2, data_dim),
        # svox also writes these; the reference loader ignores them (scripts/compress_octree.py:62-66)
        "parent_depth": np.stack([np.zeros(cap, np.int32), depth_of_node], axis=1),
        "n_internal": np.int64(cap),
        "depth_limit": np.int64(depth),
        "geom_resize_fact": np.float64(1.0),
        "n_free": np.int64(0),
    }


def tree_stats(tree) -> dict:
    child = tree["child"].reshape(-1, 8)
    cap = child.shape[0]
    n_leaf = int((child == 0).sum())
    sig = tree["data"].reshape(cap * 8, -1)[:, -1]
    occ = int(((child.reshape(-1) == 0) & (sig > 0)).sum())
    return {"nodes": cap, "leaves": n_leaf, "occupied_leaves": occ,
            "child_bytes": int(child.nbytes), "data_bytes": int(tree["data"].nbytes)}


def write_tree_npz(path: str, tree: dict, compressed: bool = False) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    (np.savez_compressed if compressed else np.savez)(path, **tree)


def look_at_pose(eye, target=(0.0, 0.0, 0.0), world_up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """4x4 c2w in the NeRF/blender convention: columns right, up, back (camera looks along -back), centre."""
    eye = np.asarray(eye, np.float64)
    back = eye - np.asarray(target, np.float64)
    back /= np.linalg.norm(back)
    right = np.cross(np.asarray(world_up, np.float64), back)
    right /= np.linalg.norm(right)
    up = np.cross(back, right)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, back, eye
    return m


def make_poses(n: int = 200, radius: float = BLENDER_RADIUS, elevation_deg: float = 30.0) -> np.ndarray:
    """n c2w matrices [n,4,4] float64 on a circle, looking at the origin."""
    el = np.deg2rad(elevation_deg)
    out = []
    for i in range(n):
        az = 2.0 * np.pi * i / n
        eye = radius * np.array([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)])
        out.append(look_at_pose(eye))
    return np.stack(out)


def poses_to_c2w12(poses: np.ndarray) -> np.ndarray:
    """[n,4,4] -> float32 [n,12] in the column-major 4x3 order Camera::_update uploads
    (renderer/src/camera.cpp:72-73): right(3), up(3), back(3), centre(3)."""
    p = np.asarray(poses)[:, :3, :4].astype(np.float32)      # [n,3,4] rows x cols
    return np.ascontiguousarray(p.transpose(0, 2, 1)).reshape(-1, 12)


def blender_focal(width: int, camera_angle_x: float = CAMERA_ANGLE_X) -> float:
    """fx = fy = 0.5f * width / tanf(0.5f * camera_angle_x) in fp32 (main_headless.cpp:257-258)."""
    a = np.float32(camera_angle_x)
    return float(np.float32(0.5) * np.float32(width) / np.tan(np.float32(0.5) * a, dtype=np.float32))


def write_blender_json(path: str, poses: np.ndarray, camera_angle_x: float = CAMERA_ANGLE_X) -> None:
    frames = [{"file_path": "./test/r_%d" % i, "rotation": 0.0,
               "transform_matrix": [[float(v) for v in row] for row in np.asarray(m)]}
              for i, m in enumerate(poses)]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump({"camera_angle_x": camera_angle_x, "frames": frames}, f, indent=1)


def write_tt_dir(root: str, poses: np.ndarray, fx: float, fy: float, cx: float, cy: float) -> str:
    """NSVF/T&T layout: <root>/intrinsics.txt and <root>/pose/<i>.txt (4x4, OpenCV convention)."""
    pose_dir = os.path.join(root, "pose")
    os.makedirs(pose_dir, exist_ok=True)
    with open(os.path.join(root, "intrinsics.txt"), "w") as f:
        f.write("%r 0.0 %r 0.0\n0.0 %r %r 0.0\n0.0 0.0 1.0 0.0\n0.0 0.0 0.0 1.0\n" % (fx, cx, fy, cy))
    flip = np.diag([1.0, -1.0, -1.0, 1.0])
    for i, m in enumerate(poses):
        np.savetxt(os.path.join(pose_dir, "%04d.txt" % i), np.asarray(m) @ flip, fmt="%.9g")
    return pose_dir


REFERENCE_OPT_JSON = {  # renderer/options/opt.json, verbatim values
    "background_brightness": 1.0, "denoise": True, "spp": 6, "enable_probe": False, "grid_max_depth": 4,
    "probe": [0.0, 0.0, 1.0], "probe_disp_size": 100, "show_grid": False, "sigma_thresh": 0.01,
    "step_size": 0.0001, "stop_thresh": 0.01,
}


def write_opt_json(path: str, **overrides) -> None:
    d = dict(REFERENCE_OPT_JSON)
    d.update(overrides)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        json.dump(d, f, indent=2)


def make_guidance_weights(seed: int = 0, in_ch: int = 8, mid_ch: int = 32, levels: int = 4,
                          branches: int = 5):
    """Random-init GuidanceNet in its DEPLOYED (re-parameterised) form, fp16:
    W1[mid,in,3,3], b1[mid], W2[2L,mid,3,3], b2[2L]   (denoiser/network.py:123-168).

    Uses only numpy so it runs anywhere; the distribution mimics nn.Conv2d's default init summed over
    `branches` 3x3 and 1x1 branches (network.py:134-141).  tests/golden holds weights exported from the
    reference's own GuidanceNet(8,32,5,2,4) -> GuidanceNetCompact for the parity fixtures."""
    rng = np.random.default_rng(seed)

    def block(cin, cout):
        w = np.zeros((cout, cin, 3, 3), np.float32)
        b = np.zeros(cout, np.float32)
        for _ in range(branches):
            k3 = 1.0 / np.sqrt(cin * 9)
            w += rng.uniform(-k3, k3, (cout, cin, 3, 3)).astype(np.float32)
            b += rng.uniform(-k3, k3, cout).astype(np.float32)
        for _ in range(branches):
            k1 = 1.0 / np.sqrt(cin)
            w[:, :, 1, 1] += rng.uniform(-k1, k1, (cout, cin)).astype(np.float32)
            b += rng.uniform(-k1, k1, cout).astype(np.float32)
        return w.astype(np.float16), b.astype(np.float16)

    w1, b1 = block(in_ch, mid_ch)
    w2, b2 = block(mid_ch, 2 * levels)
    return {"w1": w1, "b1": b1, "w2": w2, "b2": b2}
