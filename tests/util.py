"""Shared helpers for the tests."""
import ctypes as C

import numpy as np

TRACE_KEYS = ("steps", "term", "src_bits", "t_bits", "leaf_hash", "depth_sum", "n_hits", "hit_leaf", "hit_cnt")


def encode_nodes(tree):
    """numpy statement of the SoA node-word encoding (rt_octree_b200/csrc/rto_tree.cu build_nodes_kernel)."""
    child = tree["child"].reshape(-1).astype(np.int64)
    cap = child.size // 8
    sig = np.ascontiguousarray(tree["data"].reshape(cap * 8, -1)[:, -1]).view(np.uint16).astype(np.uint32)
    node_of = np.arange(cap * 8) // 8
    return np.where(child == 0, 0x80000000 | sig, node_of + child).astype(np.uint32)


def tree_depth(tree):
    child = tree["child"].reshape(-1, 8)
    cur, d = np.array([0]), 0
    while cur.size:
        d += 1
        c = child[cur]
        nz = c != 0
        cur = (cur[:, None] + c)[nz]
    return d


def host_walk(lib, tree, c2w12, W, H, fx, fy, spp, rng, ndc=(-1.0, 0.0, 0.0), step_size=1e-4, sigma_thresh=1e-2,
              thresh=None, max_seq=0, pix_range=None, grid=False, byte_bricks=True, deferred=True, fused=True):
    """grid=True: march over the sparse brick grid (VERIFY build) instead of the ancestor-stack walker; with the defaults
    that is the PRODUCTION marcher (walk_grid_fused: byte plane, deferred hits, fused table indices).  fused=False selects
    the v9 loop (walk_grid), byte_bricks the plane it reads (4-byte leaf words = RTO_GRID8=0), deferred the collision path."""
    nodes = encode_nodes(tree)
    lib.host_ray_use_byte_bricks(1 if byte_bricks else 0)
    lib.host_ray_use_deferred_hits(1 if deferred else 0)   # collisions resolved through the leaf-id planes after the march
    lib.host_ray_use_fused_index(1 if (fused and byte_bricks and deferred) else 0)
    if grid:
        child = np.ascontiguousarray(tree["child"].reshape(-1), np.int32)
        data = np.ascontiguousarray(tree["data"].reshape(-1)).view(np.uint16)
        nb = C.c_int64(0)
        K = lib.host_ray_set_grid(child.ctypes.data, data.ctypes.data, int(tree["data_dim"]), child.size // 8, tree_depth(tree), C.byref(nb))
        assert K > 0, "grid not built for this tree depth"
    else:
        lib.host_ray_set_grid(None, None, 0, 0, 0, None)
    b, e = pix_range if pix_range else (0, W * H)
    n = e - b
    out = dict(steps=np.zeros(n, np.uint32), term=np.zeros(n, np.int32), src_bits=np.zeros(n, np.uint32),
               t_bits=np.zeros(n, np.uint32), leaf_hash=np.zeros(n, np.uint64), depth_sum=np.zeros(n, np.uint32),
               n_hits=np.zeros(n, np.uint32), n_loads=np.zeros(n, np.uint32), hit_leaf=np.zeros((n, spp), np.int32),
               hit_cnt=np.zeros((n, spp), np.uint32), leaf_seq=np.zeros((n, max(max_seq, 1)), np.int32))
    off = np.ascontiguousarray(tree["offset"], np.float32)
    sc = np.ascontiguousarray(tree["invradius3"], np.float32)
    c2w = np.ascontiguousarray(c2w12, np.float32)
    th = None if thresh is None else np.ascontiguousarray(thresh, np.float32)
    order = ("steps", "term", "src_bits", "t_bits", "leaf_hash", "depth_sum", "n_hits", "n_loads", "hit_leaf", "hit_cnt", "leaf_seq")
    rc = lib.host_ray_walk(nodes.ctypes.data, tree_depth(tree), c2w.ctypes.data, off.ctypes.data, sc.ctypes.data, fx, fy,
                           ndc[0], ndc[1], ndc[2], step_size, sigma_thresh, W, H, spp, rng[0], rng[1], b, e,
                           None if th is None else th.ctypes.data, *[out[k].ctypes.data for k in order], max_seq)
    assert rc == 0
    if max_seq == 0:
        out["leaf_seq"] = None
    return out


def psnr(a, b):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


class GpuTrace:
    """Device-side trace buffers (torch used only as the allocator) + the rto_trace POD."""

    def __init__(self, capi, n, spp, max_seq=0, thresh=True, marcher=0):
        import torch

        dev = "cuda"
        self.t = dict(steps=torch.zeros(n, dtype=torch.int32, device=dev), term=torch.zeros(n, dtype=torch.int32, device=dev),
                      src_bits=torch.zeros(n, dtype=torch.int32, device=dev), t_bits=torch.zeros(n, dtype=torch.int32, device=dev),
                      leaf_hash=torch.zeros(n, dtype=torch.int64, device=dev), depth_sum=torch.zeros(n, dtype=torch.int32, device=dev),
                      n_hits=torch.zeros(n, dtype=torch.int32, device=dev), n_loads=torch.zeros(n, dtype=torch.int32, device=dev),
                      hit_leaf=torch.zeros((n, spp), dtype=torch.int32, device=dev),
                      hit_cnt=torch.zeros((n, spp), dtype=torch.int32, device=dev))
        if max_seq > 0:
            self.t["leaf_seq"] = torch.zeros((n, max_seq), dtype=torch.int32, device=dev)
        if thresh:
            self.t["thresh"] = torch.zeros((n, spp), dtype=torch.float32, device=dev)
        pod = capi.TracePOD()
        for k, v in self.t.items():
            setattr(pod, k, v.data_ptr())
        pod.max_seq = max_seq
        pod.marcher = marcher   # 0 = tree walker, 1 = production brick-grid marcher
        self.pod = pod

    def host(self):
        import torch

        torch.cuda.synchronize()
        out = {}
        for k, v in self.t.items():
            a = v.cpu().numpy()
            if k in ("steps", "src_bits", "t_bits", "depth_sum", "n_hits", "n_loads", "hit_cnt"):
                a = a.view(np.uint32)
            if k == "leaf_hash":
                a = a.view(np.uint64)
            out[k] = a
        return out
