"""The CUDA kernels' per-ray code (rt_octree_b200/csrc/rto_ray.cuh: integer-coordinate, ancestor-resume traversal
over the SoA node words) instantiated on the host by tests/host_ray_harness.cpp, against the oracle's root-restart
floating-point restatement of the reference.  Every trace field must be BIT-identical.  (The GPU run of the same
header is checked by tests/test_gpu_render.py.)"""
import numpy as np
import pytest

from util import TRACE_KEYS, host_walk


@pytest.mark.parametrize("spp", [1, 2, 3, 4, 6, 8, 16, 32])
def test_host_ray_bit_exact(oracle, host_ray_lib, mid_tree, poses8, spp):
    from rt_octree_b200 import synthetic as S

    W, H = 120, 90
    fx = S.blender_focal(W)
    for pi in (0, 5):
        rng = oracle.frame_rng(pi)
        o = oracle.render(mid_tree, poses8[pi], W, H, fx, fx, spp, rng, max_seq=48)
        h = host_walk(host_ray_lib, mid_tree, poses8[pi], W, H, fx, fx, spp, rng, max_seq=48)
        for k in TRACE_KEYS + ("leaf_seq",):
            assert np.array_equal(h[k], o[k]), (k, spp, pi)
        assert o["steps"].max() > 40 and o["n_hits"].max() >= 1
        # the ancestor-resume descent loads far fewer node words than the reference's root restart (+1 sigma load)
        assert h["n_loads"].sum() < 0.5 * (o["depth_sum"].sum() + o["steps"].sum())


def test_host_ray_ndc_and_anisotropic(oracle, host_ray_lib, poses8):
    """NDC warp (volrend.cu:36-56) + anisotropic scale, forward-facing camera."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=6, shell=1.0, halo=0.05, seed=5, invradius3=(0.45, 0.3, 0.5), offset=(0.5, 0.45, 0.55))
    W, H = 64, 48
    fx = 60.0
    ndc = (float(W), float(H), fx)
    pose = S.poses_to_c2w12(np.stack([S.look_at_pose((0.1, 0.05, 0.2), target=(0.0, 0.0, -1.0), world_up=(0, 1, 0))]))[0]
    rng = oracle.frame_rng(0)
    for spp in (1, 6):
        o = oracle.render(tree, pose, W, H, fx, fx, spp, rng, ndc=ndc, max_seq=32)
        h = host_walk(host_ray_lib, tree, pose, W, H, fx, fx, spp, rng, ndc=ndc, max_seq=32)
        for k in TRACE_KEYS + ("leaf_seq",):
            assert np.array_equal(h[k], o[k]), k
        assert o["steps"].sum() > 0


def test_host_ray_injected_thresholds(oracle, host_ray_lib, small_tree, poses8):
    """Threshold injection path used by the GPU tests (lg2.approx values come from the device)."""
    from rt_octree_b200 import synthetic as S

    W, H, spp = 40, 30, 6
    fx = S.blender_focal(W)
    th = np.sort(np.random.default_rng(0).exponential(1.0, (W * H, spp)).astype(np.float32), axis=1)
    rng = oracle.frame_rng(2)
    o = oracle.render(small_tree, poses8[2], W, H, fx, fx, spp, rng, thresh=th)
    h = host_walk(host_ray_lib, small_tree, poses8[2], W, H, fx, fx, spp, rng, thresh=th)
    for k in TRACE_KEYS:
        assert np.array_equal(h[k], o[k]), k


def test_empty_and_degenerate_trees(oracle, host_ray_lib, poses8):
    """A root-only tree (8 leaves), all empty and all dense."""
    for sigma in (0.0, 50.0):
        data = np.zeros((1, 2, 2, 2, 28), np.float16)
        data[..., -1] = sigma
        data[..., 0] = 0.5
        tree = {"data_dim": np.int64(28), "data_format": np.array("SH9"), "invradius3": np.full(3, 0.375, np.float32),
                "offset": np.full(3, 0.5, np.float32), "child": np.zeros((1, 2, 2, 2), np.int32), "data": data}
        W, H, spp = 24, 24, 4
        rng = oracle.frame_rng(0)
        o = oracle.render(tree, poses8[0], W, H, 60.0, 60.0, spp, rng)
        h = host_walk(host_ray_lib, tree, poses8[0], W, H, 60.0, 60.0, spp, rng)
        for k in TRACE_KEYS:
            assert np.array_equal(h[k], o[k]), k
        if sigma == 0.0:
            assert o["aux"][3].max() == 0.0 and np.all(o["aux"][:3] == 1.0)   # pure background
        else:
            assert o["aux"][3].max() == 1.0


@pytest.mark.parametrize("byte_bricks,deferred,fused", [(True, True, True), (True, True, False), (True, False, False), (False, False, False)],
                         ids=["fused_index(production)", "bytes+leaf_planes(v9)", "bytes", "words"])
@pytest.mark.parametrize("spp", [1, 6, 32])
def test_host_grid_walk_bit_exact(oracle, host_ray_lib, mid_tree, poses8, spp, byte_bricks, deferred, fused):
    """The sparse brick grid walker (rto_ray.cuh walk_grid: 1-2 loads per step, no descent) against the oracle; the VERIFY
    build also checks at every step that the grid's (depth, sigma) equal the tree's, and at every collision that the
    leaf-id plane names the leaf the root descent finds (term == -777 flags a mismatch)."""
    from rt_octree_b200 import synthetic as S

    W, H = 120, 90
    fx = S.blender_focal(W)
    for pi in (0, 5):
        rng = oracle.frame_rng(pi)
        o = oracle.render(mid_tree, poses8[pi], W, H, fx, fx, spp, rng, max_seq=48)
        h = host_walk(host_ray_lib, mid_tree, poses8[pi], W, H, fx, fx, spp, rng, max_seq=48, grid=True, byte_bricks=byte_bricks,
                      deferred=deferred, fused=fused)
        assert not (h["term"] == -777).any(), "grid (depth, sigma) disagrees with the tree"
        for k in TRACE_KEYS + ("leaf_seq",):
            assert np.array_equal(h[k], o[k]), (k, spp, pi)
        # 1-2 loads per step; the byte plane adds a third only in cells with non-zero sigma
        assert h["n_loads"].sum() <= (2.2 if byte_bricks else 2) * o["steps"].sum()


def test_host_grid_walk_other_depths(oracle, host_ray_lib, poses8):
    """Depth 4 (K = 1, the shallowest gridded tree), depth 6 with anisotropic scale + NDC."""
    from rt_octree_b200 import synthetic as S

    t4 = S.make_tree(depth=4, shell=1.0, halo=0.3, seed=9)
    W, H = 64, 48
    fx = S.blender_focal(W)
    rng = oracle.frame_rng(1)
    o = oracle.render(t4, poses8[1], W, H, fx, fx, 6, rng)
    h = host_walk(host_ray_lib, t4, poses8[1], W, H, fx, fx, 6, rng, grid=True)
    assert not (h["term"] == -777).any()
    for k in TRACE_KEYS:
        assert np.array_equal(h[k], o[k]), k
    tree = S.make_tree(depth=6, shell=1.0, halo=0.05, seed=5, invradius3=(0.45, 0.3, 0.5), offset=(0.5, 0.45, 0.55))
    fx = 60.0
    ndc = (float(W), float(H), fx)
    pose = S.poses_to_c2w12(np.stack([S.look_at_pose((0.1, 0.05, 0.2), target=(0.0, 0.0, -1.0), world_up=(0, 1, 0))]))[0]
    o = oracle.render(tree, pose, W, H, fx, fx, 6, rng, ndc=ndc)
    h = host_walk(host_ray_lib, tree, pose, W, H, fx, fx, 6, rng, ndc=ndc, grid=True)
    assert not (h["term"] == -777).any()
    for k in TRACE_KEYS:
        assert np.array_equal(h[k], o[k]), k


def _sphere_sdf(p):
    return np.sqrt((p.astype(np.float32) ** 2).sum(-1)) - np.float32(0.35)


@pytest.mark.parametrize("depth", [5, 7, 8, 9, 10, 11])
def test_host_fused_index_every_grid_level(oracle, host_ray_lib, poses8, depth):
    """The fused-index marcher (rto_ray.cuh FusedIdx: per-K magic adds, biased march table, sign trick on the z add for the K
    whose exponent sum leaves bit 31 clear) on trees of every remaining grid level K = depth - 3 (K = 1, 3 above): all trace
    fields and the visited-leaf sequence equal the oracle's, the grid agrees with the tree at every step (term != -777), and
    the v9 loop gives the same record."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=depth, shell=0.02, halo=0.04 if depth < 10 else 0.01, seed=depth, sdf=_sphere_sdf)
    W, H, spp = 48, 36, 6
    fx = S.blender_focal(W)
    rng = oracle.frame_rng(3)
    o = oracle.render(tree, poses8[3], W, H, fx, fx, spp, rng, max_seq=32)
    h = host_walk(host_ray_lib, tree, poses8[3], W, H, fx, fx, spp, rng, max_seq=32, grid=True)
    v9 = host_walk(host_ray_lib, tree, poses8[3], W, H, fx, fx, spp, rng, max_seq=32, grid=True, fused=False)
    assert not (h["term"] == -777).any(), "march table disagrees with the tree"
    for k in TRACE_KEYS + ("leaf_seq",):
        assert np.array_equal(h[k], o[k]), (k, depth)
        assert np.array_equal(h[k], v9[k]), (k, depth)
    assert o["n_hits"].max() >= 1 and o["depth_sum"].max() > 0
