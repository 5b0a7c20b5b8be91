"""GPU parity tests of the render path, through the C ABI (rt_octree_b200/capi.py -> librtoctree_b200.so):
traversal outputs BIT-EXACT against the oracle (thresholds dst[] are taken from the GPU because MUFU lg2.approx has
no CPU equivalent; everything downstream of them is compared bit for bit), aux/RGB within 1e-5 of the oracle."""
import numpy as np
import pytest

from util import TRACE_KEYS, GpuTrace

pytestmark = pytest.mark.gpu


def _setup(capi, tree, W, H, fx, fy=None):
    t = capi.N3Tree(tree)
    ctx = capi.RenderContext(W, H)
    cam = capi.Camera(W, H, fx, fy if fy else fx)
    return t, ctx, cam


def _opts(capi, spp, denoise=False, **kw):
    o = capi.RenderOptions()
    o.spp = spp
    o.denoise = denoise
    for k, v in kw.items():
        setattr(o, k, v)
    return o


@pytest.mark.parametrize("marcher", [0, 1], ids=["tree_walker", "production_grid"])
@pytest.mark.parametrize("spp", [1, 2, 3, 4, 6, 8, 16, 32])
def test_trace_bit_exact_vs_oracle(capi, oracle, mid_tree, poses8, spp, marcher):
    """Every trace field + the visited-leaf sequence, bit for bit against the oracle, for BOTH marching loops: the tree
    walker and the production brick-grid marcher (rto_trace.marcher = 1: the loop rto_render runs, record switched on)."""
    from rt_octree_b200 import synthetic as S

    W, H = 200, 152
    fx = S.blender_focal(W)
    t, ctx, cam = _setup(capi, mid_tree, W, H, fx)
    for pi in (0, 5):
        cam.transform = poses8[pi]
        ctx.rng_set_frame(pi)
        assert ctx.rng_get() == oracle.frame_rng(pi)
        tr = GpuTrace(capi, W * H, spp, max_seq=64, marcher=marcher)
        capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=tr.pod)
        g = tr.host()
        aux = ctx.read_aux()
        img = ctx.read_image()
        hit = g["steps"] > 0
        th = g["thresh"]
        assert np.all(np.diff(th[hit], axis=1) >= 0) and np.all(th[hit] >= 0)
        # thresholds vs the CPU's log2f: same uniforms, MUFU lg2.approx within 2 ulp-ish relative error
        o_cpu = oracle.render(mid_tree, poses8[pi], W, H, fx, fx, spp, oracle.frame_rng(pi), trace=False)
        o = oracle.render(mid_tree, poses8[pi], W, H, fx, fx, spp, oracle.frame_rng(pi), thresh=th, max_seq=64, want_img=True)
        for k in TRACE_KEYS + ("leaf_seq",):
            assert np.array_equal(g[k], o[k]), "trace field %s differs (spp %d pose %d): %d rays" % (
                k, spp, pi, int((g[k] != o[k]).reshape(W * H, -1).any(1).sum()))
        assert g["steps"].max() > 40 and g["n_hits"].max() >= 1 and (g["term"] >= 0).any()
        assert np.array_equal(aux[3], o["aux"][3])                      # alpha = k/SPP exactly
        assert np.abs(aux - o["aux"]).max() < 1e-5                       # rgb: only ex2.approx / fma order apart
        assert np.abs(img - o["img"]).max() < 1e-5 and np.all(img[..., 3] == 1.0)
        # against CPU-generated thresholds the image differs only where a threshold moved across a leaf boundary
        assert np.mean(aux[3] != o_cpu["aux"][3]) < 2e-3
        assert g["n_loads"].sum() < 0.5 * (o["depth_sum"].sum() + o["steps"].sum())
        assert (g["term"] != -777).all()      # production marcher: grid depth / sigma agreed with the tree at every step


def test_rng_uniforms_exact(capi, oracle, small_tree, poses8):
    """exp(-dst) recovers 1-u: the pcg32 stream (advance(idx*SPP), next_uint) is reproduced exactly on the GPU."""
    from rt_octree_b200 import synthetic as S

    W, H, spp = 64, 48, 6
    t, ctx, cam = _setup(capi, small_tree, W, H, S.blender_focal(W))
    cam.transform = poses8[1]
    ctx.rng_set_frame(7)
    tr = GpuTrace(capi, W * H, spp)
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=tr.pod)
    g = tr.host()
    bits = oracle.uniform_bits(oracle.frame_rng(7), (0, W * H), spp)
    u = ((bits >> 9) | 0x3F800000).astype(np.uint32).view(np.float32) - np.float32(1.0)
    expect = np.sort(-np.log((np.float32(1.0) - u).astype(np.float64)), axis=1)
    hit = g["steps"] > 0
    assert hit.sum() > 100
    got = g["thresh"][hit].astype(np.float64)
    assert np.allclose(got, expect[hit], rtol=2e-6, atol=2e-7)


def test_ndc_anisotropic_rgba_and_options(capi, oracle, poses8):
    from rt_octree_b200 import synthetic as S

    # NDC + anisotropic scale
    tree = S.make_tree(depth=6, shell=1.0, halo=0.05, seed=5, invradius3=(0.45, 0.3, 0.5), offset=(0.5, 0.45, 0.55))
    W, H, fx = 64, 48, 60.0
    pose = S.poses_to_c2w12(np.stack([S.look_at_pose((0.1, 0.05, 0.2), target=(0.0, 0.0, -1.0), world_up=(0, 1, 0))]))[0]
    t, ctx, cam = _setup(capi, tree, W, H, fx)
    t.set_ndc(W, H, fx)
    cam.transform = pose
    for spp, kw in ((1, {}), (6, dict(step_size=3e-4, sigma_thresh=7.0, background_brightness=0.25))):
        ctx.rng_set_frame(0)
        tr = GpuTrace(capi, W * H, spp, max_seq=32)
        capi.launch_renderer(t, cam, _opts(capi, spp, **kw), ctx, trace=tr.pod)
        g = tr.host()
        o = oracle.render(tree, pose, W, H, fx, fx, spp, oracle.frame_rng(0), ndc=(W, H, fx), thresh=g["thresh"], max_seq=32,
                          step_size=kw.get("step_size", 1e-4), sigma_thresh=kw.get("sigma_thresh", 1e-2),
                          background=kw.get("background_brightness", 1.0))
        for k in TRACE_KEYS + ("leaf_seq",):
            assert np.array_equal(g[k], o[k]), k
        assert np.abs(ctx.read_aux() - o["aux"]).max() < 1e-5
        assert o["steps"].sum() > 0
    # RGBA leaves (data_dim 4, rt_core.cuh:322-326)
    base = S.make_tree(depth=5, shell=1.0, halo=0.05, seed=2)
    rgba = dict(base)
    d = base["data"].reshape(-1, 28)
    rgba["data"] = np.ascontiguousarray(np.concatenate([np.abs(d[:, :3]).clip(0, 1), d[:, -1:]], axis=1)).reshape(-1, 2, 2, 2, 4)
    rgba["data_dim"] = np.int64(4)
    rgba["data_format"] = np.array("RGBA")
    W, H = 80, 60
    fx = S.blender_focal(W)
    t2, ctx2, cam2 = _setup(capi, rgba, W, H, fx)
    cam2.transform = poses8[2]
    ctx2.rng_set_frame(2)
    tr = GpuTrace(capi, W * H, 4)
    capi.launch_renderer(t2, cam2, _opts(capi, 4), ctx2, trace=tr.pod)
    g = tr.host()
    o = oracle.render(rgba, poses8[2], W, H, fx, fx, 4, oracle.frame_rng(2), thresh=g["thresh"])
    for k in TRACE_KEYS:
        assert np.array_equal(g[k], o[k]), k
    assert np.abs(ctx2.read_aux() - o["aux"]).max() < 1e-6


@pytest.mark.parametrize("basis_dim", [1, 4, 16, 25])
def test_other_sh_orders(capi, oracle, poses8, basis_dim):
    """SH1 / SH4 / SH16 / SH25 trees (lumisphere.hpp:38-81; payload strides 8 / 16 / 48 / 80 halfs) through the generic
    shading path: traversal bit-exact, colours within 1e-5 of the oracle, grid kernel == tree walker."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=6, shell=1.0, halo=0.05, seed=20 + basis_dim, basis_dim=basis_dim)
    W, H = 80, 64
    fx = S.blender_focal(W)
    t, ctx, cam = _setup(capi, tree, W, H, fx)
    i = t.info
    assert i.basis_dim == basis_dim and i.data_dim == 3 * basis_dim + 1 and i.payload_stride_halfs == (3 * basis_dim + 7) // 8 * 8
    cam.transform = poses8[4]
    ctx.rng_set_frame(4)
    capi.launch_renderer(t, cam, _opts(capi, 6), ctx)                   # production kernel (brick grid)
    aux_grid = ctx.read_aux().copy()
    tr = GpuTrace(capi, W * H, 6)
    capi.launch_renderer(t, cam, _opts(capi, 6), ctx, trace=tr.pod)     # tree walker + trace
    g = tr.host()
    assert np.array_equal(ctx.read_aux(), aux_grid)
    o = oracle.render(tree, poses8[4], W, H, fx, fx, 6, oracle.frame_rng(4), thresh=g["thresh"])
    for k in TRACE_KEYS:
        assert np.array_equal(g[k], o[k]), k
    assert np.array_equal(aux_grid[3], o["aux"][3]) and aux_grid[3].max() == 1.0
    assert np.abs(aux_grid - o["aux"]).max() < 1e-5


def test_rect_render_equals_full_frame(capi, mid_tree, poses8):
    """Tile split: bands rendered separately reproduce the full frame bit for bit."""
    from rt_octree_b200 import synthetic as S

    W, H, spp = 160, 120, 6
    t, ctx, cam = _setup(capi, mid_tree, W, H, S.blender_focal(W))
    cam.transform = poses8[3]
    ctx.rng_set_frame(3)
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx)
    full = ctx.read_aux().copy()
    ctx2 = capi.RenderContext(W, H)
    ctx2.rng_set_frame(3)
    for (y0, y1) in ((0, 37), (37, 90), (90, 120)):
        capi.launch_renderer(t, cam, _opts(capi, spp), ctx2, rect=(0, y0, W, y1))
    assert np.array_equal(ctx2.read_aux(), full)
    ctx3 = capi.RenderContext(W, H)
    ctx3.rng_set_frame(3)
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx3, rect=(13, 5, 101, 77))
    part = ctx3.read_aux()
    assert np.array_equal(part[:, 5:77, 13:101], full[:, 5:77, 13:101])
    assert np.all(part[:, :5] == 0) and np.all(part[:, :, 101:] == 0)


def test_frame_rng_is_pure_function_of_frame(capi, mid_tree, poses8):
    """ctx.rng.advance() per frame (main_headless.cpp:479,506) == rng_set_frame(f): frame sharding is reproducible."""
    from rt_octree_b200 import synthetic as S

    W, H = 96, 72
    t, ctx, cam = _setup(capi, mid_tree, W, H, S.blender_focal(W))
    o = _opts(capi, 6)
    ctx.rng_seed()
    for _ in range(100):
        ctx.rng_advance()
    seq = []
    for f in range(3):
        cam.transform = poses8[f]
        capi.launch_renderer(t, cam, o, ctx)
        seq.append(ctx.read_aux().copy())
        ctx.rng_advance()
    for f in (2, 0):
        cam.transform = poses8[f]
        ctx.rng_set_frame(f)
        capi.launch_renderer(t, cam, o, ctx)
        assert np.array_equal(ctx.read_aux(), seq[f])


@pytest.mark.parametrize("size", [(67, 45), (13, 7), (1, 1), (8, 4), (17, 129)])
def test_ragged_image_sizes(capi, oracle, small_tree, poses8, size):
    """Image sizes that are not multiples of the 16x8 super-tile / 8x4 warp tile (partially filled tiles, a single pixel):
    both kernels bit-exact against the oracle, nothing written outside the image."""
    from rt_octree_b200 import synthetic as S

    W, H = size
    fx = S.blender_focal(max(W, 64)) * 0.5
    t, ctx, cam = _setup(capi, small_tree, W, H, fx)
    cam.transform = poses8[1]
    for spp, denoise in ((6, False), (1, True)):
        ctx.rng_set_frame(1)
        capi.launch_renderer(t, cam, _opts(capi, spp, denoise=denoise), ctx)          # brick-grid kernel
        aux_grid = ctx.read_aux().copy()
        tr = GpuTrace(capi, W * H, spp)
        capi.launch_renderer(t, cam, _opts(capi, spp, denoise=denoise), ctx, trace=tr.pod)
        g = tr.host()
        assert np.array_equal(ctx.read_aux(), aux_grid)
        o = oracle.render(small_tree, poses8[1], W, H, fx, fx, spp, oracle.frame_rng(1), thresh=g["thresh"], want_img=True)
        for k in TRACE_KEYS:
            assert np.array_equal(g[k], o[k]), k
        assert np.array_equal(aux_grid[3], o["aux"][3]) and np.abs(aux_grid - o["aux"]).max() < 1e-5
        if not denoise:
            assert np.abs(ctx.read_image() - o["img"]).max() < 1e-5


def test_root_only_trees(capi, oracle, poses8):
    """A one-node tree (8 leaves, depth 1: no brick grid, the production kernel walks the tree), all empty and all dense."""
    for sigma in (0.0, 50.0):
        data = np.zeros((1, 2, 2, 2, 28), np.float16)
        data[..., -1] = sigma
        data[..., 0] = 0.5
        tree = {"data_dim": np.int64(28), "data_format": np.array("SH9"), "invradius3": np.full(3, 0.375, np.float32),
                "offset": np.full(3, 0.5, np.float32), "child": np.zeros((1, 2, 2, 2), np.int32), "data": data}
        W, H, spp = 24, 24, 4
        t, ctx, cam = _setup(capi, tree, W, H, 60.0)
        assert t.info.max_depth == 1 and t.info.grid_level == 0 and t.info.n_leaves == 8
        cam.transform = poses8[0]
        ctx.rng_set_frame(0)
        capi.launch_renderer(t, cam, _opts(capi, spp), ctx)
        aux = ctx.read_aux().copy()
        tr = GpuTrace(capi, W * H, spp)
        capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=tr.pod)
        g = tr.host()
        assert np.array_equal(ctx.read_aux(), aux)
        o = oracle.render(tree, poses8[0], W, H, 60.0, 60.0, spp, oracle.frame_rng(0), thresh=g["thresh"])
        for k in TRACE_KEYS:
            assert np.array_equal(g[k], o[k]), k
        assert np.array_equal(aux[3], o["aux"][3]) and np.abs(aux - o["aux"]).max() < 1e-5
        if sigma == 0.0:
            assert aux[3].max() == 0.0 and np.all(aux[:3] == 1.0)   # pure background
        else:
            assert aux[3].max() == 1.0


def test_errors(capi, small_tree):
    t, ctx, cam = _setup(capi, small_tree, 32, 32, 40.0)
    o = _opts(capi, 6)
    o.spp = 5
    with pytest.raises(capi.RtoError, match="spp == 5 not supported"):
        capi.launch_renderer(t, cam, o, ctx)
    o.spp = 6
    cam2 = capi.Camera(16, 16, 40.0)
    with pytest.raises(capi.RtoError, match="does not match context"):
        capi.launch_renderer(t, cam2, o, ctx)
    o.enable_probe = True
    with pytest.raises(capi.RtoError, match="enable_probe"):
        capi.launch_renderer(t, cam, o, ctx)
    i = t.info
    assert i.max_depth == 6 and i.payload_stride_halfs == 32 and i.capacity == small_tree["child"].shape[0]
    assert i.n_leaves == int((small_tree["child"] == 0).sum())


def test_full_size_properties(capi, oracle):
    """BASELINE config size (800x800, SPP 6, depth-9 tree): size-independent properties + oracle on a pixel sample."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=9, shell=1.0, halo=0.05, seed=0)
    poses = S.poses_to_c2w12(S.make_poses(200))
    W = H = 800
    fx = S.blender_focal(W)
    t, ctx, cam = _setup(capi, tree, W, H, fx)
    cam.transform = poses[17]
    ctx.rng_set_frame(17)
    spp = 6
    tr = GpuTrace(capi, W * H, spp)
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=tr.pod)
    g = tr.host()
    aux = ctx.read_aux()
    # alpha is a multiple of 1/SPP and equals the collision count; squares channel; background where nothing was hit
    k = np.rint(aux[3] * spp)
    assert np.array_equal(np.float32(k) * np.float32(1.0 / spp), aux[3])
    assert np.array_equal(k.reshape(-1), g["hit_cnt"].sum(1))
    assert np.array_equal(aux[4:], aux[:4] * aux[:4])
    assert np.all(aux[:3, aux[3] == 0] == 1.0)
    assert np.all((g["term"] >= 0) == (g["hit_cnt"].sum(1) == spp))
    assert 0.05 < (aux[3] > 0).mean() < 0.6
    # idempotence
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx)
    assert np.array_equal(ctx.read_aux(), aux)
    # oracle on three row bands of the full-size frame
    for y in (100, 400, 401, 655):
        b, e = y * W, (y + 1) * W
        o = oracle.render(tree, poses[17], W, H, fx, fx, spp, oracle.frame_rng(17), pix_range=(b, e), thresh=g["thresh"][b:e])
        for key in TRACE_KEYS:
            assert np.array_equal(g[key][b:e], o[key]), key
        assert np.abs(aux[:, y] - o["aux"][:, y]).max() < 1e-5


@pytest.fixture(scope="module")
def bench_tree():
    """The EXACT tree bench.py times (same keyword arguments, imported from bench.py)."""
    import bench
    from rt_octree_b200 import synthetic as S

    return S.make_tree(**bench.TREE_KW)


@pytest.mark.parametrize("spp,denoise", [(6, True), (1, False)], ids=["config3_spp6_denoise", "config2_spp1"])
def test_production_marcher_full_frame_bench_workload(capi, oracle, bench_tree, spp, denoise):
    """BASELINE configs 2 and 3 on the exact bench workload (bench.py TREE_KW, 800x800, bench poses, frame rng): the
    PRODUCTION brick-grid marcher's traversal record (steps, termination index, src / t bits, leaf-sequence hash, depth sum,
    hit leaves and counts) equals the oracle's on EVERY pixel of the frame, and the untraced production launch writes the
    same buffers as the traced one."""
    import bench
    from rt_octree_b200 import synthetic as S

    W, H = bench.W, bench.H
    poses, fx = bench.workload_poses()
    t, ctx, cam = _setup(capi, bench_tree, W, H, fx)
    assert t.info.max_depth == 9 and t.info.grid_level == 6
    for f in (0, 117):
        cam.transform = poses[f % len(poses)]
        ctx.rng_set_frame(f, bench.WARMUP_RNG)
        opt = _opts(capi, spp, denoise=denoise)
        capi.launch_renderer(t, cam, opt, ctx)                                   # what bench.py times
        aux_prod = ctx.read_aux().copy()
        img_prod = ctx.read_image().copy()
        tr = GpuTrace(capi, W * H, spp, marcher=1)
        capi.launch_renderer(t, cam, opt, ctx, trace=tr.pod)                     # same loop, record on
        g = tr.host()
        assert np.array_equal(ctx.read_aux(), aux_prod)
        o = oracle.render(bench_tree, poses[f % len(poses)], W, H, fx, fx, spp, oracle.frame_rng(f, bench.WARMUP_RNG),
                          thresh=g["thresh"], want_img=True)
        for key in TRACE_KEYS:
            assert np.array_equal(g[key], o[key]), "%s differs on %d rays (frame %d)" % (
                key, int((g[key] != o[key]).reshape(W * H, -1).any(1).sum()), f)
        assert np.array_equal(aux_prod[3], o["aux"][3]) and np.abs(aux_prod - o["aux"]).max() < 1e-5
        if not denoise:
            assert np.abs(img_prod - o["img"]).max() < 1e-5
        assert g["steps"].max() > 300 and (g["term"] >= 0).mean() > 0.05


@pytest.mark.parametrize("spp", [1, 6, 16])
def test_grid_kernel_equals_tree_walker(capi, oracle, mid_tree, poses8, spp):
    """The production (non-trace) kernel marches over the sparse brick grid; the TRACE kernel walks the tree.  Same hits
    => same shading code => the aux buffers must be BIT-identical, and both equal the oracle's alpha."""
    from rt_octree_b200 import synthetic as S

    W, H = 240, 176
    fx = S.blender_focal(W)
    t, ctx, cam = _setup(capi, mid_tree, W, H, fx)
    i = t.info
    assert i.grid_level == i.max_depth - 3 and i.n_bricks > 0 and i.grid_bytes > 0
    for pi in (2, 7):
        cam.transform = poses8[pi]
        ctx.rng_set_frame(pi)
        capi.launch_renderer(t, cam, _opts(capi, spp), ctx)                 # grid path
        aux_grid = ctx.read_aux().copy()
        img_grid = ctx.read_image().copy()
        tr = GpuTrace(capi, W * H, spp)
        capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=tr.pod)   # tree walker
        g = tr.host()
        assert np.array_equal(ctx.read_aux(), aux_grid)
        assert np.array_equal(ctx.read_image(), img_grid)
        o = oracle.render(mid_tree, poses8[pi], W, H, fx, fx, spp, oracle.frame_rng(pi), thresh=g["thresh"], trace=False)
        assert np.array_equal(aux_grid[3], o["aux"][3]) and np.abs(aux_grid - o["aux"]).max() < 1e-5
        assert aux_grid[3].max() == 1.0


def test_byte_brick_plane_equals_word_plane(capi, mid_tree, poses8, monkeypatch):
    """The marching loop reads the byte plane of the bricks (depth | dense flag) and fetches the 4-byte leaf word only in
    cells with non-zero sigma; RTO_GRID8=0 reads the leaf words on every step.  Same words => identical buffers, also
    with a NEGATIVE sigma threshold (then every sigma = 0 cell counts as dense) and with negative-zero sigmas."""
    from rt_octree_b200 import synthetic as S

    W, H = 240, 176
    fx = S.blender_focal(W)
    tree = dict(mid_tree)
    data = tree["data"].copy()
    sig = data[..., -1].view(np.uint16)
    leaf0 = (tree["child"] == 0) & (sig == 0)
    flip = leaf0 & (np.random.RandomState(1).rand(*leaf0.shape) < 0.3)
    sig[flip] = 0x8000                                       # -0.0: non-zero bits, still "not above" any threshold >= 0
    tree["data"] = data
    for tr_, thresh in ((mid_tree, 1e-2), (mid_tree, -1.0), (tree, 1e-2), (tree, 0.0)):
        t, ctx, cam = _setup(capi, tr_, W, H, fx)
        cam.transform = poses8[3]
        out = {}
        for g8 in ("1", "0", "nodefer", "nofused"):   # "1" = production: byte plane, deferred hits, fused table indices
            monkeypatch.setenv("RTO_GRID8", "0" if g8 == "0" else "1")
            monkeypatch.setenv("RTO_DEFER_HITS", "0" if g8 == "nodefer" else "1")   # collisions: leaf-id planes vs root descent
            monkeypatch.setenv("RTO_FUSED_INDEX", "0" if g8 == "nofused" else "1")  # table indices: shift-built (v9) vs fp adder
            ctx.rng_set_frame(3)
            capi.launch_renderer(t, cam, _opts(capi, 6, sigma_thresh=thresh), ctx)
            out[g8] = (ctx.read_aux().copy(), ctx.read_image().copy())
        monkeypatch.delenv("RTO_GRID8")
        monkeypatch.delenv("RTO_DEFER_HITS")
        monkeypatch.delenv("RTO_FUSED_INDEX")
        assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])
        assert np.array_equal(out["1"][0], out["nodefer"][0]) and np.array_equal(out["1"][1], out["nodefer"][1])
        assert np.array_equal(out["1"][0], out["nofused"][0]) and np.array_equal(out["1"][1], out["nofused"][1])
        assert out["1"][0][3].max() == 1.0


def test_tt_shaped_depth10_1080p(capi, oracle):
    """BASELINE config 4 shape: anisotropic depth-10 tree, 1920x1080, OpenCV-convention poses through the tt loader math.
    Size-independent properties on the full frame + bit-exact oracle on sampled rows."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=10, shell=0.03, halo=0.02, seed=1, invradius3=(0.30, 0.42, 0.36), offset=(0.5, 0.52, 0.48))
    W, H = 1920, 1080
    fx = 1166.0
    poses = S.poses_to_c2w12(S.make_poses(200, radius=3.2, elevation_deg=20.0))
    t, ctx, cam = _setup(capi, tree, W, H, fx)
    i = t.info
    assert i.max_depth == 10 and i.grid_level == 7
    cam.transform = poses[41]
    ctx.rng_set_frame(41)
    spp = 6
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx)           # grid kernel
    aux = ctx.read_aux().copy()
    tr = GpuTrace(capi, W * H, spp)
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=tr.pod)
    g = tr.host()
    assert np.array_equal(ctx.read_aux(), aux)                    # grid == tree walker, bit for bit
    k = np.rint(aux[3] * spp)
    assert np.array_equal(np.float32(k) * np.float32(1.0 / spp), aux[3]) and np.array_equal(k.reshape(-1), g["hit_cnt"].sum(1))
    assert np.array_equal(aux[4:], aux[:4] * aux[:4]) and aux[3].max() == 1.0
    # the production marcher's own record, FULL FRAME against the oracle (every field, every pixel)
    trg = GpuTrace(capi, W * H, spp, marcher=1)
    capi.launch_renderer(t, cam, _opts(capi, spp), ctx, trace=trg.pod)
    gg = trg.host()
    assert np.array_equal(ctx.read_aux(), aux)
    o = oracle.render(tree, poses[41], W, H, fx, fx, spp, oracle.frame_rng(41), thresh=gg["thresh"])
    for key in TRACE_KEYS:
        assert np.array_equal(gg[key], o[key]), "production marcher: %s differs" % key
        assert np.array_equal(g[key], o[key]), "tree walker: %s differs" % key
    assert np.array_equal(aux[3], o["aux"][3]) and np.abs(aux - o["aux"]).max() < 1e-5


def test_4k_tile_split_bands(capi, mid_tree, poses8, net_weights):
    """BASELINE config 5 shape on one GPU: 3840x2160 SPP 6 + denoise rendered as 8 row bands (+6-row halo) equals the
    full frame bit for bit (the multi-GPU version gathers the bands with NCCL: tools/tile_split_check.py)."""
    from rt_octree_b200 import sharding as SH, synthetic as S

    W, H = 3840, 2160
    fx = float(np.float32(S.blender_focal(W)))
    t = capi.N3Tree(mid_tree)
    cam = capi.Camera(W, H, fx, fx)
    cam.transform = poses8[6]
    o = _opts(capi, 6, denoise=True)
    net = capi.Denoiser(net_weights)
    ctx = capi.RenderContext(W, H)
    ctx.rng_set_frame(6)
    capi.launch_renderer(t, cam, o, ctx)
    net.denoise(cam, ctx)
    full = ctx.read_image().copy()
    out = np.zeros_like(full)
    for band in SH.tile_bands(H, 8):
        c = capi.RenderContext(W, H)
        c.rng_set_frame(6)
        y0, y1 = SH.render_rows_for_band(band, H, True)
        capi.launch_renderer(t, cam, o, c, rect=(0, y0, W, y1))
        net.denoise(cam, c, rows=band)
        out[band[0]:band[1]] = c.read_image()[band[0]:band[1]]
        c.close()
    assert np.array_equal(out, full)
    assert np.isfinite(full).all() and np.all(full[..., 3] == 1.0)
