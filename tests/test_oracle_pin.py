"""Pins the CPU oracle (oracle/rt_oracle.c) against the REFERENCE:
  * tests/golden/trace_ref_cpu.npz  — aux buffers produced by the reference's own trace_ray (rt_core.cuh
    host-compiled by oracle/build_ref.sh; generator tools/make_golden.py) — runs everywhere;
  * oracle/_ref/libref_cpu.so       — the same reference code, live, on more cases — when it has been built.
alpha (= k/SPP) must be bit-identical: it is decided purely by the traversal; rgb differs only by the
host compiler's choice of fused multiply-adds in the SH dot products (<= 1e-6).
"""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_reference_golden(oracle, poses8):
    from rt_octree_b200 import synthetic as S

    g = np.load(os.path.join(ROOT, "tests", "golden", "trace_ref_cpu.npz"))
    tree = S.make_tree(depth=int(g["tree_depth"]), shell=float(g["tree_shell"]), halo=float(g["tree_halo"]), seed=int(g["tree_seed"]))
    W, H, fx = int(g["W"]), int(g["H"]), float(g["fx"])
    n_cases = 0
    for key in g.files:
        if not key.startswith("aux_spp"):
            continue
        spp = int(key.split("_")[1][3:])
        pi = int(key.split("_")[2][4:])
        o = oracle.render(tree, poses8[pi], W, H, fx, fx, spp, oracle.frame_rng(pi))
        ref = g[key]
        assert np.array_equal(o["aux"][3], ref[3]), "alpha differs from the reference (%s)" % key
        assert np.abs(o["aux"] - ref).max() <= 1e-6, key
        assert o["aux"][3].max() > 0, "degenerate fixture"
        n_cases += 1
    assert n_cases == 6


@pytest.mark.parametrize("spp", [1, 2, 3, 4, 6, 8, 16, 32])
def test_oracle_matches_live_reference_cpu(oracle, mid_tree, poses8, spp):
    if oracle.ref_cpu_lib() is None:
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    W, H = 96, 80
    from rt_octree_b200 import synthetic as S

    fx = S.blender_focal(W)
    for pi in (1, 6):
        rng = oracle.frame_rng(pi)
        ref = oracle.ref_cpu_render(mid_tree, poses8[pi], W, H, fx, fx, spp, rng)
        o = oracle.render(mid_tree, poses8[pi], W, H, fx, fx, spp, rng)
        assert np.array_equal(o["aux"][3], ref[3])
        assert np.abs(o["aux"] - ref).max() <= 1e-6
        # aux layout (volrend.cu:187-202): ch4..7 are the squares of ch0..3
        assert np.array_equal(o["aux"][4:], o["aux"][:4] * o["aux"][:4])


@pytest.mark.parametrize("basis_dim", [1, 4, 16, 25])
def test_oracle_other_sh_orders_match_live_reference_cpu(oracle, poses8, basis_dim):
    """The reference supports SH1/4/9/16/25 (lumisphere.hpp:38-81); the shipped scenes use SH9.  Pin the other orders too."""
    if oracle.ref_cpu_lib() is None:
        pytest.skip("oracle/_ref/libref_cpu.so not built (needs /root/reference)")
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=6, shell=1.0, halo=0.05, seed=20 + basis_dim, basis_dim=basis_dim)
    assert int(tree["data_dim"]) == 3 * basis_dim + 1
    W, H = 80, 64
    fx = S.blender_focal(W)
    rng = oracle.frame_rng(4)
    ref = oracle.ref_cpu_render(tree, poses8[4], W, H, fx, fx, 6, rng)
    o = oracle.render(tree, poses8[4], W, H, fx, fx, 6, rng)
    assert o["aux"][3].max() == 1.0
    assert np.array_equal(o["aux"][3], ref[3])
    assert np.abs(o["aux"] - ref).max() <= 1e-6


def test_unsupported_spp_raises(oracle, small_tree, poses8):
    with pytest.raises(ValueError, match="spp == 5 not supported"):
        oracle.render(small_tree, poses8[0], 8, 8, 100.0, 100.0, 5, oracle.frame_rng(0))


def test_pcg32_known_values(oracle):
    # pcg32(20230418): state/inc after seed(initstate, 1) — pcg32.h:53-59
    st, inc = oracle.pcg32_seed(20230418)
    assert inc == 3
    M, mask = 0x5851F42D4C957F2D, (1 << 64) - 1
    s = 0
    s = (s * M + 3) & mask
    s = (s + 20230418) & mask
    s = (s * M + 3) & mask
    assert st == s
    # advance(k) == k single steps
    x = st
    for _ in range(37):
        x = (x * M + inc) & mask
    assert oracle.lib().rto_oracle_pcg32_advance(st, inc, 37) == x
    # frame rng = 100+f advances of 2^32 composed
    a = oracle.frame_rng(3)[0]
    b = st
    for _ in range(103):
        b = oracle.lib().rto_oracle_pcg32_advance(b, inc, 1 << 32)
    assert a == b


def test_guidance_net_oracle_matches_reference_module(oracle):
    """GuidanceNetCompact fp16 forward of the reference's own module on CPU (tools/make_golden.py)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "guidance_net_ref.npz"))
    assert bool(g["have_fp16"])
    w = {k: g[k] for k in ("w1", "b1", "w2", "b2")}
    wm, gm = oracle.guidance_net(g["aux"], w, fused_bias=True)   # PyTorch CPU fp16 conv rounds once after the bias
    d = np.abs(gm - g["guidance_fp16"])
    assert d.max() <= 2.0 ** -9 + 1e-7          # at most one fp16 ulp in [2,4) (accumulation order only)
    assert (d == 0).mean() > 0.99
    assert np.abs(wm - g["weight_fp16"]).max() < 1e-3
    assert np.allclose(wm.sum(0), 1.0, atol=1e-6)
    # and against the fp32 un-compacted 5-branch model: only fp16 quantisation apart
    assert np.abs(gm - g["guidance_fp32_full"]).max() < 8e-3
    # the GPU (ATen cuDNN) rounding variant differs from it by <= 1.5 fp16 ulp
    wm2, gm2 = oracle.guidance_net(g["aux"], w, fused_bias=False)
    assert np.abs(gm2 - gm).max() <= 3 * 2.0 ** -10 + 1e-7


def _filter_bruteforce(weight, guidance, img_in):
    """Independent numpy statement of filtering.cu:108-228 (float64 accumulation)."""
    L, H, W = weight.shape
    out = np.zeros((H, W, 4))
    out[..., 3] = 1.0
    for l in range(L):
        S = l + 1
        g = np.full((H + 2 * S, W + 2 * S), -np.inf)
        g[S:S + H, S:S + W] = guidance[l]
        rgb = np.zeros((H + 2 * S, W + 2 * S, 3))
        rgb[S:S + H, S:S + W] = img_in[..., :3]
        num = np.zeros((H, W, 3))
        den = np.zeros((H, W))
        mx = np.full((H, W), -np.inf)
        for dy in range(2 * S + 1):
            for dx in range(2 * S + 1):
                mx = np.maximum(mx, g[dy:dy + H, dx:dx + W])
        for dy in range(2 * S + 1):
            for dx in range(2 * S + 1):
                k = np.exp(g[dy:dy + H, dx:dx + W] - mx)
                den += k
                num += rgb[dy:dy + H, dx:dx + W] * k[..., None]
        out[..., :3] += weight[l][..., None] * num / den[..., None]
    return out


@pytest.mark.parametrize("L", [1, 4, 6])
def test_filter_oracle_matches_bruteforce(oracle, L):
    rs = np.random.default_rng(L)
    H, W = 21, 35
    guidance = rs.uniform(0, 6, (L, H, W)).astype(np.float32)
    weight = rs.uniform(0, 1, (L, H, W)).astype(np.float32)
    weight /= weight.sum(0, keepdims=True)
    img = rs.uniform(0, 1, (H, W, 4)).astype(np.float32)
    out = oracle.filtering(weight, guidance, img)
    ref = _filter_bruteforce(weight, guidance, img)
    assert np.abs(out - ref).max() < 2e-6
    assert np.all(out[..., 3] == 1.0)
    # constant image is a fixed point (weights sum to 1 per level and over levels)
    const = np.full((H, W, 4), 0.37, np.float32)
    assert np.abs(oracle.filtering(weight, guidance, const)[..., :3] - 0.37).max() < 1e-6


def test_filter_rejects_unsupported_levels(oracle):
    z = np.zeros((7, 4, 4), np.float32)
    with pytest.raises(ValueError, match="Kernel size == 15 not supported"):
        oracle.filtering(z, z, np.zeros((4, 4, 4), np.float32))
