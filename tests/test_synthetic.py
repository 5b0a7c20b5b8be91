"""Input generators: tree.npz key set / invariants, pose conventions, quantised-tree decode."""
import numpy as np

from util import tree_depth


def test_tree_structure(small_tree):
    t = small_tree
    child = t["child"].reshape(-1, 8)
    cap = child.shape[0]
    assert t["child"].dtype == np.int32 and t["data"].dtype == np.float16
    assert t["data"].shape == (cap, 2, 2, 2, 28) and int(t["data_dim"]) == 28 and str(t["data_format"]) == "SH9"
    tgt = np.arange(cap)[:, None] + child
    assert np.all((child == 0) | ((tgt > 0) & (tgt < cap)))
    # every non-root node is referenced exactly once
    refs = np.bincount(tgt[child != 0], minlength=cap)
    assert refs[0] == 0 and np.all(refs[1:] == 1)
    assert tree_depth(t) == 6
    sig = t["data"].reshape(cap * 8, 28)[:, -1]
    leaf = child.reshape(-1) == 0
    assert (sig[leaf] > 0).any() and (sig[leaf] == 0).any()


def test_tree_deterministic():
    from rt_octree_b200 import synthetic as S

    a, b = S.make_tree(depth=5, seed=7), S.make_tree(depth=5, seed=7)
    assert np.array_equal(a["child"], b["child"]) and np.array_equal(a["data"], b["data"])


def test_npz_roundtrip_through_host_loader_fields(tmp_path, small_tree):
    from rt_octree_b200 import synthetic as S

    p = tmp_path / "tree.npz"
    S.write_tree_npz(str(p), small_tree)
    z = np.load(str(p))
    for k in ("data_dim", "data_format", "invradius3", "offset", "child", "data"):
        assert k in z.files
    assert str(z["data_format"]) == "SH9"


def test_poses_blender_convention():
    from rt_octree_b200 import synthetic as S

    P = S.make_poses(4)
    for m in P:
        R = m[:3, :3]
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-12)
        assert np.isclose(np.linalg.norm(m[:3, 3]), S.BLENDER_RADIUS)
        # camera looks along -back towards the origin
        assert np.allclose(-m[:3, 2], -m[:3, 3] / np.linalg.norm(m[:3, 3]), atol=1e-12)
    c = S.poses_to_c2w12(P)
    assert c.shape == (4, 12) and np.allclose(c[0, 9:], P[0][:3, 3])
    assert abs(S.blender_focal(800) - 1111.111) < 1e-2


def test_quantized_decode_matches_reference_layout(capi):
    """n3tree.cpp:279-340: data[i, j + n_ret + k*n_basis] = quant_colors[j, quant_map[j,i], k]."""
    rs = np.random.default_rng(0)
    cap, N, basis = 3, 2, 4
    n_child = cap * 8
    data_dim = 3 * basis + 1
    n_ret = 1
    qc = rs.normal(size=(basis - n_ret, 65536, 3)).astype(np.float16)
    qm = rs.integers(0, 65536, size=(basis - n_ret, cap, 2, 2, 2)).astype(np.uint16)
    sigma = rs.uniform(0, 9, (cap, 2, 2, 2)).astype(np.float16)
    ret = rs.normal(size=(n_ret, cap, 2, 2, 2, 3)).astype(np.float16)
    z = {"quant_colors": qc, "quant_map": qm, "sigma": sigma, "data_retained": ret}
    d = capi.decode_quantized(z, cap, N, data_dim)
    # scalar restatement of the reference loops
    exp = np.zeros((n_child, data_dim), np.float16)
    qmf = qm.reshape(basis - n_ret, n_child)
    for i in range(n_child):
        for j in range(basis - n_ret):
            for k in range(3):
                exp[i, j + n_ret + k * basis] = qc[j, qmf[j, i], k]
        exp[i, data_dim - 1] = sigma.reshape(-1)[i]
        for j in range(n_ret):
            for k in range(3):
                exp[i, j + k * basis] = ret.reshape(n_ret, n_child, 3)[j, i, k]
    assert np.array_equal(d, exp)
