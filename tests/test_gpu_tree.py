"""GPU tests of the device octree loader (rt_octree_b200/csrc/rto_tree.cu), through the C ABI: every HBM plane is built on
the GPU and read back with rto_tree_read_plane.

* node words / payload vs a numpy statement of the layout (bit-exact),
* the brick grid built on the device vs the host builder (rto_grid_host.h, RTO_GRID_BUILD=host): identical tables,
* quantised (svox-compressed) files: codebook gather on the GPU vs the host decode of n3tree.cpp:279-340 (bit-exact),
* structure check on the device: out-of-range offsets, cycles, depth."""
import ctypes as C
import os

import numpy as np
import pytest

from util import encode_nodes, tree_depth

pytestmark = pytest.mark.gpu


def _host_grid_tree(capi, tree):
    os.environ["RTO_GRID_BUILD"] = "host"
    try:
        return capi.N3Tree(tree)
    finally:
        del os.environ["RTO_GRID_BUILD"]


def _payload_expected(tree, stride):
    cap = tree["child"].shape[0]
    dd = int(tree["data_dim"])
    d = np.ascontiguousarray(tree["data"]).reshape(cap * 8, dd).view(np.uint16)
    out = np.zeros((cap * 8, stride), np.uint16)
    out[:, :dd - 1] = d[:, :dd - 1]
    return out


@pytest.mark.parametrize("depth", [4, 5, 6, 8])
def test_planes_and_device_grid_equal_host_builder(capi, depth):
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=depth, shell=1.0, halo=0.1, seed=depth)
    t = capi.N3Tree(tree)
    i = t.info
    assert i.max_depth == tree_depth(tree) == depth
    assert i.n_leaves == int((tree["child"] == 0).sum())
    assert np.array_equal(t.read_plane("nodes"), encode_nodes(tree))
    pay = t.read_plane("payload").view(np.uint16).reshape(-1, i.payload_stride_halfs)
    assert np.array_equal(pay, _payload_expected(tree, i.payload_stride_halfs))
    assert i.grid_level == depth - 3
    h = _host_grid_tree(capi, tree)
    assert h.info.grid_level == i.grid_level and h.info.n_bricks == i.n_bricks and h.info.grid_bytes == i.grid_bytes
    top_d, top_h = t.read_plane("grid_top"), h.read_plane("grid_top")
    assert top_d.size == 8 ** (depth - 3)
    assert np.array_equal(top_d, top_h), "%d top-table cells differ" % int((top_d != top_h).sum())
    br_d, br_h = t.read_plane("grid_bricks"), h.read_plane("grid_bricks")
    assert np.array_equal(br_d, br_h), "%d brick cells differ" % int((br_d != br_h).sum())
    # march table of the fused-index marcher (rto_ray.cuh march_top_entry): leaf words unchanged, brick ids as biased offsets
    K = depth - 3
    tm = t.read_plane("grid_march_top")
    cell = np.arange(top_d.size, dtype=np.int64)
    off = 512 * (cell >> (2 * K)) + 64 * ((cell >> K) & ((1 << K) - 1)) + 8 * (cell & ((1 << K) - 1))
    leafcell = (top_d & 0x80000000) != 0
    want = np.where(leafcell, top_d.astype(np.int64), top_d.astype(np.int64) * 512 + (1 << 18) - off)
    assert np.array_equal(tm.astype(np.int64), want) and not (tm[~leafcell] & 0x80000000).any()
    assert np.array_equal(tm, h.read_plane("grid_march_top"))
    # every brick cell is a leaf word whose depth field lies in (K, K+3]
    if br_d.size:
        d = ((br_d >> 23) & 0xff).astype(np.int64) - 127
        assert np.all(br_d & 0x80000000) and d.min() > depth - 3 and d.max() == depth
        # byte plane (rto_ray.cuh brick_byte): depth | 0x80 where the sigma bits are non-zero
        b8 = t.read_plane("grid_bricks8")
        assert np.array_equal(b8, (d | np.where(br_d & 0xffff, 0x80, 0)).astype(np.uint8))
        assert np.array_equal(b8, h.read_plane("grid_bricks8"))
        assert (b8 & 0x80).any() and not (b8 & 0x80).all()


def test_device_grid_bench_tree_equals_host_builder(capi):
    """The bench tree (depth 9, 2.04 M nodes, ~31 k bricks): same tables from both builders."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=9, shell=1.0, halo=0.25, seed=0)
    t = capi.N3Tree(tree)
    h = _host_grid_tree(capi, tree)
    assert t.info.n_bricks == h.info.n_bricks > 10000
    assert np.array_equal(t.read_plane("grid_top"), h.read_plane("grid_top"))
    assert np.array_equal(t.read_plane("grid_bricks"), h.read_plane("grid_bricks"))


def _leaf_of_cells(tree, level):
    """numpy: flat leaf index (node*8+octant, the reference's sub_ptr) and leaf depth of every cell of the 2^level grid,
    by descending all cells at once (n3tree_query.hpp:13-48 on integer coordinates)."""
    child = tree["child"].reshape(-1).astype(np.int64)
    S = 1 << level
    x, y, z = np.meshgrid(np.arange(S), np.arange(S), np.arange(S), indexing="ij")
    x, y, z = x.reshape(-1), y.reshape(-1), z.reshape(-1)
    node = np.zeros(x.size, np.int64)
    leaf = np.full(x.size, -1, np.int64)
    depth = np.zeros(x.size, np.int64)
    for d in range(1, level + 1):
        sh = level - d
        e = node * 8 + ((((x >> sh) & 1) << 2) | (((y >> sh) & 1) << 1) | ((z >> sh) & 1))
        live = leaf < 0
        is_leaf = live & (child[e] == 0)
        leaf[is_leaf], depth[is_leaf] = e[is_leaf], d
        go = live & ~is_leaf
        node[go] = node[go] + child[e[go]]
    return leaf.reshape(S, S, S), depth.reshape(S, S, S)


@pytest.mark.parametrize("depth", [4, 6, 7])
def test_leaf_id_planes(capi, depth):
    """Leaf-id planes (what collisions are resolved through after the march): every level-K leaf cell and every brick cell
    names the leaf a root descent finds, for both grid builders."""
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=depth, shell=1.0, halo=0.2, seed=40 + depth)
    K = depth - 3
    fine_leaf, _ = _leaf_of_cells(tree, depth)
    coarse_leaf, coarse_depth = _leaf_of_cells(tree, K)
    for t in (capi.N3Tree(tree), _host_grid_tree(capi, tree)):
        top = t.read_plane("grid_top").reshape((1 << K,) * 3)
        lt = t.read_plane("grid_leaf_top").reshape((1 << K,) * 3)
        lb = t.read_plane("grid_leaf_bricks").reshape(-1, 8, 8, 8)
        is_leaf = (top & 0x80000000) != 0
        assert np.array_equal(is_leaf, coarse_depth > 0)
        assert np.array_equal(lt[is_leaf].astype(np.int64), coarse_leaf[is_leaf])
        assert (is_leaf.any() or K == 1) and (~is_leaf).any() and lb.shape[0] == t.info.n_bricks
        cx, cy, cz = np.nonzero(~is_leaf)
        for bx, by, bz in list(zip(cx, cy, cz))[:: max(1, len(cx) // 400)]:      # a few hundred bricks, every cell of each
            want = fine_leaf[bx * 8:bx * 8 + 8, by * 8:by * 8 + 8, bz * 8:bz * 8 + 8]
            assert np.array_equal(lb[top[bx, by, bz]].astype(np.int64), want)


def test_permuted_node_order_builds_the_same_grid(capi, small_tree):
    """svox files are not breadth-first: shuffle the node numbering (negative relative offsets appear) and check that the
    geometry the grid encodes does not change."""
    tree = dict(small_tree)
    child = tree["child"].reshape(-1, 8).astype(np.int64)
    cap = child.shape[0]
    rs = np.random.RandomState(5)
    perm = np.concatenate([[0], 1 + rs.permutation(cap - 1)])   # new id of old node i (root stays 0)
    tgt_old = np.arange(cap)[:, None] + child                  # absolute target, old numbering
    new_child = np.zeros_like(child)
    new_child[perm] = np.where(child != 0, perm[np.where(child != 0, tgt_old, 0)] - perm[:, None], 0)
    assert (new_child < 0).any()
    data = tree["data"].reshape(cap, 8, -1)
    new_data = np.empty_like(data)
    new_data[perm] = data
    tree["child"] = new_child.astype(np.int32).reshape(cap, 2, 2, 2)
    tree["data"] = new_data.reshape(tree["data"].shape)
    a, b = capi.N3Tree(small_tree), capi.N3Tree(tree)
    assert a.info.max_depth == b.info.max_depth and a.info.n_bricks == b.info.n_bricks
    assert np.array_equal(a.read_plane("grid_top"), b.read_plane("grid_top"))      # brick ids follow the geometry, not the node ids
    assert np.array_equal(a.read_plane("grid_bricks"), b.read_plane("grid_bricks"))
    h = _host_grid_tree(capi, tree)
    assert np.array_equal(b.read_plane("grid_bricks"), h.read_plane("grid_bricks"))


@pytest.mark.parametrize("n_ret", [0, 1, 3])
def test_quantized_tree_decoded_on_the_gpu(capi, small_tree, n_ret):
    """compress_octree.py:68-119 layout -> rto_tree_create_quantized; planes equal the ones built from the host decode
    (capi.decode_quantized restates n3tree.cpp:279-340)."""
    cap = small_tree["child"].shape[0]
    basis, dd = 9, 28
    rs = np.random.RandomState(11 + n_ret)
    z = {"data_dim": np.int64(dd), "data_format": np.array("SH9"), "invradius3": small_tree["invradius3"],
         "offset": small_tree["offset"], "child": small_tree["child"],
         "quant_colors": rs.normal(size=(basis - n_ret, 65536, 3)).astype(np.float16),
         "quant_map": rs.randint(0, 65536, size=(basis - n_ret, cap, 2, 2, 2)).astype(np.uint16),
         "sigma": np.ascontiguousarray(small_tree["data"][..., -1])}
    if n_ret:
        z["data_retained"] = rs.normal(size=(n_ret, cap, 2, 2, 2, 3)).astype(np.float16)
    dense = dict(small_tree)
    dense["data"] = capi.decode_quantized(z, cap, 2, dd).reshape(cap, 2, 2, 2, dd)
    q, d = capi.N3Tree(z), capi.N3Tree(dense)
    assert q.info.n_leaves == d.info.n_leaves and q.info.max_depth == d.info.max_depth
    assert np.array_equal(q.read_plane("nodes"), d.read_plane("nodes"))
    assert np.array_equal(q.read_plane("payload").view(np.uint16), d.read_plane("payload").view(np.uint16))
    assert np.array_equal(q.read_plane("grid_bricks"), d.read_plane("grid_bricks"))


def test_malformed_trees_are_rejected(capi):
    lib = capi.load()
    h = C.c_void_p()
    off = np.zeros(3, np.float32)
    data = np.zeros((3, 8, 28), np.float16)

    def create(child):
        child = np.ascontiguousarray(child, np.int32)
        return lib.rto_tree_create(C.byref(h), child.ctypes.data, data.ctypes.data, child.shape[0], 2, 28, capi.FORMAT_SH, 9,
                                   off.ctypes.data, off.ctypes.data)

    c = np.zeros((1, 8), np.int32)
    c[0, 3] = 5                                    # points outside a 1-node tree
    assert create(c) == capi.RTO_ERR_INVALID and b"malformed tree" in lib.rto_last_error()
    c = np.zeros((3, 8), np.int32)
    c[0, 0], c[1, 2], c[2, 7] = 1, 1, -1           # 0 -> 1 -> 2 -> 1: a cycle
    assert create(c) == capi.RTO_ERR_INVALID and b"malformed tree" in lib.rto_last_error()
    c = np.zeros((3, 8), np.int32)
    c[0, 0], c[0, 1], c[0, 2], c[0, 3] = 1, 1, 2, 2   # shared subtrees: 1 + 4 visits for 3 nodes
    assert create(c) == capi.RTO_ERR_INVALID
    c = np.zeros((3, 8), np.int32)
    c[0, 0], c[1, 5] = 1, 1                        # a proper 3-level chain
    assert create(c) == capi.RTO_OK
    info = capi.TreeInfoPOD()
    assert lib.rto_tree_get_info(h, C.byref(info)) == capi.RTO_OK
    assert info.max_depth == 3 and info.n_leaves == 22 and info.grid_level == 0
    lib.rto_tree_destroy(h)


def test_deep_chain_depth_limit(capi):
    """23 levels is the coordinate precision (RTO_COORD_BITS); a 24-level chain is refused."""
    lib = capi.load()
    off = np.zeros(3, np.float32)
    for levels, expect in ((23, capi.RTO_OK), (24, capi.RTO_ERR_UNSUPPORTED)):
        child = np.zeros((levels, 8), np.int32)
        child[:-1, 0] = 1
        data = np.zeros((levels, 8, 28), np.float16)
        h = C.c_void_p()
        rc = lib.rto_tree_create(C.byref(h), child.ctypes.data, data.ctypes.data, levels, 2, 28, capi.FORMAT_SH, 9,
                                 off.ctypes.data, off.ctypes.data)
        assert rc == expect, lib.rto_last_error()
        if rc == capi.RTO_OK:
            lib.rto_tree_destroy(h)
