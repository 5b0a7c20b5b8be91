"""pytest configuration: markers, oracle/harness builds and shared fixtures.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI load/exports (runs without a GPU, a few minutes).
`-m gpu`       : the parity tests proper; they call through the C ABI (rt_octree_b200/capi.py) on cuda:0.
Nothing in the gpu tests reads /root/reference: the reference binaries travel prebuilt in oracle/_ref/.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def host_ray_lib():
    """Test-only host instantiation of rto_ray.cuh (tests/host_ray_harness.cpp)."""
    import ctypes as C

    src = os.path.join(ROOT, "tests", "host_ray_harness.cpp")
    hdr = os.path.join(ROOT, "rt_octree_b200", "csrc", "rto_ray.cuh")
    hdr2 = os.path.join(ROOT, "rt_octree_b200", "csrc", "rto_grid_host.h")
    so = os.path.join(ROOT, "tests", "_build", "libhost_ray.so")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2)):
        subprocess.run(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-mf16c", "-fPIC", "-shared",
                        src, "-o", so], check=True)
    lib = C.CDLL(so)
    P, I, F, U64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
    lib.host_ray_walk.restype = I
    lib.host_ray_walk.argtypes = [P, I, P, P, P, F, F, F, F, F, F, F, I, I, I, U64, U64, I, I, P] + [P] * 11 + [I]
    lib.host_ray_set_grid.restype = I
    lib.host_ray_set_grid.argtypes = [P, P, I, C.c_int64, I, P]
    lib.host_ray_use_byte_bricks.restype = None
    lib.host_ray_use_byte_bricks.argtypes = [I]
    lib.host_ray_use_deferred_hits.restype = None
    lib.host_ray_use_deferred_hits.argtypes = [I]
    lib.host_ray_use_fused_index.restype = None
    lib.host_ray_use_fused_index.argtypes = [I]
    return lib


@pytest.fixture(scope="session")
def small_tree():
    from rt_octree_b200 import synthetic as S

    return S.make_tree(depth=6, shell=1.0, halo=0.05, seed=3)


@pytest.fixture(scope="session")
def mid_tree():
    from rt_octree_b200 import synthetic as S

    return S.make_tree(depth=8, shell=1.0, halo=0.1, seed=0)


@pytest.fixture(scope="session")
def poses8():
    from rt_octree_b200 import synthetic as S

    return S.poses_to_c2w12(S.make_poses(8))


@pytest.fixture(scope="session")
def net_weights():
    """The reference's GuidanceNet(8,32,5,2,4) weights (seed 0), exported by tools/make_golden.py."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "guidance_net_ref.npz"))
    return {k: g[k] for k in ("w1", "b1", "w2", "b2")}


@pytest.fixture(scope="session")
def capi():
    from rt_octree_b200 import capi as c

    c.load()
    return c
