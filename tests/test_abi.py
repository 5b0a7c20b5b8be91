"""The C-ABI shared library loads and exports every symbol include/rtoctree_b200.h declares; host-side mirrors of the
reference interface (RenderOptions JSON binding, DataFormat parse, error conventions).  No compute calls (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rtoctree_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rto_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "librtoctree_b200.so does not export %s" % s
    assert sorted(capi.EXPORTS) == syms, "capi.EXPORTS out of sync with the header"
    assert lib.rto_abi_version() == 2


def test_no_gpu_fails_loudly(capi):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.RtoError) as e:
        capi.RenderContext(8, 8)
    assert e.value.code == capi.RTO_ERR_CUDA


def test_argument_validation_without_gpu(capi):
    lib = capi.load()
    h = C.c_void_p()
    child = np.zeros(8, np.int32)
    data = np.zeros(8 * 28, np.float16)
    off = np.zeros(3, np.float32)
    args = lambda **kw: [C.byref(h), child.ctypes.data, data.ctypes.data, kw.get("cap", 1), kw.get("N", 2), kw.get("dd", 28),
                         kw.get("fmt", capi.FORMAT_SH), kw.get("bd", 9), off.ctypes.data, off.ctypes.data]
    assert lib.rto_tree_create(*args(N=4)) == capi.RTO_ERR_UNSUPPORTED
    assert b"N = 2" in lib.rto_last_error()
    assert lib.rto_tree_create(*args(fmt=capi.FORMAT_SG)) == capi.RTO_ERR_UNSUPPORTED
    assert lib.rto_tree_create(*args(bd=7)) == capi.RTO_ERR_INVALID
    assert lib.rto_tree_create(*args(dd=27)) == capi.RTO_ERR_INVALID
    # (the structure check of the child array — offsets in range, no cycles, depth — runs on the device:
    #  tests/test_gpu_tree.py::test_malformed_trees_are_rejected)
    q = lambda **kw: [C.byref(h), child.ctypes.data, 1, 2, 28, capi.FORMAT_SH, 9, off.ctypes.data, off.ctypes.data,
                      data.ctypes.data, data.ctypes.data, kw.get("nq", 9), data.ctypes.data, None, kw.get("nr", 0)]
    assert lib.rto_tree_create_quantized(*q(nq=10)) == capi.RTO_ERR_INVALID   # 3*10 colour slots do not fit data_dim-1 = 27
    assert lib.rto_tree_create_quantized(*q(nq=8, nr=1)) == capi.RTO_ERR_INVALID   # retained basis without its array
    assert lib.rto_tree_read_plane(None, 0, data.ctypes.data, 0) == capi.RTO_ERR_INVALID
    assert lib.rto_context_create(C.byref(h), 0, 10) == capi.RTO_ERR_INVALID
    # pipelined-caller and band read-back entry points validate before touching the device
    nofn = capi.FRAME_RETIRED_FN(0)
    assert lib.rto_frame_sequence(None, None, 0, None, 0, 0, 0, 0, 0, nofn, None) == capi.RTO_ERR_INVALID
    slots = (C.c_void_p * 2)(None, None)
    poses = np.zeros((1, 12), np.float32)
    assert lib.rto_frame_sequence(slots, slots, 2, poses.ctypes.data, 1, 100, 0, 1, 0, nofn, None) == capi.RTO_ERR_INVALID
    assert b"frames[0] is NULL" in lib.rto_last_error()
    assert lib.rto_context_read_rows_rgba8(None, data.ctypes.data, 0, 1, None) == capi.RTO_ERR_INVALID
    assert lib.rto_context_read_image_rows(None, data.ctypes.data, 0, 1, None) == capi.RTO_ERR_INVALID
    w = np.zeros(4096, np.float16)
    assert lib.rto_net_create(C.byref(h), w.ctypes.data, w.ctypes.data, w.ctypes.data, w.ctypes.data, 8, 32, 7) == capi.RTO_ERR_UNSUPPORTED
    assert b"Kernel size == 15 not supported" in lib.rto_last_error()   # filtering.cu:362-366 message
    assert lib.rto_net_create(C.byref(h), w.ctypes.data, w.ctypes.data, w.ctypes.data, w.ctypes.data, 3, 32, 4) == capi.RTO_ERR_UNSUPPORTED


def test_render_options_json_binding(capi, tmp_path):
    from rt_octree_b200 import synthetic as S

    p = tmp_path / "opt.json"
    S.write_opt_json(str(p))
    o = capi.RenderOptions.from_json(str(p))
    assert (o.spp, o.denoise, o.step_size, o.sigma_thresh, o.background_brightness) == (6, True, 1e-4, 1e-2, 1.0)
    pod = o.pod()
    assert pod.spp == 6 and pod.denoise == 1 and pod.enable_probe == 0
    d = dict(S.REFERENCE_OPT_JSON)
    del d["stop_thresh"]  # NLOHMANN_DEFINE_TYPE_INTRUSIVE uses .at(): every key is mandatory (render_options.hpp:61-77)
    with pytest.raises(KeyError):
        capi.RenderOptions.from_json(d)
    dflt = capi.RenderOptions()
    assert dflt.spp == 1 and dflt.denoise is True


def test_data_format_parse(capi):
    assert capi.parse_data_format("SH9") == (capi.FORMAT_SH, 9)
    assert capi.parse_data_format("SH16") == (capi.FORMAT_SH, 16)
    assert capi.parse_data_format("SG25") == (capi.FORMAT_SG, 25)
    assert capi.parse_data_format("ASG8") == (capi.FORMAT_ASG, 8)
    assert capi.parse_data_format("RGBA") == (capi.FORMAT_RGBA, -1)


def test_launch_renderer_rejects_bad_spp(capi):
    o = capi.RenderOptions()
    o.spp = 5
    with pytest.raises(capi.RtoError, match="spp == 5 not supported"):
        capi.launch_renderer(None, capi.Camera(8, 8), o, None)


def test_denoiser_requires_module_path(capi):
    with pytest.raises(RuntimeError, match="No torchscript module is given to denoiser"):
        capi.Denoiser("")
