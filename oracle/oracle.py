"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/rt_oracle.c) and of the host-compiled
reference functions (oracle/_ref/libref_cpu.so, built by oracle/build_ref.sh from /root/reference in place).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_P, _I, _F, _U64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64

SEED = 20230418  # RenderContext::rng, render_context.hpp:16


def build(force: bool = False) -> str:
    """Compile oracle/rt_oracle.c (and, when /root/reference is present, oracle/_ref)."""
    so = os.path.join(HERE, "_build", "librt_oracle.so")
    src = os.path.join(HERE, "rt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.rto_oracle_render.restype = _I
        _lib.rto_oracle_render.argtypes = ([_P, _P, _I, _I, _P, _P, _F, _F, _F, _P, _I, _I, _F, _F, _I, _F, _F, _F,
                                            _U64, _U64, _I, _I, _P, _P, _P] + [_P] * 10 + [_I])
        _lib.rto_oracle_pcg32_seed.argtypes = [_U64, _U64, C.POINTER(_U64), C.POINTER(_U64)]
        _lib.rto_oracle_pcg32_advance.restype = _U64
        _lib.rto_oracle_pcg32_advance.argtypes = [_U64, _U64, _U64]
        _lib.rto_oracle_uniform_bits.argtypes = [_U64, _U64, _I, _I, _I, _P]
        _lib.rto_oracle_guidance_net.restype = _I
        _lib.rto_oracle_guidance_net.argtypes = [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I]
        _lib.rto_oracle_filter.restype = _I
        _lib.rto_oracle_filter.argtypes = [_P, _P, _P, _I, _I, _I, _P]
    return _lib


def ref_cpu_lib():
    """The reference's own rt_core.cuh host-compiled; None when oracle/_ref was never built."""
    so = os.path.join(HERE, "_ref", "libref_cpu.so")
    if not os.path.exists(so):
        return None
    l = C.CDLL(so)
    l.ref_cpu_render.restype = _I
    l.ref_cpu_render.argtypes = [_P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _F, _F, _I, _F, _F, _F, _U64, _U64, _I, _I,
                                 _P, _I]
    return l


def pcg32_seed(seed: int = SEED):
    st, inc = _U64(), _U64()
    lib().rto_oracle_pcg32_seed(seed, 1, C.byref(st), C.byref(inc))
    return st.value, inc.value


def frame_rng(frame: int, warmup: int = 100, seed: int = SEED):
    """ctx.rng when main_headless.cpp renders pose `frame`: pcg32(seed) advanced (warmup+frame) times by 2^32."""
    st, inc = pcg32_seed(seed)
    st = lib().rto_oracle_pcg32_advance(st, inc, ((warmup + frame) << 32) & ((1 << 64) - 1))
    return st, inc


def _ptr(a):
    return None if a is None else a.ctypes.data


def tree_arrays(tree: dict):
    child = np.ascontiguousarray(tree["child"].reshape(-1), dtype=np.int32)
    data = np.ascontiguousarray(tree["data"].reshape(-1)).view(np.uint16)
    data_dim = int(tree["data_dim"])
    fmt = str(tree["data_format"]) if "data_format" in tree else ("RGBA" if data_dim == 4 else "SH%d" % ((data_dim - 1) // 3))
    basis_dim = int(fmt[2:]) if fmt.startswith("SH") else -1
    offset = np.ascontiguousarray(tree["offset"], dtype=np.float32)
    scale = np.ascontiguousarray(tree["invradius3"], dtype=np.float32)
    return child, data, data_dim, basis_dim, offset, scale


def render(tree: dict, c2w12, W, H, fx, fy, spp, rng, step_size=1e-4, sigma_thresh=1e-2, background=1.0,
           ndc=(-1.0, 0.0, 0.0), pix_range=None, thresh=None, trace=True, max_seq=0, want_img=False):
    """Oracle render of one frame (or a pixel range).  Returns dict(aux=[8,H,W], img, and the trace arrays)."""
    child, data, data_dim, basis_dim, offset, scale = tree_arrays(tree)
    b, e = pix_range if pix_range else (0, W * H)
    n = e - b
    aux = np.zeros((8, H, W), np.float32)
    img = np.zeros((H, W, 4), np.float32) if want_img else None
    out = {"aux": aux, "img": img}
    tr = {}
    if trace:
        tr = dict(steps=np.zeros(n, np.uint32), term=np.zeros(n, np.int32), src_bits=np.zeros(n, np.uint32),
                  t_bits=np.zeros(n, np.uint32), leaf_hash=np.zeros(n, np.uint64), depth_sum=np.zeros(n, np.uint32),
                  n_hits=np.zeros(n, np.uint32), hit_leaf=np.zeros((n, spp), np.int32),
                  hit_cnt=np.zeros((n, spp), np.uint32),
                  leaf_seq=np.zeros((n, max_seq), np.int32) if max_seq > 0 else None)
    c2w = np.ascontiguousarray(c2w12, dtype=np.float32)
    if thresh is not None:
        thresh = np.ascontiguousarray(thresh, dtype=np.float32)
        assert thresh.shape == (n, spp)
    names = ["steps", "term", "src_bits", "t_bits", "leaf_hash", "depth_sum", "n_hits", "hit_leaf", "hit_cnt", "leaf_seq"]
    rc = lib().rto_oracle_render(child.ctypes.data, data.ctypes.data, data_dim, basis_dim, offset.ctypes.data,
                                 scale.ctypes.data, ndc[0], ndc[1], ndc[2], c2w.ctypes.data, W, H, fx, fy, spp,
                                 step_size, sigma_thresh, background, rng[0], rng[1], b, e, _ptr(thresh),
                                 aux.ctypes.data, _ptr(img), *[_ptr(tr.get(k)) for k in names], max_seq)
    if rc != 0:
        raise ValueError("spp == %d not supported." % spp)
    out.update(tr)
    return out


def ref_cpu_render(tree: dict, c2w12, W, H, fx, fy, spp, rng, step_size=1e-4, sigma_thresh=1e-2, background=1.0,
                   pix_range=None, nthreads=1):
    l = ref_cpu_lib()
    if l is None:
        raise RuntimeError("oracle/_ref/libref_cpu.so not built")
    child, data, data_dim, basis_dim, offset, scale = tree_arrays(tree)
    b, e = pix_range if pix_range else (0, W * H)
    aux = np.zeros((8, H, W), np.float32)
    c2w = np.ascontiguousarray(c2w12, dtype=np.float32)
    rc = l.ref_cpu_render(child.ctypes.data, data.ctypes.data, child.size // 8, data_dim, basis_dim,
                          offset.ctypes.data, scale.ctypes.data, c2w.ctypes.data, W, H, fx, fy, spp, step_size,
                          sigma_thresh, background, rng[0], rng[1], b, e, aux.ctypes.data, nthreads)
    if rc != 0:
        raise ValueError("spp == %d not supported." % spp)
    return aux


def uniform_bits(rng, pix_range, spp):
    b, e = pix_range
    out = np.zeros((e - b, spp), np.uint32)
    lib().rto_oracle_uniform_bits(rng[0], rng[1], b, e, spp, out.ctypes.data)
    return out


def guidance_net(aux, w, fused_bias=False):
    """aux [8,H,W] fp32, w = dict(w1,b1,w2,b2 fp16) -> (weight [L,H,W], guidance [L,H,W]).
    fused_bias=False: cuDNN/ATen GPU rounding (conv->fp16, +bias->fp16); True: single rounding (PyTorch CPU)."""
    aux = np.ascontiguousarray(aux, np.float32)
    in_ch, H, W = aux.shape
    mid = w["w1"].shape[0]
    L = w["w2"].shape[0] // 2
    wm = np.zeros((L, H, W), np.float32)
    gm = np.zeros((L, H, W), np.float32)
    arrs = [np.ascontiguousarray(w[k], np.float16).view(np.uint16) for k in ("w1", "b1", "w2", "b2")]
    rc = lib().rto_oracle_guidance_net(aux.ctypes.data, in_ch, mid, L, H, W, *[a.ctypes.data for a in arrs],
                                       wm.ctypes.data, gm.ctypes.data, int(fused_bias))
    assert rc == 0
    return wm, gm


def filtering(weight, guidance, img_in):
    """weight/guidance [L,H,W], img_in [H,W,4] -> img_out [H,W,4]  (denoiser::filtering)"""
    weight = np.ascontiguousarray(weight, np.float32)
    guidance = np.ascontiguousarray(guidance, np.float32)
    img_in = np.ascontiguousarray(img_in, np.float32)
    L, H, W = weight.shape
    out = np.zeros((H, W, 4), np.float32)
    rc = lib().rto_oracle_filter(img_in.ctypes.data, weight.ctypes.data, guidance.ctypes.data, L, H, W, out.ctypes.data)
    if rc != 0:
        raise ValueError("Kernel size == %d not supported." % (2 * L + 1))
    return out


def denoise(aux, w, fused_bias=False):
    """Denoiser::denoise on an aux buffer: GuidanceNet + filter; the noisy image is aux ch 0..2 with alpha 1."""
    wm, gm = guidance_net(aux, w, fused_bias)
    H, W = aux.shape[1:]
    img_in = np.ones((H, W, 4), np.float32)
    img_in[..., :3] = np.transpose(aux[:3], (1, 2, 0))
    return filtering(wm, gm, img_in), wm, gm
