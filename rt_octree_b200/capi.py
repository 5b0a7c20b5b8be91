"""ctypes binding of librtoctree_b200.so (include/rtoctree_b200.h) + thin Python mirrors of the reference's
host classes for this path: N3Tree / Camera / RenderOptions / RenderContext / Denoiser / launch_renderer
(SURVEY.md §8b).  PyTorch is used by callers only for device buffers and streams; nothing here imports it.

There is no CPU fallback: if the shared library is missing, loading raises (run __graft_entry__.build()).
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTO_LIB", os.path.join(_HERE, "librtoctree_b200.so"))   # RTO_LIB: A/B builds while tuning

RTO_OK, RTO_ERR_INVALID, RTO_ERR_UNSUPPORTED, RTO_ERR_CUDA, RTO_ERR_NOMEM = 0, -1, -2, -3, -4
FORMAT_RGBA, FORMAT_SH, FORMAT_SG, FORMAT_ASG = 0, 1, 2, 3
SUPPORTED_SPP = (1, 2, 3, 4, 6, 8, 16, 32)  # renderer/src/cuda/volrend.cu:266-278


class RtoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("rtoctree_b200 error %d: %s" % (code, msg))
        self.code = code


class CameraPOD(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_float), ("fy", C.c_float), ("c2w", C.c_float * 12)]


class RenderOptionsPOD(C.Structure):
    _fields_ = [("step_size", C.c_float), ("sigma_thresh", C.c_float), ("stop_thresh", C.c_float),
                ("background_brightness", C.c_float), ("denoise", C.c_int), ("spp", C.c_int),
                ("enable_probe", C.c_int)]


class TreeInfoPOD(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("N", C.c_int), ("data_dim", C.c_int), ("format", C.c_int),
                ("basis_dim", C.c_int), ("max_depth", C.c_int), ("n_leaves", C.c_int64), ("node_bytes", C.c_int64),
                ("payload_bytes", C.c_int64), ("payload_stride_halfs", C.c_int), ("grid_level", C.c_int), ("n_bricks", C.c_int64),
                ("grid_bytes", C.c_int64), ("offset", C.c_float * 3),
                ("scale", C.c_float * 3), ("ndc_width", C.c_float), ("ndc_height", C.c_float), ("ndc_focal", C.c_float)]


class TracePOD(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("steps", "term", "src_bits", "t_bits", "leaf_hash", "depth_sum", "n_hits",
                                          "n_loads", "hit_leaf", "hit_cnt", "leaf_seq", "thresh")] + [("max_seq", C.c_int),
                                                                                                      ("marcher", C.c_int)]


class FrameDescPOD(C.Structure):
    _fields_ = [("tree", C.c_void_p), ("net", C.c_void_p), ("opt", RenderOptionsPOD), ("fx", C.c_float), ("fy", C.c_float),
                ("host_rgba8", C.c_void_p), ("host_image", C.c_void_p), ("host_aux", C.c_void_p)]


FRAME_RETIRED_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.c_int)   # rto_frame_retired_fn(user, frame_index, slot)


EXPORTS = [  # every symbol include/rtoctree_b200.h declares (tests/test_abi.py checks the library exports them all)
    "rto_last_error", "rto_abi_version", "rto_set_device", "rto_device_count", "rto_synchronize", "rto_render_options_default",
    "rto_tree_create", "rto_tree_create_quantized", "rto_tree_read_plane", "rto_tree_set_ndc", "rto_tree_get_info", "rto_tree_destroy",
    "rto_context_create", "rto_context_destroy", "rto_context_aux", "rto_context_image", "rto_context_rng_seed",
    "rto_context_rng_advance", "rto_context_rng_set_frame", "rto_context_rng_get", "rto_context_read_aux", "rto_context_write_aux", "rto_context_read_image_rgba8",
    "rto_context_read_image", "rto_render", "rto_render_rect", "rto_render_trace",
    "rto_net_create", "rto_net_destroy", "rto_net_set_impl", "rto_net_set_bias_mode", "rto_denoise", "rto_denoise_rows", "rto_net_forward",
    "rto_filter", "rto_filter_forward_save", "rto_filter_backward", "rto_timer_enable", "rto_timer_reset", "rto_timer_record", "rto_timer_report", "rto_launch_count",
    "rto_context_image_rgba8", "rto_context_read_rows_rgba8", "rto_context_read_image_rows", "rto_stream_create", "rto_stream_destroy", "rto_host_alloc", "rto_host_free",
    "rto_frame_create", "rto_frame_launch", "rto_frame_launch_indexed", "rto_frame_sequence", "rto_frame_destroy",
    "rto_context_set_image_target", "rto_context_mark_image_written", "rto_peer_enable", "rto_ipc_export", "rto_ipc_open",
    "rto_ipc_close", "rto_event_create", "rto_event_create_timed", "rto_event_elapsed_ms", "rto_event_record", "rto_stream_wait_event",
    "rto_event_destroy",
]

_lib = None


def load(path: str = LIB_PATH):
    """dlopen the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RtoError(RTO_ERR_CUDA, "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    L = C.CDLL(path)
    P, I, F = C.c_void_p, C.c_int, C.c_float
    L.rto_last_error.restype = C.c_char_p
    L.rto_abi_version.restype = I
    L.rto_launch_count.restype = C.c_int64
    L.rto_set_device.argtypes = [I]
    L.rto_device_count.argtypes = [C.POINTER(I)]
    L.rto_synchronize.argtypes = [P]
    L.rto_render_options_default.argtypes = [C.POINTER(RenderOptionsPOD)]
    L.rto_render_options_default.restype = None
    L.rto_tree_create.argtypes = [C.POINTER(P), P, P, C.c_int64, I, I, I, I, P, P]
    L.rto_tree_create_quantized.argtypes = [C.POINTER(P), P, C.c_int64, I, I, I, I, P, P, P, P, I, P, P, I]
    L.rto_tree_read_plane.argtypes = [P, I, P, C.c_size_t]
    L.rto_tree_set_ndc.argtypes = [P, F, F, F]
    L.rto_tree_get_info.argtypes = [P, C.POINTER(TreeInfoPOD)]
    L.rto_tree_destroy.argtypes = [P]
    L.rto_tree_destroy.restype = None
    L.rto_context_create.argtypes = [C.POINTER(P), I, I]
    L.rto_context_destroy.argtypes = [P]
    L.rto_context_destroy.restype = None
    L.rto_context_aux.argtypes = [P]
    L.rto_context_aux.restype = P
    L.rto_context_image.argtypes = [P]
    L.rto_context_image.restype = P
    L.rto_context_rng_seed.argtypes = [P, C.c_uint64]
    L.rto_context_rng_advance.argtypes = [P, C.c_int64]
    L.rto_context_rng_set_frame.argtypes = [P, C.c_int64, C.c_int64]
    L.rto_context_rng_get.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.rto_context_read_aux.argtypes = [P, P, P]
    L.rto_context_write_aux.argtypes = [P, P, P]
    L.rto_context_read_image_rgba8.argtypes = [P, P, P]
    L.rto_context_read_image.argtypes = [P, P, P]
    L.rto_context_read_rows_rgba8.argtypes = [P, P, I, I, P]
    L.rto_context_read_image_rows.argtypes = [P, P, I, I, P]
    L.rto_render.argtypes = [P, P, C.POINTER(CameraPOD), C.POINTER(RenderOptionsPOD), P]
    L.rto_render_rect.argtypes = [P, P, C.POINTER(CameraPOD), C.POINTER(RenderOptionsPOD), I, I, I, I, P]
    L.rto_render_trace.argtypes = [P, P, C.POINTER(CameraPOD), C.POINTER(RenderOptionsPOD), C.POINTER(TracePOD), P]
    L.rto_net_create.argtypes = [C.POINTER(P), P, P, P, P, I, I, I]
    L.rto_net_destroy.argtypes = [P]
    L.rto_net_destroy.restype = None
    L.rto_net_set_impl.argtypes = [P, I]
    L.rto_net_set_bias_mode.argtypes = [P, I]
    L.rto_denoise.argtypes = [P, P, P]
    L.rto_denoise_rows.argtypes = [P, P, I, I, P]
    L.rto_net_forward.argtypes = [P, P, I, I, P, P, P]
    L.rto_filter.argtypes = [P, P, P, I, I, I, P, P]
    L.rto_filter_forward_save.argtypes = [P, P, P, I, I, I, P, P, P, P, P]
    L.rto_filter_backward.argtypes = [P, P, P, P, P, P, P, I, I, I, P, P, P]
    L.rto_context_image_rgba8.argtypes = [P]
    L.rto_context_image_rgba8.restype = P
    L.rto_stream_create.argtypes = [C.POINTER(P)]
    L.rto_stream_destroy.argtypes = [P]
    L.rto_host_alloc.argtypes = [C.POINTER(P), C.c_size_t]
    L.rto_host_free.argtypes = [P]
    L.rto_frame_create.argtypes = [C.POINTER(P), P, C.POINTER(FrameDescPOD)]
    L.rto_frame_launch.argtypes = [P, C.POINTER(C.c_float * 12), P]
    L.rto_frame_launch_indexed.argtypes = [P, C.POINTER(C.c_float * 12), C.c_int64, C.c_int64, P]
    L.rto_frame_sequence.argtypes = [C.POINTER(P), C.POINTER(P), I, P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, I, FRAME_RETIRED_FN, P]
    L.rto_frame_destroy.argtypes = [P]
    L.rto_frame_destroy.restype = None
    L.rto_context_set_image_target.argtypes = [P, P, P]
    L.rto_context_mark_image_written.argtypes = [P, I]
    L.rto_peer_enable.argtypes = [I]
    L.rto_ipc_export.argtypes = [P, P]
    L.rto_ipc_open.argtypes = [P, C.POINTER(P)]
    L.rto_ipc_close.argtypes = [P]
    L.rto_event_create.argtypes = [C.POINTER(P)]
    L.rto_event_create_timed.argtypes = [C.POINTER(P)]
    L.rto_event_elapsed_ms.argtypes = [P, P, C.POINTER(F)]
    L.rto_event_record.argtypes = [P, P]
    L.rto_stream_wait_event.argtypes = [P, P]
    L.rto_event_destroy.argtypes = [P]
    L.rto_timer_enable.argtypes = [P, I]
    L.rto_timer_reset.argtypes = [P]
    L.rto_timer_record.argtypes = [P, I]
    L.rto_timer_report.argtypes = [P, C.POINTER(F * 3), C.POINTER(I)]
    _lib = L
    return L


def _check(rc):
    if rc != RTO_OK:
        raise RtoError(rc, load().rto_last_error().decode())


def launch_count() -> int:
    return int(load().rto_launch_count())


def device_count() -> int:
    n = C.c_int(0)
    _check(load().rto_device_count(C.byref(n)))
    return int(n.value)


def set_device(device: int):
    _check(load().rto_set_device(device))


# ----------------------------------------------------------------------------------------------- RenderOptions
class RenderOptions:
    """volrend::RenderOptions (include/volrend/render_options.hpp:13-78).  `from_json` requires all 11 keys of the
    intrusive JSON binding (:61-77; nlohmann `.at()` throws on a missing key)."""
    JSON_KEYS = ("step_size", "sigma_thresh", "stop_thresh", "background_brightness", "show_grid", "grid_max_depth",
                 "enable_probe", "probe", "probe_disp_size", "denoise", "spp")

    def __init__(self):
        self.step_size = 1e-4
        self.sigma_thresh = 1e-2
        self.stop_thresh = 1e-2
        self.background_brightness = 1.0
        self.show_grid = False
        self.grid_max_depth = 4
        self.enable_probe = False
        self.probe = [0.0, 0.0, 1.0]
        self.probe_disp_size = 100
        self.denoise = True
        self.spp = 1

    @classmethod
    def from_json(cls, path_or_dict):
        d = path_or_dict
        if not isinstance(d, dict):
            with open(path_or_dict) as f:
                d = json.load(f)
        o = cls()
        for k in cls.JSON_KEYS:
            if k not in d:
                raise KeyError("key '%s' not found" % k)
            setattr(o, k, d[k])
        return o

    def pod(self) -> RenderOptionsPOD:
        return RenderOptionsPOD(float(self.step_size), float(self.sigma_thresh), float(self.stop_thresh),
                                float(self.background_brightness), int(bool(self.denoise)), int(self.spp),
                                int(bool(self.enable_probe)))


# ------------------------------------------------------------------------------------------------------ Camera
class Camera:
    """volrend::Camera as the headless driver uses it (width, height, fx, fy, transform; camera.hpp:16-68)."""

    def __init__(self, width=256, height=256, fx=1111.11, fy=-1.0):
        self.width, self.height = int(width), int(height)
        self.fx = 1111.11 if fx < 0 else float(fx)
        self.fy = self.fx if fy < 0 else float(fy)
        self.transform = np.zeros(12, np.float32)  # column-major 4x3

    def pod(self) -> CameraPOD:
        p = CameraPOD(self.width, self.height, self.fx, self.fy)
        t = np.ascontiguousarray(self.transform, np.float32).reshape(12)
        for i in range(12):
            p.c2w[i] = float(t[i])
        return p


# ------------------------------------------------------------------------------------------------------ N3Tree
def parse_data_format(s: str):
    """DataFormat::parse (src/n3tree.cpp:55-78)."""
    i = 0
    while i < len(s) and s[i].isalpha():
        i += 1
    if i < len(s):
        name, dim = s[:i], int(s[i:]) if s[i:].lstrip("-").isdigit() else 0
        fmt = {"ASG": FORMAT_ASG, "SG": FORMAT_SG, "SH": FORMAT_SH}.get(name, FORMAT_RGBA)
        return fmt, dim
    return FORMAT_RGBA, -1


class N3Tree:
    """volrend::N3Tree: open a tree.npz (or take the arrays) and load it onto the GPU."""

    def __init__(self, path_or_arrays):
        self._h = C.c_void_p()
        if isinstance(path_or_arrays, (str, os.PathLike)):
            path = str(path_or_arrays)
            if not os.path.exists(path):
                raise FileNotFoundError("Can't load because file does not exist: %s" % path)
            z = np.load(path)
            arrays = {k: z[k] for k in z.files}
        else:
            arrays = path_or_arrays
        self._load(arrays)

    def _load(self, z):
        # N3Tree::load_npz, src/n3tree.cpp:228-362
        self.data_dim = int(np.asarray(z["data_dim"]).reshape(-1)[0])
        if "data_format" in z:
            self.format, self.basis_dim = parse_data_format(str(np.asarray(z["data_format"]).reshape(-1)[0]))
        elif self.data_dim == 4:
            self.format, self.basis_dim = FORMAT_RGBA, -1
        else:
            self.format, self.basis_dim = FORMAT_SH, (self.data_dim - 1) // 3
        if "invradius3" in z:
            self.scale = np.asarray(z["invradius3"], np.float32).reshape(3).copy()
        else:
            self.scale = np.full(3, np.float32(np.asarray(z["invradius"]).reshape(-1)[0]), np.float32)
        self.offset = np.asarray(z["offset"], np.float32).reshape(3).copy()
        child = np.ascontiguousarray(z["child"], np.int32)
        if child.ndim != 4 or child.shape[0] < 1 or child.shape[2:] != (child.shape[1],) * 2:
            raise ValueError("malformed tree.npz: child must be int32 [cap,N,N,N]")
        self.N = int(child.shape[1])
        self.capacity = int(child.shape[0])
        cell = (self.capacity, self.N, self.N, self.N)

        def expect(name, arr, shape):
            # sizes are derived from `child`; the C ABI receives raw pointers and cannot check the lengths behind them
            if tuple(arr.shape) != tuple(shape):
                raise ValueError("malformed tree.npz: '%s' has shape %s, expected %s" % (name, tuple(arr.shape), tuple(shape)))

        L = load()
        off, sc = self.offset.ctypes.data, self.scale.ctypes.data
        if "quant_colors" in z:
            # codebook files go to the GPU as they are; the gather runs there (rto_tree_create_quantized)
            qc = np.asarray(z["quant_colors"])
            if qc.dtype != np.float16:
                raise ValueError("codebook must be stored in half precision")
            qm = np.ascontiguousarray(z["quant_map"], np.uint16)
            if qc.shape[0] != qm.shape[0]:
                raise ValueError("codebook and map basis numbers does not match")
            qc = np.ascontiguousarray(qc)
            sigma = np.ascontiguousarray(z["sigma"], np.float16)
            ret = np.ascontiguousarray(z["data_retained"], np.float16) if "data_retained" in z else None
            expect("quant_colors", qc, (qm.shape[0], 65536, 3))
            expect("quant_map", qm, (qm.shape[0],) + cell)
            expect("sigma", sigma, cell)
            if ret is not None:
                expect("data_retained", ret, (ret.shape[0],) + cell + (3,))
            _check(L.rto_tree_create_quantized(C.byref(self._h), child.ctypes.data, self.capacity, self.N, self.data_dim,
                                               self.format, self.basis_dim, off, sc, qc.ctypes.data, qm.ctypes.data,
                                               int(qm.shape[0]), sigma.ctypes.data,
                                               ret.ctypes.data if ret is not None else None,
                                               int(ret.shape[0]) if ret is not None else 0))
        else:
            data = z["data"]
            if data.dtype != np.float16:
                raise ValueError("data must be stored in half precision")
            data = np.ascontiguousarray(data)
            expect("data", data, cell + (self.data_dim,))
            _check(L.rto_tree_create(C.byref(self._h), child.ctypes.data, data.ctypes.data, self.capacity, self.N,
                                     self.data_dim, self.format, self.basis_dim, off, sc))

    PLANES = {"nodes": (0, np.uint32), "payload": (1, np.float16), "grid_top": (2, np.uint32), "grid_bricks": (3, np.uint32),
              "grid_bricks8": (4, np.uint8), "grid_leaf_top": (5, np.uint32), "grid_leaf_bricks": (6, np.uint32),
              "grid_march_top": (7, np.uint32)}

    def read_plane(self, name):
        """Device plane copied back to the host (rto_tree_read_plane): nodes / payload / grid_top / grid_bricks."""
        which, dt = self.PLANES[name]
        i = self.info
        nbytes = {"nodes": i.node_bytes, "payload": i.payload_bytes,
                  "grid_top": (4 << (3 * i.grid_level)) if i.grid_level else 0,
                  "grid_bricks": i.n_bricks * 2048 if i.grid_level else 0,
                  "grid_bricks8": i.n_bricks * 512 if i.grid_level else 0,
                  "grid_leaf_top": (4 << (3 * i.grid_level)) if i.grid_level else 0,
                  "grid_leaf_bricks": i.n_bricks * 2048 if i.grid_level else 0,
                  "grid_march_top": (4 << (3 * i.grid_level)) if i.grid_level else 0}[name]
        out = np.empty(nbytes // np.dtype(dt).itemsize, dt)
        _check(load().rto_tree_read_plane(self._h, which, out.ctypes.data, nbytes))
        return out

    def set_ndc(self, width, height, focal):
        _check(load().rto_tree_set_ndc(self._h, float(width), float(height), float(focal)))

    @property
    def info(self) -> TreeInfoPOD:
        i = TreeInfoPOD()
        _check(load().rto_tree_get_info(self._h, C.byref(i)))
        return i

    def close(self):
        if self._h:
            load().rto_tree_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decode_quantized(z, capacity, N, data_dim):
    """Median-cut codebook decode (src/n3tree.cpp:279-340; producer renderer/scripts/compress_octree.py:68-119)."""
    qc = np.asarray(z["quant_colors"])
    if qc.dtype != np.float16:
        raise ValueError("codebook must be stored in half precision")
    qm = np.asarray(z["quant_map"])
    n_basis_q = qm.shape[0]
    if qc.shape[0] != n_basis_q:
        raise ValueError("codebook and map basis numbers does not match")
    retained = np.asarray(z["data_retained"]) if "data_retained" in z else None
    n_ret = retained.shape[0] if retained is not None else 0
    n_basis = n_basis_q + n_ret
    n_child = capacity * N ** 3
    data = np.zeros((n_child, data_dim), np.float16)
    qm = qm.reshape(n_basis_q, n_child)
    for j in range(n_basis_q):
        cols = qc[j].reshape(-1, 3)[qm[j].astype(np.int64)]          # [n_child,3]
        for k in range(3):
            data[:, j + n_ret + k * n_basis] = cols[:, k]
    data[:, data_dim - 1] = np.asarray(z["sigma"]).reshape(n_child)
    if retained is not None:
        r = retained.reshape(n_ret, n_child, 3)
        for j in range(n_ret):
            for k in range(3):
                data[:, j + k * n_basis] = r[j, :, k]
    return np.ascontiguousarray(data)


# ----------------------------------------------------------------------------------------------- RenderContext
class RenderContext:
    """volrend::RenderContext (render_context.hpp:14-120): aux buffer, output image, pcg32 state, stage timer."""
    CHANNELS = 8

    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)
        self._h = C.c_void_p()
        _check(load().rto_context_create(C.byref(self._h), self.width, self.height))

    @property
    def aux_ptr(self) -> int:
        return load().rto_context_aux(self._h)

    @property
    def image_ptr(self) -> int:
        return load().rto_context_image(self._h)

    def rng_seed(self, seed=20230418):
        _check(load().rto_context_rng_seed(self._h, seed))

    def rng_advance(self, delta=1 << 32):
        _check(load().rto_context_rng_advance(self._h, delta))

    def rng_set_frame(self, frame, warmup=100):
        _check(load().rto_context_rng_set_frame(self._h, warmup, frame))

    def rng_get(self):
        s, i = C.c_uint64(), C.c_uint64()
        _check(load().rto_context_rng_get(self._h, C.byref(s), C.byref(i)))
        return s.value, i.value

    def read_aux(self, host: np.ndarray = None, stream=0, sync=True) -> np.ndarray:
        if host is None:
            host = np.empty((8, self.height, self.width), np.float32)
        _check(load().rto_context_read_aux(self._h, host.ctypes.data, C.c_void_p(stream)))
        if sync:
            _cuda_sync()
        return host

    def read_image_rgba8(self, host: np.ndarray = None, stream=0, sync=True) -> np.ndarray:
        """Final image as uint8 [H][W][4], `(uint8)(v * 255)` on the device (the bytes the reference CLI writes to PNG)."""
        if host is None:
            host = np.empty((self.height, self.width, 4), np.uint8)
        _check(load().rto_context_read_image_rgba8(self._h, host.ctypes.data, C.c_void_p(stream)))
        if sync:
            _cuda_sync()
        return host

    def read_rows(self, host8=None, host_image=None, rows=(0, 0), stream=0, sync=True):
        """Rows [y0, y1) of the RGBA8 copy / the float4 image into the same rows of full-frame host arrays (the band
        read-back of a tile split whose GPUs deliver their own rows: rto_context_read_rows_rgba8 / _read_image_rows)."""
        if host8 is not None:
            _check(load().rto_context_read_rows_rgba8(self._h, host8.ctypes.data, int(rows[0]), int(rows[1]), C.c_void_p(stream)))
        if host_image is not None:
            _check(load().rto_context_read_image_rows(self._h, host_image.ctypes.data, int(rows[0]), int(rows[1]), C.c_void_p(stream)))
        if sync:
            _cuda_sync()

    def write_aux(self, host: np.ndarray, stream=0, sync=True):
        """Upload a stored guidance buffer (fp32 [8][H][W], the `buf_*.bin` layout)."""
        host = np.ascontiguousarray(host, np.float32)
        if host.shape != (8, self.height, self.width):
            raise ValueError("aux buffer must be [8, %d, %d]" % (self.height, self.width))
        _check(load().rto_context_write_aux(self._h, host.ctypes.data, C.c_void_p(stream)))
        if sync:
            _cuda_sync()

    def read_image(self, host: np.ndarray = None, stream=0, sync=True) -> np.ndarray:
        if host is None:
            host = np.empty((self.height, self.width, 4), np.float32)
        _check(load().rto_context_read_image(self._h, host.ctypes.data, C.c_void_p(stream)))
        if sync:
            _cuda_sync()
        return host

    @property
    def image_rgba8_ptr(self) -> int:
        return load().rto_context_image_rgba8(self._h)

    def set_image_target(self, image_ptr=None, rgba8_ptr=None):
        """Tile split: the kernels that produce the final image store it at these full-frame DEVICE pointers (possibly in a
        peer GPU's memory) instead of this context's own buffers; None restores them."""
        _check(load().rto_context_set_image_target(self._h, C.c_void_p(image_ptr), C.c_void_p(rgba8_ptr)))

    def mark_image_written(self, rgba8_too=True):
        _check(load().rto_context_mark_image_written(self._h, int(bool(rgba8_too))))

    def timer_enable(self, on=True):
        _check(load().rto_timer_enable(self._h, int(on)))

    def timer_reset(self):
        _check(load().rto_timer_reset(self._h))

    def timer_record(self, denoise: bool):
        _check(load().rto_timer_record(self._h, int(denoise)))

    def timer_report(self):
        ms = (C.c_float * 3)()
        n = C.c_int()
        _check(load().rto_timer_report(self._h, C.byref(ms), C.byref(n)))
        return [ms[0], ms[1], ms[2]], n.value

    def close(self):
        if self._h:
            load().rto_context_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _cuda_sync(stream=0):
    _check(load().rto_synchronize(C.c_void_p(stream)))


def synchronize(stream=0):
    _cuda_sync(stream)


def launch_renderer(tree: N3Tree, cam: Camera, options: RenderOptions, ctx: RenderContext, stream=0, rect=None,
                    trace: TracePOD = None):
    """volrend::launch_renderer(tree, cam, options, ctx, stream, offscreen=true)."""
    if options.spp not in SUPPORTED_SPP:
        raise RtoError(RTO_ERR_UNSUPPORTED, "spp == %d not supported." % options.spp)
    L = load()
    cp, op = cam.pod(), options.pod()
    s = C.c_void_p(stream)
    if trace is not None:
        _check(L.rto_render_trace(ctx._h, tree._h, C.byref(cp), C.byref(op), C.byref(trace), s))
    elif rect is not None:
        _check(L.rto_render_rect(ctx._h, tree._h, C.byref(cp), C.byref(op), *[int(v) for v in rect], s))
    else:
        _check(L.rto_render(ctx._h, tree._h, C.byref(cp), C.byref(op), s))


# ---------------------------------------------------------------------------------------------------- Denoiser
class Denoiser:
    """volrend::Denoiser (denoiser.hpp:11-21).  The reference loads a TorchScript file; here the same four fp16
    tensors are read from the raw export made once by tools/make_ts_module.py --export (`<name>.npz` with w1,b1,w2,b2)
    or passed as arrays.  An empty path raises, like the reference (denoiser.cpp:13-16)."""

    def __init__(self, weights):
        self._h = C.c_void_p()
        if isinstance(weights, (str, os.PathLike)):
            if not str(weights):
                raise RuntimeError("No torchscript module is given to denoiser.")
            z = np.load(str(weights))
            weights = {k: z[k] for k in ("w1", "b1", "w2", "b2")}
        w = {k: np.ascontiguousarray(weights[k], np.float16) for k in ("w1", "b1", "w2", "b2")}
        mid, in_ch = w["w1"].shape[0], w["w1"].shape[1]
        levels = w["w2"].shape[0] // 2
        self.in_ch, self.mid_ch, self.levels = in_ch, mid, levels
        _check(load().rto_net_create(C.byref(self._h), w["w1"].ctypes.data, w["b1"].ctypes.data, w["w2"].ctypes.data,
                                     w["b2"].ctypes.data, in_ch, mid, levels))

    def set_impl(self, impl: int):
        _check(load().rto_net_set_impl(self._h, impl))

    def set_bias_mode(self, fused: bool):
        _check(load().rto_net_set_bias_mode(self._h, int(fused)))

    def denoise(self, cam: Camera, ctx: RenderContext, stream=0, rows=None):
        if rows is None:
            _check(load().rto_denoise(ctx._h, self._h, C.c_void_p(stream)))
        else:
            _check(load().rto_denoise_rows(ctx._h, self._h, int(rows[0]), int(rows[1]), C.c_void_p(stream)))

    def forward(self, aux_ptr: int, width: int, height: int, weight_ptr: int, guidance_ptr: int, stream=0):
        _check(load().rto_net_forward(self._h, C.c_void_p(aux_ptr), width, height, C.c_void_p(weight_ptr),
                                      C.c_void_p(guidance_ptr), C.c_void_p(stream)))

    def close(self):
        if self._h:
            load().rto_net_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Pinned host memory from the library (rto_host_alloc) viewed as a numpy array: destination of async read-backs."""

    def __init__(self, shape, dtype):
        self._p = C.c_void_p()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _check(load().rto_host_alloc(C.byref(self._p), n))
        self.array = np.ctypeslib.as_array((C.c_ubyte * n).from_address(self._p.value)).view(dtype).reshape(shape)
        self.ptr = self._p.value

    def close(self):
        if self._p:
            self.array = None
            load().rto_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def peer_enable(peer_device: int):
    _check(load().rto_peer_enable(int(peer_device)))


def ipc_export(dev_ptr: int) -> bytes:
    h = (C.c_ubyte * 64)()
    _check(load().rto_ipc_export(C.c_void_p(dev_ptr), h))
    return bytes(h)


def ipc_open(handle: bytes) -> int:
    h = (C.c_ubyte * 64)(*handle)
    p = C.c_void_p()
    _check(load().rto_ipc_open(h, C.byref(p)))
    return p.value


def ipc_close(dev_ptr: int):
    _check(load().rto_ipc_close(C.c_void_p(dev_ptr)))


def stream_create() -> int:
    s = C.c_void_p()
    _check(load().rto_stream_create(C.byref(s)))
    return s.value


def stream_destroy(stream: int):
    _check(load().rto_stream_destroy(C.c_void_p(stream)))


class Frame:
    """rto_frame: render -> denoise -> read-backs of one context captured as a CUDA graph, one launch per frame.
    `rgba8` / `image` / `aux` are optional PinnedBuffer destinations written by every launch."""

    def __init__(self, ctx: RenderContext, tree: N3Tree, net, options: RenderOptions, fx, fy, rgba8=None, image=None, aux=None):
        self._h = C.c_void_p()
        self._keep = (ctx, tree, net, rgba8, image, aux)
        d = FrameDescPOD(tree._h, net._h if net is not None else None, options.pod(), float(fx), float(fy),
                         rgba8.ptr if rgba8 is not None else None, image.ptr if image is not None else None,
                         aux.ptr if aux is not None else None)
        _check(load().rto_frame_create(C.byref(self._h), ctx._h, C.byref(d)))
        self._launch_indexed = load().rto_frame_launch_indexed

    @staticmethod
    def pose_array(c2w12):
        """A camera transform as the ctypes array the launch calls take (build once per pose, reuse per frame)."""
        return (C.c_float * 12)(*[float(v) for v in np.asarray(c2w12, np.float32).reshape(12)])

    def launch(self, c2w12, stream=0):
        m = c2w12 if isinstance(c2w12, C.c_float * 12) else self.pose_array(c2w12)
        _check(load().rto_frame_launch(self._h, C.byref(m), C.c_void_p(stream)))

    def launch_indexed(self, c2w12, frame, warmup=100, stream=0):
        """ctx.rng for pose `frame` of the job (rng_set_frame) + launch, one library call."""
        m = c2w12 if isinstance(c2w12, C.c_float * 12) else self.pose_array(c2w12)
        rc = self._launch_indexed(self._h, m, warmup, frame, stream)
        if rc != RTO_OK:
            _check(rc)

    def close(self):
        if self._h:
            load().rto_frame_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrameSequence:
    """rto_frame_sequence: the host loop of a pipelined driver in one library call.  `frames[k]` runs on `streams[k]`
    (raw stream handles); `poses` is a float32 [n][12] array kept alive here.  run(first, count) issues frames
    first .. first+count-1 (frame i on slot i % n_slots, pose i % n); `retired(frame_index, slot)` is called when a frame's
    host buffers are complete, before the slot is reused."""

    def __init__(self, frames, streams, poses, warmup=100, retired=None):
        self._frames = list(frames)
        n = len(self._frames)
        self._fh = (C.c_void_p * n)(*[f._h for f in self._frames])
        self._st = (C.c_void_p * n)(*[C.c_void_p(s) for s in streams])
        self._poses = np.ascontiguousarray(poses, np.float32).reshape(-1, 12)
        self._warmup = int(warmup)
        self._cb = FRAME_RETIRED_FN((lambda user, idx, slot: retired(int(idx), int(slot))) if retired else 0)
        self._run = load().rto_frame_sequence

    def run(self, first, count, drain=False):
        rc = self._run(self._fh, self._st, len(self._frames), self._poses.ctypes.data, self._poses.shape[0], self._warmup,
                       int(first), int(count), 1 if drain else 0, self._cb, None)
        if rc != RTO_OK:
            _check(rc)


def filtering(weight_ptr, guidance_ptr, img_in_ptr, levels, width, height, img_out_ptr, stream=0):
    """denoiser::filtering(stream, weight_map, guidance_map, img_in, img_out) on device pointers."""
    _check(load().rto_filter(C.c_void_p(weight_ptr), C.c_void_p(guidance_ptr), C.c_void_p(img_in_ptr), levels, width,
                             height, C.c_void_p(img_out_ptr), C.c_void_p(stream)))
