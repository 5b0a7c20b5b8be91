"""Against the UNMODIFIED reference, run on this GPU: oracle/_ref/ref_driver links the reference's own objects
(launch_renderer, Denoiser::denoise, N3Tree loader; built by oracle/build_ref.sh) and dumps aux + final image.
 * aux alpha (k/SPP) bit-identical, aux rgb within 1e-5  -> traversal + thresholds + shading parity with the reference;
 * final denoised image within max-abs 1e-3 and PSNR >= 50 dB (north star tolerance).
Nothing here reads /root/reference; the binaries travel prebuilt."""
import os
import subprocess
import sys

import numpy as np
import pytest

from util import psnr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _run_reference(tmp, tree, poses12, W, H, fx, spp, denoise, weights, nframes, warmup):
    from rt_octree_b200 import synthetic as S
    import make_ts_module as M

    S.write_tree_npz(os.path.join(tmp, "tree.npz"), tree)
    np.ascontiguousarray(poses12, np.float32).tofile(os.path.join(tmp, "poses.bin"))
    ts = M.make_ts(weights, os.path.join(tmp, "ts_test.ts"), device="cuda")
    out = os.path.join(tmp, "out")
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([REF_DRIVER, os.path.join(tmp, "tree.npz"), os.path.join(tmp, "poses.bin"), ts, str(W), str(H),
                        repr(fx), repr(fx), str(spp), str(int(denoise)), out, str(nframes), str(warmup)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = []
    for i in range(nframes):
        aux = np.fromfile(os.path.join(out, "aux_%d.bin" % i), np.float32).reshape(8, H, W)
        img = np.fromfile(os.path.join(out, "img_%d.bin" % i), np.float32).reshape(H, W, 4)
        res.append((aux, img))
    return res, r.stdout


@pytest.mark.parametrize("spp,denoise", [(6, True), (1, False), (32, False)])
def test_against_reference_binary(capi, tmp_path, mid_tree, poses8, net_weights, spp, denoise):
    if not os.path.exists(REF_DRIVER):
        pytest.skip("oracle/_ref/ref_driver not built")
    from rt_octree_b200 import synthetic as S

    W, H = 400, 304
    fx = float(np.float32(S.blender_focal(W)))
    nframes, warmup = 3, 2
    ref, log = _run_reference(str(tmp_path), mid_tree, poses8, W, H, fx, spp, denoise, net_weights, nframes, warmup)
    print(log)
    t = capi.N3Tree(os.path.join(str(tmp_path), "tree.npz"))
    ctx = capi.RenderContext(W, H)
    cam = capi.Camera(W, H, fx, fx)
    o = capi.RenderOptions()
    o.spp, o.denoise = spp, denoise
    net = capi.Denoiser(net_weights)
    for f in range(nframes):
        cam.transform = poses8[f]
        ctx.rng_set_frame(f, warmup=warmup)
        capi.launch_renderer(t, cam, o, ctx)
        if denoise:
            net.denoise(cam, ctx)
        aux, img = ctx.read_aux(), ctx.read_image()
        raux, rimg = ref[f]
        assert raux[3].max() == 1.0
        n_alpha = int((aux[3] != raux[3]).sum())
        assert n_alpha == 0, "alpha differs from the reference at %d pixels (frame %d)" % (n_alpha, f)
        assert np.abs(aux - raux).max() < 1e-5
        d = np.abs(img - rimg)
        print("frame %d: final image max-abs %.3g psnr %.1f dB" % (f, d.max(), psnr(img[..., :3], rimg[..., :3])))
        assert np.all(img[..., 3] == 1.0) and np.all(rimg[..., 3] == 1.0)
        if denoise:
            assert psnr(img[..., :3], rimg[..., :3]) >= 50.0
            assert d.max() < 1e-3
        else:
            assert d.max() < 1e-5
