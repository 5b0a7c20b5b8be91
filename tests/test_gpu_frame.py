"""GPU tests of the pipelined-caller part of the C ABI: rto_frame (one CUDA-graph launch per frame), pinned buffers, the
RGBA8 copy written by the producing kernel's own epilogue, and the image-target redirection used by the tile split."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rig(capi, tree, W, H, spp, denoise, weights):
    from rt_octree_b200 import synthetic as S

    fx = S.blender_focal(W)
    t = capi.N3Tree(tree)
    cam = capi.Camera(W, H, fx, fx)
    opt = capi.RenderOptions()
    opt.spp, opt.denoise = spp, denoise
    net = capi.Denoiser(weights) if denoise else None
    return t, cam, opt, net


@pytest.mark.parametrize("spp,denoise", [(6, True), (1, False), (16, True)])
def test_frame_graph_equals_separate_launches(capi, mid_tree, poses8, net_weights, spp, denoise):
    """rto_frame_launch (render -> net -> filter -> read-backs captured once, the render node's argument block replaced per
    frame) produces exactly the buffers of rto_render + rto_denoise + rto_context_read_*, for changing poses and rng states,
    interleaved with ordinary launches on the same context."""
    W, H = 232, 168
    t, cam, opt, net = _rig(capi, mid_tree, W, H, spp, denoise, net_weights)
    ref = capi.RenderContext(W, H)
    ctx = capi.RenderContext(W, H)
    b8, bimg, baux = capi.PinnedBuffer((H, W, 4), np.uint8), capi.PinnedBuffer((H, W, 4), np.float32), capi.PinnedBuffer((8, H, W), np.float32)
    stream = capi.stream_create()
    fr = capi.Frame(ctx, t, net, opt, cam.fx, cam.fy, rgba8=b8, image=bimg, aux=baux)
    launches0 = capi.launch_count()
    for f in (3, 0, 7, 3):
        ctx.rng_set_frame(f)
        fr.launch(poses8[f], stream=stream)
        capi.synchronize(stream)
        cam.transform = poses8[f]
        ref.rng_set_frame(f)
        capi.launch_renderer(t, cam, opt, ref)
        if denoise:
            net.denoise(cam, ref)
        want_img, want_aux, want8 = ref.read_image(), ref.read_aux(), ref.read_image_rgba8()
        assert np.array_equal(baux.array, want_aux)
        assert np.array_equal(bimg.array, want_img)
        assert np.array_equal(b8.array, want8)
        assert np.array_equal(want8, (want_img * np.float32(255)).astype(np.int32).astype(np.uint8))
        # the context's own buffers hold the frame too, and ordinary launches still work on it
        assert np.array_equal(ctx.read_image(), want_img)
    assert capi.launch_count() - launches0 >= 4 * (3 if denoise else 1)      # graph launches are counted kernel by kernel
    # a frame needs a net when the options say denoise
    if denoise:
        with pytest.raises(capi.RtoError, match="no net"):
            capi.Frame(ctx, t, None, opt, cam.fx, cam.fy)
    fr.close()
    capi.stream_destroy(stream)


def test_rgba8_written_by_the_producing_kernel(capi, mid_tree, poses8, net_weights):
    """After the RGBA8 buffer exists, rto_denoise's filter (and rto_render with denoise off) write it themselves: reading it
    back launches nothing, and the bytes equal the stand-alone conversion."""
    W, H = 200, 152
    for denoise in (True, False):
        t, cam, opt, net = _rig(capi, mid_tree, W, H, 6, denoise, net_weights)
        ctx = capi.RenderContext(W, H)
        cam.transform = poses8[2]
        ctx.rng_set_frame(2)
        capi.launch_renderer(t, cam, opt, ctx)
        if denoise:
            net.denoise(cam, ctx)
        n0 = capi.launch_count()
        first = ctx.read_image_rgba8().copy()            # buffer did not exist: stand-alone conversion kernel
        assert capi.launch_count() == n0 + 1
        ctx.rng_set_frame(5)
        cam.transform = poses8[5]
        capi.launch_renderer(t, cam, opt, ctx)
        if denoise:
            net.denoise(cam, ctx)
        n1 = capi.launch_count()
        second = ctx.read_image_rgba8()
        assert capi.launch_count() == n1                 # written by the filter / render epilogue: no extra launch
        img = ctx.read_image()
        assert np.array_equal(second, (img * np.float32(255)).astype(np.int32).astype(np.uint8))
        assert not np.array_equal(first, second)
        # bands do not claim a current full-frame copy: the conversion kernel runs again
        if denoise:
            net.denoise(cam, ctx, rows=(10, 50))
            n2 = capi.launch_count()
            third = ctx.read_image_rgba8()
            assert capi.launch_count() == n2 + 1 and np.array_equal(third, second)


def test_image_target_redirects_the_final_stores(capi, mid_tree, poses8, net_weights):
    """rto_context_set_image_target: two contexts denoise two bands of one frame, both storing into the first context's
    image and RGBA8 copy (what the tile split does across GPUs); the assembled frame equals the full-frame render."""
    from rt_octree_b200 import sharding as SH

    W, H = 216, 160
    t, cam, opt, net = _rig(capi, mid_tree, W, H, 6, True, net_weights)
    full = capi.RenderContext(W, H)
    cam.transform = poses8[6]
    full.rng_set_frame(6)
    capi.launch_renderer(t, cam, opt, full)
    net.denoise(cam, full)
    want, want8 = full.read_image(), full.read_image_rgba8()
    a, b = capi.RenderContext(W, H), capi.RenderContext(W, H)
    b.set_image_target(a.image_ptr, a.image_rgba8_ptr)
    for c, band in ((a, (0, 77)), (b, (77, H))):
        c.rng_set_frame(6)
        y0, y1 = SH.render_rows_for_band(band, H, True, net.levels)
        capi.launch_renderer(t, cam, opt, c, rect=(0, y0, W, y1))
        net.denoise(cam, c, rows=band)
    capi.synchronize()
    a.mark_image_written(True)
    n0 = capi.launch_count()
    assert np.array_equal(a.read_image(), want) and np.array_equal(a.read_image_rgba8(), want8)
    assert capi.launch_count() == n0
    assert np.all(b.read_image()[77:] == 0)           # the second context's own image was not touched
    # the rows were stored into `a`: a band read-back from `b` must refuse
    with pytest.raises(capi.RtoError, match="image target"):
        b.read_rows(host8=np.empty((H, W, 4), np.uint8), rows=(77, H))
    b.set_image_target(None, None)
    with pytest.raises(capi.RtoError):
        b.set_image_target(None, a.image_rgba8_ptr)
    # the other way to deliver a split frame: no assembly on a GPU, every context copies its own rows into ONE host frame
    # (rto_context_read_rows_rgba8 / rto_context_read_image_rows)
    c0, c1 = capi.RenderContext(W, H), capi.RenderContext(W, H)
    host8, host_img = capi.PinnedBuffer((H, W, 4), np.uint8), capi.PinnedBuffer((H, W, 4), np.float32)
    with pytest.raises(capi.RtoError, match="RGBA8"):
        c0.read_rows(host8=host8.array, rows=(0, 10))     # no RGBA8 copy yet
    for c, band in ((c0, (0, 77)), (c1, (77, H))):
        assert c.image_rgba8_ptr                           # from now on the filter epilogue writes the RGBA8 rows too
        c.rng_set_frame(6)
        y0, y1 = SH.render_rows_for_band(band, H, True, net.levels)
        capi.launch_renderer(t, cam, opt, c, rect=(0, y0, W, y1))
        net.denoise(cam, c, rows=band)
        c.read_rows(host8=host8.array, host_image=host_img.array, rows=band)
    assert np.array_equal(host8.array, want8) and np.array_equal(host_img.array, want)
    with pytest.raises(capi.RtoError, match="row range"):
        c0.read_rows(host8=host8.array, rows=(10, H + 1))


def test_frame_sequence_equals_frame_by_frame(capi, mid_tree, poses8, net_weights):
    """rto_frame_sequence (the pipelined driver's host loop in one library call: slot waits, rng + pose per frame, graph
    launches) delivers every frame's RGBA8 buffer exactly as a frame-by-frame rto_frame_launch_indexed loop does, reports
    every frame once, in issue order per slot, before its slot is reused, and drains on request."""
    W, H, n_slots, n_frames = 200, 152, 3, 11
    t, cam, opt, net = _rig(capi, mid_tree, W, H, 6, True, net_weights)
    ctxs = [capi.RenderContext(W, H) for _ in range(n_slots)]
    bufs = [capi.PinnedBuffer((H, W, 4), np.uint8) for _ in range(n_slots)]
    streams = [capi.stream_create() for _ in range(n_slots)]
    frames = [capi.Frame(ctxs[k], t, net, opt, cam.fx, cam.fy, rgba8=bufs[k]) for k in range(n_slots)]
    got, order = {}, []

    def retired(idx, slot):
        assert slot == idx % n_slots
        order.append(idx)
        got[idx] = bufs[slot].array.copy()

    seq = capi.FrameSequence(frames, streams, poses8, warmup=100, retired=retired)
    first = 5
    seq.run(first, 4)                       # frames 5..8: 5 is retired when slot 5 % 3 is reused by frame 8
    assert order == [5]
    seq.run(first + 4, n_frames - 4, drain=True)
    assert sorted(order) == list(range(first, first + n_frames)) and len(order) == n_frames
    for k in range(n_slots):                # per slot, frames retire in issue order
        mine = [i for i in order if i % n_slots == k]
        assert mine == sorted(mine)
    # frame by frame on one context
    ref = capi.RenderContext(W, H)
    b = capi.PinnedBuffer((H, W, 4), np.uint8)
    fr = capi.Frame(ref, t, net, opt, cam.fx, cam.fy, rgba8=b)
    for i in range(first, first + n_frames):
        fr.launch_indexed(poses8[i % len(poses8)], i, 100, streams[0])
        capi.synchronize(streams[0])
        assert np.array_equal(got[i], b.array), i
    assert len({got[i].tobytes() for i in got}) > 1
    # argument checks
    with pytest.raises(capi.RtoError):
        capi.FrameSequence(frames, streams, poses8).run(-1, 2)
    fr.close()
    for f in frames:
        f.close()
    for st in streams:
        capi.stream_destroy(st)
