"""GPU parity tests of the training-side operator `filtering_autograd` (forward-with-save + backward).

Checkers: (1) a float64 torch restatement of the filter differentiated by torch autograd, (2) when oracle/_ref was
built, the UNMODIFIED reference extension (`_denoiser_ref` = denoiser/extension/bindings.cpp + filtering.cu).
Tolerance: the op is fp32 with `__expf`; outputs and gradients are compared at 2e-5 absolute on O(1) values
(the reference itself sums grad_guidance with atomics in arbitrary order).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 2e-5


def _torch_filter64(weight, guidance, img):
    """float64 restatement: out(p) = sum_l w_l(p) * sum_q k rgb(q) / sum_q k, k = exp(g_l(q)), q in the (2l+3)^2 window
    clipped to the image (filtering.cu:120-228)."""
    import torch
    import torch.nn.functional as F

    B, L, H, W = guidance.shape
    out = torch.zeros((B, H, W, 3), dtype=torch.float64, device=img.device)
    rgb = img[..., :3].permute(0, 3, 1, 2)                       # [B,3,H,W]
    for l in range(L):
        S = l + 1
        g = guidance[:, l]
        gp = F.pad(g, (S, S, S, S), value=float("-inf"))
        rp = F.pad(rgb, (S, S, S, S), value=0.0)
        taps_g, taps_r = [], []
        for dy in range(2 * S + 1):
            for dx in range(2 * S + 1):
                taps_g.append(gp[:, dy:dy + H, dx:dx + W])
                taps_r.append(rp[:, :, dy:dy + H, dx:dx + W])
        tg = torch.stack(taps_g, 1)                              # [B,T,H,W]
        tr = torch.stack(taps_r, 1)                              # [B,T,3,H,W]
        k = torch.softmax(tg, dim=1)
        f = (k[:, :, None] * tr).sum(1)                          # [B,3,H,W]
        out = out + (weight[:, l, None] * f).permute(0, 2, 3, 1)
    return torch.cat([out, torch.ones_like(out[..., :1])], -1)


def _inputs(B, L, H, W, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    weight = torch.softmax(torch.randn((B, L, H, W), generator=g), 1).cuda()
    guidance = (torch.rand((B, L, H, W), generator=g) * 6.0).cuda()      # relu6 range of the net's output
    img = torch.rand((B, H, W, 4), generator=g).cuda()
    dout = torch.randn((B, H, W, 4), generator=g).cuda()
    return weight, guidance, img, dout


@pytest.mark.parametrize("B,L,H,W", [(1, 4, 33, 47), (2, 6, 20, 17), (1, 1, 5, 3), (3, 2, 16, 64)])
def test_filtering_autograd_matches_torch_autograd(B, L, H, W):
    import torch

    from rt_octree_b200 import training as T

    weight, guidance, img, dout = _inputs(B, L, H, W, 1000 * B + L)
    w = weight.clone().requires_grad_(True)
    g = guidance.clone().requires_grad_(True)
    out = T.filtering_autograd(w, g, img, True)
    out.backward(dout)

    w64 = weight.double().requires_grad_(True)
    g64 = guidance.double().requires_grad_(True)
    ref = _torch_filter64(w64, g64, img.double())
    ref.backward(dout.double())

    assert (out.double() - ref).abs().max().item() < TOL
    assert (w.grad.double() - w64.grad).abs().max().item() < TOL * 5
    assert (g.grad.double() - g64.grad).abs().max().item() < TOL * 5
    # forward with and without saving is the same kernel
    out2 = T.filtering_autograd(weight, guidance, img, False)
    assert torch.equal(out2, out.detach())


def test_filtering_autograd_matches_inference_filter(capi):
    """requires_grad=False goes through rto_filter, the op the renderer calls; the saving variant is bit-identical."""
    import torch

    from rt_octree_b200 import training as T

    weight, guidance, img, _ = _inputs(1, 4, 64, 80, 7)
    a = T.filtering_autograd(weight, guidance, img, False)
    out = torch.zeros_like(img[0])
    capi.filtering(weight[0].data_ptr(), guidance[0].data_ptr(), img[0].data_ptr(), 4, 80, 64, out.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(a[0], out)


def test_filtering_autograd_errors():
    import torch

    from rt_octree_b200 import training as T

    w = torch.zeros((1, 7, 8, 8), device="cuda")
    with pytest.raises(RuntimeError, match="Kernel size == 15 not supported"):
        T.filtering_autograd(w, w, torch.zeros((1, 8, 8, 4), device="cuda"), False)
    with pytest.raises(RuntimeError, match="CUDA"):
        T.filtering_autograd(w.cpu(), w.cpu(), torch.zeros((1, 8, 8, 4)), False)
    w4 = torch.zeros((1, 4, 8, 8), device="cuda", requires_grad=True)
    out = T.filtering_autograd(w4, w4.detach(), torch.zeros((1, 8, 8, 4), device="cuda"), False)
    with pytest.raises(RuntimeError, match="requires_grad=False"):
        out.sum().backward()


def test_backward_is_deterministic():
    """Gather formulation: two backward passes give bit-identical gradients (the reference's atomicAdd scatter does not
    guarantee that)."""
    import torch

    from rt_octree_b200 import training as T

    weight, guidance, img, dout = _inputs(1, 6, 96, 128, 3)
    grads = []
    for _ in range(2):
        g = guidance.clone().requires_grad_(True)
        T.filtering_autograd(weight, g, img, True).backward(dout)
        grads.append(g.grad.clone())
    assert torch.equal(grads[0], grads[1])


def _ref_ext():
    d = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(d, "_denoiser_ref.so")):
        pytest.skip("oracle/_ref/_denoiser_ref.so not built")
    import torch  # noqa: F401  (the extension links libtorch)

    if d not in sys.path:
        sys.path.insert(0, d)
    import _denoiser_ref

    return _denoiser_ref


@pytest.mark.parametrize("B,L,H,W", [(2, 4, 100, 100), (1, 6, 64, 48)])
def test_filtering_autograd_matches_reference_extension(B, L, H, W):
    """Against the reference's own compiled operator (forward and both gradients)."""
    from rt_octree_b200 import training as T

    ref = _ref_ext()
    weight, guidance, img, dout = _inputs(B, L, H, W, 77 + L)
    res = []
    for op in (T.filtering_autograd, ref.filtering_autograd):
        w = weight.clone().requires_grad_(True)
        g = guidance.clone().requires_grad_(True)
        out = op(w, g, img, True)
        out.backward(dout)
        res.append((out.detach(), w.grad, g.grad))
    (o, gw, gg), (ro, rgw, rgg) = res
    assert (o[..., :3] - ro[..., :3]).abs().max().item() < 2e-6
    assert (gw - rgw).abs().max().item() < TOL
    assert (gg - rgg).abs().max().item() < TOL


def test_reference_training_step_runs_on_this_op():
    """The reference's training recipe (runner.py:70-86: GuidanceNet -> softmax/guidance -> filtering_autograd -> L1 loss
    -> GradScaler backward -> Adam), restated with plain torch modules around OUR operator: the loss goes down."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    from rt_octree_b200 import training as T

    torch.manual_seed(0)
    L, H, W = 4, 64, 64
    net = nn.Sequential(nn.Conv2d(8, 8, 3, padding=1), nn.ReLU6(), nn.Conv2d(8, 2 * L, 3, padding=1), nn.ReLU6()).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
    gt = torch.stack([xx, yy, 0.5 * (xx + yy), torch.ones_like(xx)], -1)[None].cuda()
    noisy = (gt + 0.2 * torch.randn_like(gt)).clamp(0, 1)
    aux = torch.cat([noisy[..., :3], gt[..., 3:], noisy[..., :3] ** 2, gt[..., 3:]], -1).permute(0, 3, 1, 2).contiguous()
    losses = []
    for _ in range(30):
        opt.zero_grad(set_to_none=True)
        x = net(aux).float()
        wm = F.softmax(x[:, :L].contiguous(), dim=1)
        gm = x[:, L:].contiguous()
        out = T.filtering_autograd(wm, gm, noisy, True)
        loss = (out[..., :3] - gt[..., :3]).abs().mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] * 0.9, losses
