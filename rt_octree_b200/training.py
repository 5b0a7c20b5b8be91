"""Training-side mirror of the reference's Python operator `_denoiser.filtering_autograd`.

Reference: denoiser/extension/bindings.cpp:5-13 (the pybind entry), denoiser/extension/filtering.cu:580-699 (the
autograd Function), denoiser/network.py:8-47,76-83 (the only caller).  Same name, arguments and semantics:

    img_out = filtering_autograd(weight_map[B,L,H,W], guidance_map[B,L,H,W], imgs_in[B,H,W,4], requires_grad=False)

with gradients for weight_map and guidance_map only.  torch is plumbing here (device memory, the autograd tape, the
current stream); the arithmetic is `rto_filter_forward_save` / `rto_filter_backward` / `rto_filter` of the C-ABI library.
There is no CPU path: non-CUDA tensors raise.

`install_as_denoiser_extension()` registers this module as `_denoiser` in `sys.modules`, so the reference's
`denoiser/network.py` (`try: import _denoiser`) binds to it instead of JIT-building its own extension.
"""
from __future__ import annotations

import sys
import types

import torch

from . import capi


def _check_inputs(weight_map, guidance_map, imgs_in):
    for name, t in (("weight_map", weight_map), ("guidance_map", guidance_map), ("imgs_in", imgs_in)):
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")           # filtering.h CHECK_CUDA
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name} must be float32")
    if guidance_map.dim() != 4 or weight_map.shape != guidance_map.shape:
        raise RuntimeError("weight_map and guidance_map must both be [B, L, H, W]")
    B, L, H, W = guidance_map.shape
    if tuple(imgs_in.shape) != (B, H, W, 4):
        raise RuntimeError("imgs_in must be [B, H, W, 4]")
    if not (weight_map.device == guidance_map.device == imgs_in.device):
        raise RuntimeError("inputs must live on one device")
    return B, L, H, W


class Filtering(torch.autograd.Function):
    """Filtering (filtering.cu:580-699): forward loops over the batch like the reference; tensors saved for backward
    are the inputs plus rgb_filtered [B,L,H,W,4], max_map and inv_kernel_sum [B,L,H,W]."""

    @staticmethod
    def forward(ctx, weight_map, guidance_map, imgs_in, requires_grad=False):
        B, L, H, W = _check_inputs(weight_map, guidance_map, imgs_in)
        weight_map, guidance_map, imgs_in = weight_map.contiguous(), guidance_map.contiguous(), imgs_in.contiguous()
        lib = capi.load()
        img_out = torch.empty_like(imgs_in)
        ctx.requires_grad_ = bool(requires_grad)
        with torch.cuda.device(imgs_in.device):
            stream = torch.cuda.current_stream().cuda_stream
            if requires_grad:
                rgb_f = torch.empty((B, L, H, W, 4), device=imgs_in.device, dtype=torch.float32)
                max_map = torch.empty((B, L, H, W), device=imgs_in.device, dtype=torch.float32)
                inv_sum = torch.empty((B, L, H, W), device=imgs_in.device, dtype=torch.float32)
                for i in range(B):
                    capi._check(lib.rto_filter_forward_save(
                        weight_map[i].data_ptr(), guidance_map[i].data_ptr(), imgs_in[i].data_ptr(), L, W, H,
                        img_out[i].data_ptr(), rgb_f[i].data_ptr(), max_map[i].data_ptr(), inv_sum[i].data_ptr(), stream))
                ctx.save_for_backward(weight_map, guidance_map, imgs_in, rgb_f, max_map, inv_sum)
            else:
                for i in range(B):
                    capi._check(lib.rto_filter(weight_map[i].data_ptr(), guidance_map[i].data_ptr(),
                                               imgs_in[i].data_ptr(), L, W, H, img_out[i].data_ptr(), stream))
        return img_out

    @staticmethod
    def backward(ctx, grad_output):
        if not ctx.requires_grad_:
            raise RuntimeError("filtering_autograd was called with requires_grad=False; nothing was saved for backward")
        weight_map, guidance_map, imgs_in, rgb_f, max_map, inv_sum = ctx.saved_tensors
        B, L, H, W = guidance_map.shape
        grad_output = grad_output.contiguous()
        grad_weight = torch.empty_like(weight_map)
        grad_guidance = torch.empty_like(guidance_map)
        lib = capi.load()
        with torch.cuda.device(imgs_in.device):
            stream = torch.cuda.current_stream().cuda_stream
            for i in range(B):
                capi._check(lib.rto_filter_backward(
                    grad_output[i].data_ptr(), imgs_in[i].data_ptr(), weight_map[i].data_ptr(), guidance_map[i].data_ptr(),
                    rgb_f[i].data_ptr(), max_map[i].data_ptr(), inv_sum[i].data_ptr(), L, W, H,
                    grad_weight[i].data_ptr(), grad_guidance[i].data_ptr(), stream))
        return grad_weight, grad_guidance, None, None


def filtering_autograd(weight_map, guidance_map, imgs_in, requires_grad=False):
    """denoiser::filtering_autograd (filtering.cu:701-710 / bindings.cpp:5-13)."""
    return Filtering.apply(weight_map, guidance_map, imgs_in, requires_grad)


def install_as_denoiser_extension():
    """Make `import _denoiser` (denoiser/network.py:8) resolve to this implementation."""
    mod = types.ModuleType("_denoiser")
    mod.filtering_autograd = filtering_autograd
    mod.__doc__ = "rt_octree_b200 drop-in for the reference's _denoiser extension"
    sys.modules["_denoiser"] = mod
    return mod
