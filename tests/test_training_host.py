"""Host-side checks of the training operator mirror (no GPU): API surface, argument validation, and that the
reference's own `denoiser/network.py` binds to it through `import _denoiser` (only where /root/reference exists)."""
import inspect
import os
import sys

import pytest


def test_signature_matches_reference_binding():
    """bindings.cpp:5-13: filtering_autograd(weight_map, guidance_map, imgs_in, requires_grad=False)."""
    from rt_octree_b200 import training as T

    sig = inspect.signature(T.filtering_autograd)
    assert list(sig.parameters) == ["weight_map", "guidance_map", "imgs_in", "requires_grad"]
    assert sig.parameters["requires_grad"].default is False


def test_cpu_tensors_fail_loudly():
    import torch

    from rt_octree_b200 import training as T

    w = torch.zeros((1, 4, 8, 8))
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        T.filtering_autograd(w, w, torch.zeros((1, 8, 8, 4)))


def test_checker_rejects_before_any_launch():
    import torch

    from rt_octree_b200 import training as T

    w = torch.zeros((1, 4, 8, 8))
    with pytest.raises(RuntimeError):
        T._check_inputs(w, w, torch.zeros((1, 8, 8, 4)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/denoiser"), reason="reference tree not present")
def test_reference_network_module_binds_to_shim():
    """denoiser/network.py:7-46 tries `import _denoiser` before JIT-building its extension; after
    install_as_denoiser_extension() it must pick ours up and its `filtering()` must reach our operator."""
    import torch

    from rt_octree_b200 import training as T

    mod = T.install_as_denoiser_extension()
    sys.path.insert(0, "/root/reference")
    try:
        sys.modules.pop("denoiser.network", None)
        import denoiser.network as N

        assert N._denoiser is mod
        net = N.GuidanceNet(8, 8, 2, 2, 4)
        assert callable(net.filtering)
        # on CPU the call must reach our operator and be refused there (no CPU fallback)
        aux = torch.zeros((1, 8, 16, 16))
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            net.filtering(aux, torch.zeros((1, 16, 16, 4)))
    finally:
        sys.path.remove("/root/reference")
        sys.modules.pop("_denoiser", None)
        for k in [k for k in sys.modules if k == "denoiser" or k.startswith("denoiser.")]:
            sys.modules.pop(k)
