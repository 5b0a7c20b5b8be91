"""tools/make_ts_module.py: the TorchScript stand-in handed to the reference binary reproduces the reference module
(tests/golden/guidance_net_ref.npz), and the weight export round-trips."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_ts_module_equals_reference_module(tmp_path, net_weights):
    import torch

    import make_ts_module as M

    g = np.load(os.path.join(ROOT, "tests", "golden", "guidance_net_ref.npz"))
    p = str(tmp_path / "ts_test.ts")
    M.make_ts(net_weights, p, device="cpu")
    m = torch.jit.load(p)
    with torch.no_grad():
        w, gd = m(torch.from_numpy(g["aux"])[None])
    assert np.array_equal(gd[0].numpy(), g["guidance_fp16"])
    assert np.array_equal(w[0].numpy(), g["weight_fp16"])
    back = M.export_weights(p)
    for k in ("w1", "b1", "w2", "b2"):
        assert np.array_equal(back[k], net_weights[k])


def test_export_weights_from_reference_ts(net_weights):
    """ts_ref_cpu.ts was written by the REFERENCE's compact_and_compile (denoiser/network.py:170-208) in this container
    (tools/make_golden.py): a traced closure whose weights are prim::Constant tensors.  The exporter must recover the
    same four fp16 tensors as the reference module's own parameters."""
    import make_ts_module as M

    back = M.export_weights(os.path.join(ROOT, "tests", "golden", "ts_ref_cpu.ts"))
    for k in ("w1", "b1", "w2", "b2"):
        assert back[k].dtype == np.float16 and np.array_equal(back[k], net_weights[k]), k
