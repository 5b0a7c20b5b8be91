#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF, run in this container:

  guidance_net_ref.npz   weights of the reference's GuidanceNet(8,32,5,2,4) (torch.manual_seed(0)) re-parameterised by
                         the reference's own GuidanceNetCompact (denoiser/network.py:123-168) and cast .half() as
                         compact_and_compile does (:170-180); a small aux input; and the reference module's outputs
                         (weight_map, guidance_map) evaluated (a) by the fp16 compact module on CPU, i.e. the deployed
                         graph `cast_and_forward` (:195-197), and (b) by the un-compacted fp32 5-branch model.
  trace_ref_cpu.npz      aux buffers rendered by the reference's own trace_ray (rt_core.cuh host-compiled through
                         oracle/ref_cpu_shim.cpp) on a small synthetic tree, for several spp / poses.

Needs /root/reference (import of denoiser.network with the CUDA extension module `_denoiser` stubbed out, because
network.py:7-47 JIT-builds it at import) and oracle/_ref/libref_cpu.so (oracle/build_ref.sh cpu).
Run from the repo root:  python tools/make_golden.py
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("REF_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def golden_guidance_net():
    import torch

    sys.modules["_denoiser"] = types.ModuleType("_denoiser")  # stub: avoids the JIT CUDA build at import
    sys.path.insert(0, REF)
    # network.py calls torch.utils.cpp_extension.load at import unless _denoiser imports; make the guard pass
    import importlib

    net_mod = importlib.import_module("denoiser.network")
    torch.manual_seed(0)
    model = net_mod.GuidanceNet(8, 32, 5, 2, 4).eval()   # configs/blender.txt:21-25
    compact = net_mod.GuidanceNetCompact(model).eval()
    g = torch.Generator().manual_seed(1)
    aux = torch.rand((1, 8, 24, 40), generator=g)
    aux[:, 4:] = aux[:, :4] ** 2                          # like the renderer: ch4..7 are squares of ch0..3
    with torch.no_grad():
        w32, g32 = model(aux)                             # fp32, 5-branch (CPU autocast is a no-op here)
        c32w, c32g = compact(aux)                         # fp32 compact
        half = compact.half()
        w1, b1 = half.layers[0].conv.weight, half.layers[0].conv.bias
        w2, b2 = half.layers[1].conv.weight, half.layers[1].conv.bias
        try:
            w16, g16 = half(aux.half())                   # the deployed graph: cast_and_forward
            have16 = True
        except Exception as e:  # CPU half conv unsupported in this torch build
            print("fp16 CPU forward unavailable:", e)
            w16, g16, have16 = c32w, c32g, False
    # the reference's own export path: compact_and_compile -> traced TorchScript (constants on CPU here)
    torch.manual_seed(0)
    model2 = net_mod.GuidanceNet(8, 32, 5, 2, 4).eval()
    ts = net_mod.compact_and_compile(model2, "cpu")
    torch.jit.save(ts, os.path.join(OUT, "ts_ref_cpu.ts"))
    np.savez(os.path.join(OUT, "guidance_net_ref.npz"),
             w1=w1.numpy(), b1=b1.numpy(), w2=w2.numpy(), b2=b2.numpy(), aux=aux[0].numpy(),
             weight_fp16=w16[0].float().numpy(), guidance_fp16=g16[0].float().numpy(), have_fp16=np.bool_(have16),
             weight_fp32_full=w32[0].numpy(), guidance_fp32_full=g32[0].numpy(),
             weight_fp32_compact=c32w[0].numpy(), guidance_fp32_compact=c32g[0].numpy())
    print("guidance_net_ref.npz: fp16 path", have16, "| full-vs-compact fp32 max diff",
          float((g32 - c32g).abs().max()), "| fp16-vs-fp32 max diff", float((g16.float() - c32g).abs().max()))


def golden_trace():
    from oracle import oracle as O
    from rt_octree_b200 import synthetic as S

    tree = S.make_tree(depth=6, shell=1.0, halo=0.05, seed=3)
    poses = S.poses_to_c2w12(S.make_poses(8))
    W, H = 48, 40
    fx = S.blender_focal(W)
    out = {"W": W, "H": H, "fx": np.float32(fx), "tree_depth": 6, "tree_shell": 1.0, "tree_halo": 0.05, "tree_seed": 3}
    for spp in (1, 6, 32):
        for pi in (0, 3):
            rng = O.frame_rng(pi)
            out["aux_spp%d_pose%d" % (spp, pi)] = O.ref_cpu_render(tree, poses[pi], W, H, fx, fx, spp, rng)
    np.savez_compressed(os.path.join(OUT, "trace_ref_cpu.npz"), **out)
    print("trace_ref_cpu.npz written")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    golden_trace()
    golden_guidance_net()
