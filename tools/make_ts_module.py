#!/usr/bin/env python
"""Builds a TorchScript file the REFERENCE binary can load (`--ts_module`, renderer/src/denoiser/denoiser.cpp:20) from the
raw GuidanceNet tensors (w1,b1,w2,b2 fp16), on the device it will run on.

Why this exists: the reference's compact_and_compile (denoiser/network.py:170-208) traces on the training GPU, so its
ts_*.ts carries CUDA constants; /root/reference is not present on the GPU box and a CPU-traced file would carry CPU
constants (the reference loads without a device map).  The module below is the same graph — x.half() -> [conv2d 'same'
+ bias -> relu6] x2 -> float -> softmax(ch :L) , ch L: — and tests/test_ts_module.py checks on CPU that it reproduces
the reference module's outputs bit for bit (tests/golden/guidance_net_ref.npz).

Also provides the inverse used once per trained model: export_weights(ts_path) -> dict of the four fp16 tensors
(they are prim::Constant nodes of the traced graph, or parameters of a scripted module).
"""
import argparse
import sys

import numpy as np
import torch
import torch.nn.functional as F


class DeployedGuidanceNet(torch.nn.Module):
    def __init__(self, w):
        super().__init__()
        self.w1 = torch.nn.Parameter(torch.from_numpy(np.asarray(w["w1"], np.float16).copy()), requires_grad=False)
        self.b1 = torch.nn.Parameter(torch.from_numpy(np.asarray(w["b1"], np.float16).copy()), requires_grad=False)
        self.w2 = torch.nn.Parameter(torch.from_numpy(np.asarray(w["w2"], np.float16).copy()), requires_grad=False)
        self.b2 = torch.nn.Parameter(torch.from_numpy(np.asarray(w["b2"], np.float16).copy()), requires_grad=False)
        self.levels = int(self.w2.shape[0] // 2)

    def forward(self, aux_buffer):
        x = aux_buffer.half()
        x = F.relu6(F.conv2d(x, self.w1, self.b1, padding="same"))
        x = F.relu6(F.conv2d(x, self.w2, self.b2, padding="same"))
        x = x.float()
        weight_map = F.softmax(x[:, :self.levels, ...].contiguous(), dim=1)
        guidance_map = x[:, self.levels:, ...].contiguous()
        return weight_map, guidance_map


def make_ts(weights: dict, out_path: str, device: str = "cuda"):
    m = DeployedGuidanceNet(weights).eval().to(device)
    ex = torch.rand((1, 8, 32, 32), device=device)
    with torch.no_grad():
        ts = torch.jit.trace(m, (ex,))
    torch.jit.save(ts, out_path)
    return out_path


def export_weights(ts_path: str) -> dict:
    """The four fp16 tensors of a ts_*.ts, identified by shape (Appendix B)."""
    m = torch.jit.load(ts_path, map_location="cpu")
    tensors = [p.detach() for p in m.parameters()]
    if not tensors:
        g = m.graph if hasattr(m, "graph") else m.forward.graph
        for n in g.findAllNodes("prim::Constant"):
            if n.hasAttribute("value") and n.kindOf("value") == "t":
                tensors.append(n.t("value"))
    t4 = [t for t in tensors if t.dim() == 4]
    t1 = [t for t in tensors if t.dim() == 1]
    if len(t4) != 2 or len(t1) != 2:
        raise ValueError("expected two conv weights and two biases, found %d/%d" % (len(t4), len(t1)))
    # w1 maps the 8 aux channels; w2 consumes w1's output channels
    w1, w2 = (t4[0], t4[1]) if t4[0].shape[1] == 8 and t4[1].shape[1] == t4[0].shape[0] else (t4[1], t4[0])
    b1, b2 = (t1[0], t1[1]) if t1[0].shape[0] == w1.shape[0] and t1[1].shape[0] == w2.shape[0] else (t1[1], t1[0])
    if w1.shape[0] == w2.shape[0]:
        raise ValueError("ambiguous shapes (mid == 2L): walk the conv nodes in topological order instead")
    return {k: v.cpu().half().numpy() for k, v in (("w1", w1), ("b1", b1), ("w2", w2), ("b2", b2))}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", help="npz with w1,b1,w2,b2 (fp16)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--export", help="instead: read this ts file and write its weights to --out (npz)")
    a = ap.parse_args()
    if a.export:
        np.savez(a.out, **export_weights(a.export))
    else:
        z = np.load(a.weights)
        make_ts({k: z[k] for k in ("w1", "b1", "w2", "b2")}, a.out, a.device)
    print(a.out)
