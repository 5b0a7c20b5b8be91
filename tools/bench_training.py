"""Time the training operator (forward-with-save + backward) against the reference extension when oracle/_ref has it.
Usage (GPU box): python tools/bench_training.py [B L H W]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rt_octree_b200 import training as T  # noqa: E402


def timeit(op, weight, guidance, img, dout, iters=20):
    def step():
        w = weight.clone().requires_grad_(True)
        g = guidance.clone().requires_grad_(True)
        out = op(w, g, img, True)
        out.backward(dout)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B, L, H, W = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (4, 4, 800, 800)
    g = torch.Generator().manual_seed(0)
    weight = torch.softmax(torch.randn((B, L, H, W), generator=g), 1).cuda()
    guidance = (torch.rand((B, L, H, W), generator=g) * 6).cuda()
    img = torch.rand((B, H, W, 4), generator=g).cuda()
    dout = torch.randn((B, H, W, 4), generator=g).cuda()
    res = {"shape": [B, L, H, W], "ours_ms_fwd_bwd": timeit(T.filtering_autograd, weight, guidance, img, dout)}
    d = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(d, "_denoiser_ref.so")):
        sys.path.insert(0, d)
        import _denoiser_ref

        res["reference_ms_fwd_bwd"] = timeit(_denoiser_ref.filtering_autograd, weight, guidance, img, dout)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
