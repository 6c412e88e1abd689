"""GPU probe: how does the tensor-core accumulation error of the conv kernels grow with the reduction length K?

Inputs are pre-rounded to bf16 so that in every mode the *products* are exact and the only error left is the
accumulation inside tcgen05.mma (vs fp32 CUDA-core FMA chains in 'exact' mode).  Reference = fp64 on the CPU.
Prints max|err|/max|ref| and rms(err)/rms(ref) for K = 9*cin, random-sign and all-positive data."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aide_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
NAMES = {0: "exact", 1: "parity", 2: "fast"}


def stats(got, ref):
    e = got.cpu().double() - ref
    return (e.abs().max() / ref.abs().max()).item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(), \
        (e.mean() / ref.abs().mean()).item()


for positive in (False, True):
    for cin in (32, 128, 512, 1024):
        g = torch.Generator().manual_seed(cin)
        x = torch.randn(1, cin, 16, 16, generator=g)
        w = torch.randn(64, cin, 3, 3, generator=g) * (9 * cin) ** -0.5
        if positive:
            x, w = x.abs(), w.abs()
        x, w = x.bfloat16().float(), w.bfloat16().float()
        ref = F.conv2d(x.double(), w.double(), None, padding=1)
        line = f"{'pos ' if positive else 'rand'} K={9 * cin:5d}"
        for fmt in (0, 1, 2):
            a = ops.from_nchw(x.to(dev), fmt)
            z, _ = ops.conv3x3(a, w.to(dev), None)
            torch.cuda.synchronize()
            mx, rms, bias = stats(ops.nhwc_to_nchw(z), ref)
            line += f" | {NAMES[fmt]} max {mx:.2e} rms {rms:.2e} bias {bias:+.1e}"
        print(line, flush=True)

# wgrad: reduction over pixels
for npix_side in (16, 64, 128):
    g = torch.Generator().manual_seed(npix_side)
    x = torch.randn(1, 32, npix_side, npix_side, generator=g).bfloat16().float()
    dz = torch.randn(1, 32, npix_side, npix_side, generator=g).bfloat16().float()
    ref = torch.nn.grad.conv2d_weight(x.double(), (32, 32, 3, 3), dz.double(), padding=1)
    line = f"wgrad K={npix_side * npix_side:6d}"
    for fmt in (0, 1, 2):
        dw = ops.conv3x3_wgrad(ops.from_nchw(x.to(dev), fmt), ops.from_nchw(dz.to(dev), fmt))
        torch.cuda.synchronize()
        mx, rms, bias = stats(dw, ref)
        line += f" | {NAMES[fmt]} max {mx:.2e} rms {rms:.2e}"
    print(line, flush=True)
