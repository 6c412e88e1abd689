"""Small-shape launches of every tcgen05 / TMA kernel, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
    compute-sanitizer --tool synccheck python tools/sanitize_probe.py

Shapes are chosen so that each kernel family is hit: the halo forward/dgrad kernel (maps >= 8x8), the first-generation
forward kernel (2x2 / 4x4 maps), the halo wgrad kernel (Cin >= 128) and the first-generation wgrad kernel (Cin 32/64),
in the split-precision (f16x2, tf32x2) and bf16 operand formats; results are checked against torch on the way."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aide_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
SHAPES = [(1, 16, 16, 32, 32), (1, 16, 24, 64, 64), (2, 8, 8, 128, 256), (1, 4, 4, 256, 128), (1, 32, 16, 128, 64)]
TOL = {1: 3e-5, 2: 3e-2, 3: 3e-5}


def relmax(a, b):
    return ((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max()).item()


worst = 0.0
for fmt in (3, 2, 1):
    for (N, H, W, cin, cout) in SHAPES:
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, cin, H, W, generator=g)
        w = torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5
        b = torch.randn(cout, generator=g)
        dz = torch.randn(N, cout, H, W, generator=g)
        a = ops.from_nchw(x.to(dev), fmt)
        z, part = ops.conv3x3(a, w.to(dev), b.to(dev), stats=True)
        dza = ops.from_nchw(dz.to(dev), fmt)
        dx = ops.conv3x3_dgrad(dza, w.to(dev))
        dw = ops.conv3x3_wgrad(a, dza)
        torch.cuda.synchronize()
        e = [relmax(ops.nhwc_to_nchw(z), F.conv2d(x, w, b, padding=1)),
             relmax(ops.nhwc_to_nchw(dx), torch.nn.grad.conv2d_input(x.shape, w, dz, padding=1)),
             relmax(dw, torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1))]
        assert max(e) < TOL[fmt], (fmt, (N, H, W, cin, cout), e)
        worst = max(worst, max(e) / TOL[fmt])
        print(f"fmt {fmt} {N}x{H}x{W} {cin}->{cout}: fwd {e[0]:.1e} dgrad {e[1]:.1e} wgrad {e[2]:.1e}", flush=True)
print(f"SANITIZE_PROBE_OK worst error / tolerance {worst:.2f}")
