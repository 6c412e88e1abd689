"""First-contact diagnostics on the GPU box: run each kernel family once, print errors instead of stopping."""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aide_b200 import ops, lib  # noqa: E402

dev = torch.device("cuda:0")
print("device", torch.cuda.get_device_name(0), "has_tma", lib.aide_has_tma(), flush=True)


def rel(a, b):
    return ((a.cpu().double() - b.cpu().double()).abs().max() / b.cpu().double().abs().max().clamp_min(1e-30)).item()


def run(name, fn):
    try:
        out = fn()
        torch.cuda.synchronize()
        print(f"[ok ] {name}: {out}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"[ERR] {name}: {type(e).__name__}: {e}", flush=True)
        traceback.print_exc()


def conv_case(fmt, N, H, W, cin, cout):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5
    b = torch.randn(cout, generator=g)
    dz = torch.randn(N, cout, H, W, generator=g)
    ref = F.conv2d(x, w, b, padding=1)
    a = ops.from_nchw(x.to(dev), fmt)
    z, part = ops.conv3x3(a, w.to(dev), b.to(dev), stats=True)
    torch.cuda.synchronize()
    e_f = rel(ops.nhwc_to_nchw(z), ref)
    e_s = rel(part.sum(0)[0], ref.sum((0, 2, 3)))
    dza = ops.from_nchw(dz.to(dev), fmt)
    e_d = rel(ops.nhwc_to_nchw(ops.conv3x3_dgrad(dza, w.to(dev))), torch.nn.grad.conv2d_input(x.shape, w, dz, padding=1))
    torch.cuda.synchronize()
    e_w = rel(ops.conv3x3_wgrad(a, dza), torch.nn.grad.conv2d_weight(x, w.shape, dz, padding=1))
    return f"fwd {e_f:.2e} stats {e_s:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e}"


for fmt, nm in ((0, "exact"), (2, "fast"), (1, "parity")):
    for shp in [(1, 16, 16, 64, 64), (2, 16, 16, 32, 32), (2, 8, 8, 128, 256), (1, 20, 12, 64, 32)]:
        run(f"conv {nm} {shp}", lambda: conv_case(fmt, *shp))
