"""GPU probe for the conv3x3 backward-filter kernels: first generation (conv_tc.cu, AIDE_WGRAD_HALO=0) vs the
halo / three-taps-per-CTA kernel (conv_wgrad_halo.cu).  Correctness against torch (fp32, TF32 off) on small shapes, then
per-layer timing on every fuseunet layer shape at batch 8 / 256x256 (L2 flushed before each timed launch).

    python tools/wgrad_probe.py [--fmts 3 2] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aide_b200 as A  # noqa: E402
from aide_b200 import engine as E, ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fmts", nargs="+", type=int, default=[3, 2])
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--json", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
NAMES = {1: "tf32x2", 2: "bf16", 3: "f16x2"}
TOL = {1: 3e-5, 2: 3e-2, 3: 3e-5}


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


bad = 0
for fmt in args.fmts:
    for (N, H, W, cin, cout) in [(2, 8, 8, 128, 256), (1, 16, 16, 128, 64), (2, 20, 24, 256, 128), (1, 4, 40, 128, 128),
                                 (3, 32, 32, 128, 64), (1, 16, 16, 1024, 512), (2, 16, 24, 32, 32), (1, 20, 12, 64, 32),
                                 (2, 16, 40, 32, 64), (3, 24, 40, 96, 96)]:
        g = torch.Generator().manual_seed(cin + H)
        x = torch.randn(N, cin, H, W, generator=g).to(dev)
        dz = torch.randn(N, cout, H, W, generator=g).to(dev)
        ref = torch.nn.grad.conv2d_weight(x.double(), (cout, cin, 3, 3), dz.double(), padding=1)
        errs = {}
        for halo in (0, 1):
            os.environ["AIDE_WGRAD_HALO"] = str(halo)
            try:
                dw = ops.conv3x3_wgrad(ops.from_nchw(x, fmt), ops.from_nchw(dz, fmt))
                torch.cuda.synchronize()
                errs[halo] = rel(dw, ref)
            except Exception as ex:  # noqa: BLE001
                errs[halo] = float("nan")
                print(f"[ERR] {NAMES[fmt]} halo={halo} {(N, H, W, cin, cout)}: {type(ex).__name__}: {ex}", flush=True)
        ok = errs[1] == errs[1] and errs[1] < TOL[fmt]
        bad += 0 if ok else 1
        print(f"[{'ok ' if ok else 'BAD'}] {NAMES[fmt]:6s} {str((N, H, W, cin, cout)):26s} old {errs[0]:.2e} halo {errs[1]:.2e}", flush=True)
print("CHECK", "PASS" if bad == 0 else "FAIL", flush=True)
if bad:
    sys.exit(3)

plan_net = E.plan_fuseunet(2)
shapes = {}
for u in plan_net.units:
    if not u.first:
        key = (u.cin, u.cout, 256 >> u.level)
        shapes[key] = shapes.get(key, 0) + 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B = args.batch
result = {}
for fmt in args.fmts:
    tot = {0: 0.0, 1: 0.0}
    tot_flop = 0.0
    rows = []
    for (cin, cout, hw), count in sorted(shapes.items(), key=lambda kv: -kv[0][2]):
        x = ops.Act(B, hw, hw, cin, fmt, dev)
        x.planes.normal_()
        dz = ops.Act(B, hw, hw, cout, fmt, dev)
        dz.planes.normal_()
        inv = torch.full((1,), 1.0 / 256.0, device=dev)
        dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        ms = {}
        for halo in (0, 1):
            os.environ["AIDE_WGRAD_HALO"] = str(halo)
            nbytes = A.lib.aide_conv3x3_wgrad_workspace_bytes(fmt, cin, cout, B, hw, hw)
            ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)

            def launch():
                ops.call("aide_conv3x3_wgrad", fmt, x.p0, x.p1, x.C, 0, cin, dz.p0, dz.p1,
                         inv.data_ptr() if fmt == 3 else None, cout, B, hw, hw, ws.data_ptr(), nbytes, dw.data_ptr(), st)
            launch()
            t = 0.0
            for _ in range(3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); launch(); e1.record()
                torch.cuda.synchronize()
                t += e0.elapsed_time(e1)
            ms[halo] = t / 3
        os.environ["AIDE_WGRAD_HALO"] = "1"
        flop = 2.0 * B * hw * hw * cout * cin * 9
        tot_flop += flop * count
        for h in (0, 1):
            tot[h] += ms[h] * count
        rows.append(dict(cin=cin, cout=cout, hw=hw, n=count, ms_old=round(ms[0], 4), ms_halo=round(ms[1], 4),
                         tf_old=round(flop / ms[0] / 1e9, 1), tf_halo=round(flop / ms[1] / 1e9, 1)))
        print(f"{NAMES[fmt]:6s} {cin:4d}->{cout:3d} @{hw:3d} x{count}: old {ms[0]:.4f} ms ({flop / ms[0] / 1e9:6.1f} TF)  "
              f"halo {ms[1]:.4f} ms ({flop / ms[1] / 1e9:6.1f} TF)", flush=True)
        del x, dz
    print(f"== {NAMES[fmt]} wgrad total (incl. split-K reduce): old {tot[0]:.3f} ms ({tot_flop / tot[0] / 1e9:.1f} TF)  "
          f"halo {tot[1]:.3f} ms ({tot_flop / tot[1] / 1e9:.1f} TF)", flush=True)
    result[NAMES[fmt]] = dict(layers=rows, total_ms_old=tot[0], total_ms_halo=tot[1])
if args.json:
    with open(args.json, "w") as f:
        json.dump(result, f, indent=1)
