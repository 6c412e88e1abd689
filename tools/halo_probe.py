"""GPU probe for the halo-reuse conv kernel (csrc/conv_halo_tc.cu).

1. correctness of the shifted-descriptor addressing (tap views of one halo tile: UMMA start address moved by whole
   pixel rows, SBO = 10 rows) per operand format, against torch conv2d on the same device (fp32, TF32 disabled).
   Measured in round 1: the plain start address works; setting the descriptor's base-offset field does NOT;
2. per-layer timing, first-generation kernel (AIDE_CONV_HALO=0) vs halo kernel, every fuseunet layer shape at
   batch 8 / 256x256, L2 flushed before each timed launch.

    python tools/halo_probe.py [--no-timing] [--fmts 3 1 2]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aide_b200 as A  # noqa: E402
from aide_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--no-timing", action="store_true")
ap.add_argument("--fmts", nargs="+", type=int, default=[3, 1, 2])
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--json", default="")
ap.add_argument("--sweep", action="store_true", help="time every (cout tile, pixel-tile blocking) the planner can build")
ap.add_argument("--nacc", action="store_true", help="accuracy/time of the accumulator-splitting levels on long chains")
ap.add_argument("--sweep-full", action="store_true", help="--sweep also over the stacked / plain weight planes and the row width")
ap.add_argument("--max-cout", type=int, default=0, help="restrict the timing / sweep to layers with cout <= this")
ap.add_argument("--model", default="fuseunet", choices=["fuseunet", "unet", "both"])
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--skip-check", action="store_true")
ap.add_argument("--skip-layers", action="store_true", help="skip the old-vs-halo per-layer timing")
ap.add_argument("--dgrad", action="store_true", help="also sweep the dgrad role of every layer (cin <-> cout)")
ap.add_argument("--batches", nargs="+", type=int, default=[], help="--sweep-full over several batch sizes (one JSON)")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
NAMES = {1: "tf32x2", 2: "bf16", 3: "f16x2"}
TOL = {1: 3e-5, 2: 3e-2, 3: 3e-5}


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def plan(fmt, cin, cout, N, H, W):
    out = (C.c_int * 9)()
    if A.lib.aide_conv3x3_plan_info(fmt, cin, cout, N, H, W, out):
        return None
    return dict(BN=out[0], MB=out[1], nacc=out[2], nbuf=out[3], rb=out[4], aS=out[5], bS=out[6], stack=out[8] & 1,
                res=(out[8] >> 1) & 1, occ=4 if out[8] & 8 else 2 if out[8] & 4 else 1)


def check(fmt, N, H, W, cin, cout):
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(N, cin, H, W, generator=g).to(dev)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    ref = F.conv2d(x, w, b, padding=1)
    a = ops.from_nchw(x, fmt)
    z, part = ops.conv3x3(a, w, b, stats=True)
    torch.cuda.synchronize()
    got = ops.nhwc_to_nchw(z)
    e = rel(got, ref)
    es = rel(part.sum(0)[0], ref.sum((0, 2, 3)))
    return e, es


SHAPES = [(1, 16, 8, 32, 32), (2, 16, 16, 64, 64), (1, 32, 32, 64, 64), (2, 8, 8, 128, 256), (1, 20, 12, 64, 32),
          (3, 24, 40, 96, 96), (1, 16, 128, 32, 64), (2, 32, 32, 256, 512), (4, 64, 64, 64, 256), (1, 16, 16, 1024, 512)]

ok_mode = {}
for mode in (() if args.skip_check else (0,)):
    for fmt in args.fmts:
        worst, bad = 0.0, 0
        for shp in SHAPES:
            try:
                e, es = check(fmt, *shp)
            except Exception as ex:  # noqa: BLE001
                print(f"[ERR] mode {mode} {NAMES[fmt]} {shp}: {type(ex).__name__}: {ex}", flush=True)
                bad += 1
                continue
            worst = max(worst, e)
            flag = "ok " if e < TOL[fmt] and es < 1e-3 else "BAD"
            if flag == "BAD":
                bad += 1
            print(f"[{flag}] desc_mode {mode} {NAMES[fmt]:7s} {str(shp):28s} rel {e:.2e} stats {es:.2e} "
                  f"plan {plan(fmt, shp[3], shp[4], shp[0], shp[1], shp[2])}", flush=True)
        ok_mode[(mode, fmt)] = bad == 0
        print(f"== desc_mode {mode} {NAMES[fmt]}: {'PASS' if bad == 0 else 'FAIL'} (worst rel {worst:.2e})", flush=True)
good = args.skip_check or [m for m in (0,) if all(ok_mode[(m, f)] for f in args.fmts)]
print("CHECK", "PASS" if good else "FAIL", flush=True)
if not good or args.no_timing:
    sys.exit(0 if good else 3)

# ---------------------------------------------------------------------------------------------- timing
from aide_b200 import engine as E  # noqa: E402

shapes = {}
for plan_net in ([E.plan_fuseunet(2)] if args.model != "unet" else []) + ([E.plan_unet(2)] if args.model != "fuseunet" else []):
    for u in plan_net.units:
        if not u.first and (not args.max_cout or u.cout <= args.max_cout):
            key = (u.cin, u.cout, args.size >> u.level)
            shapes[key] = shapes.get(key, 0) + 1
            if args.dgrad and u.cin % 32 == 0:
                key = (u.cout, u.cin, args.size >> u.level)
                shapes[key] = shapes.get(key, 0) + 1
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
B = args.batch
result = {}
for fmt in ([] if args.skip_layers else args.fmts):
    tot = {0: 0.0, 1: 0.0}
    tot_flop = 0.0
    rows = []
    for (cin, cout, hw), count in sorted(shapes.items(), key=lambda kv: -kv[0][2]):
        x = ops.Act(B, hw, hw, cin, fmt, dev)
        x.planes.normal_()
        w = torch.randn(cout, cin, 3, 3, device=dev) * (9 * cin) ** -0.5
        bias = torch.zeros(cout, device=dev)
        w0, w1, keep = ops.weight_prep(w, fmt)
        z = torch.empty((B, hw, hw, cout), dtype=torch.float32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        ms = {}
        for halo in (0, 1):
            os.environ["AIDE_CONV_HALO"] = str(halo)
            nrows = A.lib.aide_conv3x3_stat_rows(fmt, cin, cout, B, hw, hw)
            part = torch.empty((nrows, 2, cout), dtype=torch.float32, device=dev)

            def launch():
                ops.call("aide_conv3x3_fwd", fmt, x.p0, x.p1, x.C, 0, cin, w0, w1, bias.data_ptr(), z.data_ptr(), cout,
                         0, cout, B, hw, hw, part.data_ptr(), st)
            for _ in range(2):
                launch()
            t = 0.0
            for _ in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); launch(); e1.record()
                torch.cuda.synchronize()
                t += e0.elapsed_time(e1)
            ms[halo] = t / 4
        os.environ["AIDE_CONV_HALO"] = "1"
        flop = 2.0 * B * hw * hw * cout * cin * 9
        tot_flop += flop * count
        for h in (0, 1):
            tot[h] += ms[h] * count
        pl = plan(fmt, cin, cout, B, hw, hw)
        rows.append(dict(cin=cin, cout=cout, hw=hw, n=count, ms_old=round(ms[0], 4), ms_halo=round(ms[1], 4),
                         tf_old=round(flop / ms[0] / 1e9, 1), tf_halo=round(flop / ms[1] / 1e9, 1), plan=pl))
        print(f"{NAMES[fmt]:7s} {cin:4d}->{cout:3d} @{hw:3d} x{count}: old {ms[0]:.4f} ms ({flop / ms[0] / 1e9:6.1f} TF)  "
              f"halo {ms[1]:.4f} ms ({flop / ms[1] / 1e9:6.1f} TF)  {pl}", flush=True)
        del x, z
    print(f"== {NAMES[fmt]} forward conv total: old {tot[0]:.3f} ms ({tot_flop / tot[0] / 1e9:.1f} TF)  "
          f"halo {tot[1]:.3f} ms ({tot_flop / tot[1] / 1e9:.1f} TF)", flush=True)
    result[NAMES[fmt]] = dict(layers=rows, total_ms_old=tot[0], total_ms_halo=tot[1],
                              tflops_old=tot_flop / tot[0] / 1e9, tflops_halo=tot_flop / tot[1] / 1e9)
if args.json:
    with open(args.json, "w") as f:
        json.dump(result, f, indent=1)


# ---------------------------------------------------------------------------------------------- (BN, MB) sweep
def time_layer(fmt, cin, cout, hw, reps=3):
    x = ops.Act(B, hw, hw, cin, fmt, dev)
    x.planes.normal_()
    w = torch.randn(cout, cin, 3, 3, device=dev) * (9 * cin) ** -0.5
    bias = torch.zeros(cout, device=dev)
    w0, w1, keep = ops.weight_prep(w, fmt)
    z = torch.empty((B, hw, hw, cout), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    nrows = A.lib.aide_conv3x3_stat_rows(fmt, cin, cout, B, hw, hw)
    part = torch.empty((nrows, 2, cout), dtype=torch.float32, device=dev)

    def launch():
        ops.call("aide_conv3x3_fwd", fmt, x.p0, x.p1, x.C, 0, cin, w0, w1, bias.data_ptr(), z.data_ptr(), cout,
                 0, cout, B, hw, hw, part.data_ptr(), st)
    launch()
    t = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        torch.cuda.synchronize()
        t += e0.elapsed_time(e1)
    return t / reps


if args.sweep_full:
    full = {}
    for fmt, B in [(f, b) for f in args.fmts for b in (args.batches or [args.batch])]:
        for (cin, cout, hw), count in sorted(shapes.items(), key=lambda kv: -kv[0][2]):
            flop = 2.0 * B * hw * hw * cout * cin * 9
            KEYS = ("AIDE_CONV_BN", "AIDE_CONV_MB", "AIDE_CONV_STACK", "AIDE_CONV_RB", "AIDE_CONV_WRES", "AIDE_CONV_OCC")
            for k in KEYS:
                os.environ.pop(k, None)
            dflt = plan(fmt, cin, cout, B, hw, hw)
            res = []
            for bn in (32, 64, 128, 256):
                if cout % bn:
                    continue
                for mb in (1, 2, 4):
                    for stack in (0, 1):
                        for rb in (64, 128):
                            for occ in (1, 2, 4):
                                wres = 0
                                os.environ.update(AIDE_CONV_BN=str(bn), AIDE_CONV_MB=str(mb), AIDE_CONV_STACK=str(stack),
                                                  AIDE_CONV_RB=str(rb), AIDE_CONV_WRES="0", AIDE_CONV_OCC=str(occ))
                                pl = plan(fmt, cin, cout, B, hw, hw)
                                if pl is None or pl["BN"] != bn or pl["MB"] != mb or pl["stack"] != stack or pl["rb"] != rb \
                                        or pl["occ"] != occ:
                                    continue
                                try:
                                    ms = time_layer(fmt, cin, cout, hw)
                                except Exception as ex:  # noqa: BLE001
                                    print(f"[ERR] sweep {cin}->{cout}@{hw} BN={bn} MB={mb} stack={stack} rb={rb} res={wres}: {ex}",
                                          flush=True)
                                    continue
                                res.append((ms, bn, mb, stack, rb, pl))
            for k in KEYS:
                os.environ.pop(k, None)
            res.sort(key=lambda r: r[0])
            line = "  ".join(f"BN{bn}/MB{mb}/s{stack}/r{rb}/o{pl['occ']}/a{pl['nacc']}b{pl['nbuf']}A{pl['aS']}B{pl['bS']}:{flop / ms / 1e9:.0f}"
                             for ms, bn, mb, stack, rb, pl in res[:12])
            t_def = time_layer(fmt, cin, cout, hw)
            print(f"SWEEPF {NAMES[fmt]:7s} {cin:4d}->{cout:3d} @{hw:3d} default {dflt} {flop / t_def / 1e9:.0f} TF | {line}",
                  flush=True)
            full[f"{NAMES[fmt]}:{cin}:{cout}:{hw}:{B}"] = dict(
                default=dflt, default_ms=round(t_def, 4), gflop=round(flop / 1e9, 2),
                configs=[dict(ms=round(ms, 4), BN=bn, MB=mb, stack=stack, rb=rb, res=pl["res"], occ=pl["occ"], nacc=pl["nacc"], nbuf=pl["nbuf"],
                              aS=pl["aS"], bS=pl["bS"]) for ms, bn, mb, stack, rb, pl in res])
    if args.json:
        with open(args.json.replace(".json", "_full.json"), "w") as f:
            json.dump(full, f, indent=1)

if args.sweep:
    sweep = {}
    for fmt in args.fmts:
        for (cin, cout, hw), count in sorted(shapes.items(), key=lambda kv: -kv[0][2]):
            flop = 2.0 * B * hw * hw * cout * cin * 9
            os.environ.pop("AIDE_CONV_BN", None)
            os.environ.pop("AIDE_CONV_MB", None)
            dflt = plan(fmt, cin, cout, B, hw, hw)
            res = []
            for bn in (32, 64, 128, 256):
                if cout % bn:
                    continue
                for mb in (1, 2, 4):
                    os.environ["AIDE_CONV_BN"], os.environ["AIDE_CONV_MB"] = str(bn), str(mb)
                    pl = plan(fmt, cin, cout, B, hw, hw)
                    if pl is None or pl["BN"] != bn or pl["MB"] != mb:
                        continue
                    try:
                        ms = time_layer(fmt, cin, cout, hw)
                    except Exception as ex:  # noqa: BLE001
                        print(f"[ERR] sweep {NAMES[fmt]} {cin}->{cout}@{hw} BN={bn} MB={mb}: {ex}", flush=True)
                        continue
                    res.append((ms, bn, mb, pl))
            os.environ.pop("AIDE_CONV_BN", None)
            os.environ.pop("AIDE_CONV_MB", None)
            res.sort(key=lambda r: r[0])
            line = "  ".join(f"BN{bn}/MB{mb}/a{pl['nacc']}b{pl['nbuf']}r{pl['rb']}:{flop / ms / 1e9:.0f}" for ms, bn, mb, pl in res)
            print(f"SWEEP {NAMES[fmt]:7s} {cin:4d}->{cout:3d} @{hw:3d} default BN{dflt['BN']}/MB{dflt['MB']} | {line}", flush=True)
            sweep[f"{NAMES[fmt]}:{cin}:{cout}:{hw}"] = [(round(ms, 4), bn, mb) for ms, bn, mb, _ in res]
    if args.json:
        with open(args.json.replace(".json", "_sweep.json"), "w") as f:
            json.dump(sweep, f, indent=1)

# ---------------------------------------------------------------------------------------------- accumulator splitting
if args.nacc:
    for fmt in [f for f in args.fmts if f != 2]:
        for cin, cout, hw in ((1024, 512, 32), (512, 256, 64), (256, 128, 128)):
            g = torch.Generator().manual_seed(7)
            x = torch.randn(2, cin, hw, hw, generator=g).to(dev)
            w = (torch.randn(cout, cin, 3, 3, generator=g) * (9 * cin) ** -0.5).to(dev)
            ref = F.conv2d(x.double(), w.double(), None, padding=1)
            for cap in (1, 2, 4):
                os.environ["AIDE_CONV_NACC_MAX"] = str(cap)
                a = ops.from_nchw(x, fmt)
                z, _ = ops.conv3x3(a, w, None)
                torch.cuda.synchronize()
                got = ops.nhwc_to_nchw(z).double()
                err = got - ref
                ms = time_layer(fmt, cin, cout, hw)
                print(f"NACC {NAMES[fmt]:7s} {cin}->{cout}@{hw} cap {cap}: max {err.abs().max().item() / ref.abs().max().item():.2e} "
                      f"rms {err.pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item():.2e} "
                      f"bias {err.mean().item() / ref.abs().mean().item():+.1e}  {ms:.4f} ms plan {plan(fmt, cin, cout, B, hw, hw)}",
                      flush=True)
            os.environ.pop("AIDE_CONV_NACC_MAX", None)
