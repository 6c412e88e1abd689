"""Algorithmic GB/s of the HBM-bound kernels, one launch each at the shapes of the step's largest levels, L2 flushed
before every timed launch (CUDA events).  Peak = MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth).

    python tools/hbm_probe.py [--batch 8] [--json out.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aide_b200 import ops  # noqa: E402
from aide_b200._lib import call, lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--json", default="")
args = ap.parse_args()
dev = torch.device("cuda:0")
FMT = 3
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    PEAK = 6544.0
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
rows_out = []


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ms = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    return ms / reps


def report(name, nbytes, fn):
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    rows_out.append(dict(kernel=name, mbytes=round(nbytes / 1e6, 1), ms=round(ms, 4), gbs=round(gbs), frac=round(gbs / PEAK, 3)))
    print(f"{name:58s} {nbytes / 1e6:9.1f} MB {ms:8.4f} ms {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of measured copy peak", flush=True)


none = (None, None, 0, 0)
for (B, hw, Cc) in ((4 * args.batch, 256, 64), (4 * args.batch, 128, 128), (args.batch, 256, 64), (args.batch, 64, 256)):
    N = B
    z = torch.randn(N, hw, hw, Cc, device=dev)
    ss = torch.randn(2, Cc, device=dev)
    mr = torch.rand(2, Cc, device=dev) + 0.5
    y = ops.Act(N, hw, hw, Cc, FMT, dev)
    pl = ops.Act(N, hw // 2, hw // 2, Cc, FMT, dev)
    el = N * hw * hw * Cc
    tag = f"[{N}x{hw}x{hw}x{Cc}]"
    report("bn_relu_apply " + tag, el * (4 + 4), lambda: call("aide_bn_relu_apply", FMT, z.data_ptr(), N, hw, hw, Cc, ss.data_ptr(), *y.view(), *none, *none, st))
    report("bn_relu_apply + pool " + tag, el * (4 + 4 + 1), lambda: call("aide_bn_relu_apply", FMT, z.data_ptr(), N, hw, hw, Cc, ss.data_ptr(), *y.view(), *pl.view(), *none, st))
    if N == args.batch:
        # backward kernels at the train batch
        d0 = torch.randn(N, hw, hw, Cc, device=dev)
        p0 = torch.randn(N, hw // 2, hw // 2, Cc, device=dev)
        rows = lib.aide_bn_bwd_rows(N, hw, hw, Cc)
        g = torch.empty(N, hw, hw, Cc, device=dev)
        part1 = torch.empty(rows, 2, Cc, device=dev)
        part2 = torch.empty(rows, Cc, device=dev)
        gs = torch.zeros(4, device=dev)
        tk = torch.zeros(16, dtype=torch.int32, device=dev)
        dptr, dct, dco = (C.c_void_p * 3)(d0.data_ptr()), (C.c_int * 3)(Cc), (C.c_int * 3)(0)
        pptr, pct, pco = (C.c_void_p * 3)(p0.data_ptr()), (C.c_int * 3)(Cc), (C.c_int * 3)(0)
        report("bn_relu_bwd_reduce (1 direct) " + tag, el * 12,
               lambda: call("aide_bn_relu_bwd_reduce", z.data_ptr(), ss.data_ptr(), mr.data_ptr(), N, hw, hw, Cc, dptr, dct, dco, 1,
                            pptr, pct, pco, 0, g.data_ptr(), part1.data_ptr(), gs.data_ptr(), st))
        report("bn_relu_bwd_reduce (1 direct + 1 pooled) " + tag, el * 13,
               lambda: call("aide_bn_relu_bwd_reduce", z.data_ptr(), ss.data_ptr(), mr.data_ptr(), N, hw, hw, Cc, dptr, dct, dco, 1,
                            pptr, pct, pco, 1, g.data_ptr(), part1.data_ptr(), gs.data_ptr(), st))
        dz = ops.Act(N, hw, hw, Cc, FMT, dev)
        small = torch.empty(3, Cc, device=dev)
        gamma = torch.rand(Cc, device=dev) + 0.5
        report("bn_relu_bwd_apply " + tag, el * 12,
               lambda: call("aide_bn_relu_bwd_apply", FMT, g.data_ptr(), z.data_ptr(), mr.data_ptr(), gamma.data_ptr(), part1.data_ptr(),
                            rows, N, hw, hw, Cc, dz.p0, dz.p1, small[1].data_ptr(), small[0].data_ptr(), small[2].data_ptr(),
                            part2.data_ptr(), gs.data_ptr(), gs[1:].data_ptr(), tk.data_ptr(), st))
        del d0, p0, g, dz
    del z, y, pl
# upsample (decoder: level 1 -> 0 carries 128 channels, level 2 -> 1 256 channels)
for (N, h, Cc) in ((4 * args.batch, 128, 128), (4 * args.batch, 64, 256), (args.batch, 128, 128)):
    x = ops.Act(N, h, h, Cc, FMT, dev)
    x.planes.normal_()
    yy = ops.Act(N, 2 * h, 2 * h, Cc, FMT, dev)
    el_lo = N * h * h * Cc
    tag = f"[{N}x{h}x{h}x{Cc} -> x2]"
    report("upsample2x_fwd " + tag, el_lo * 4 * 5, lambda: call("aide_upsample2x_fwd", FMT, *x.view(), *yy.view(), N, h, h, Cc, st))
    if N == args.batch:
        dhi = torch.randn(N, 2 * h, 2 * h, Cc, device=dev)
        dlo = torch.empty(N, h, h, Cc, device=dev)
        report("upsample2x_bwd " + tag, el_lo * 4 * 5, lambda: call("aide_upsample2x_bwd", dhi.data_ptr(), Cc, 0, dlo.data_ptr(), N, h, h, Cc, st))
        del dhi, dlo
    del x, yy
# module boundary and head
for N in (4 * args.batch, args.batch):
    src = torch.randn(N, 3, 256, 256, device=dev)
    dst = torch.empty(N, 256, 256, 3, device=dev)
    report(f"nchw_to_nhwc (3 channels) [{N}x3x256x256]", N * 3 * 65536 * 8,
           lambda: call("aide_nchw_to_nhwc", 0, src.data_ptr(), dst.data_ptr(), None, 3, 0, N, 3, 256, 256, st))
    x = ops.Act(N, 256, 256, 64, FMT, dev)
    x.planes.normal_()
    w = torch.randn(2, 64, device=dev)
    b = torch.randn(2, device=dev)
    out = torch.empty(N, 2, 256, 256, device=dev)
    report(f"conv1x1_fwd (head) [{N}x256x256x64]", N * 65536 * (64 * 4 + 8),
           lambda: call("aide_conv1x1_fwd", FMT, *x.view(), 64, w.data_ptr(), b.data_ptr(), out.data_ptr(), 2, N, 256, 256, st))
    if N == args.batch:
        rows = lib.aide_conv1x1_bwd_rows(N, 256, 256, 64)
        part = torch.empty(rows, 2 * 64 + 2, device=dev)
        dx = torch.empty(N, 256, 256, 64, device=dev)
        dwdb = torch.empty(2 * 64 + 2, device=dev)
        dl = torch.randn(N, 2, 256, 256, device=dev)
        report(f"conv1x1_bwd (head) [{N}x256x256x64]", N * 65536 * (64 * 4 + 64 * 4 + 8),
               lambda: call("aide_conv1x1_bwd", FMT, *x.view(), 64, w.data_ptr(), dl.data_ptr(), 2, N, 256, 256, dx.data_ptr(),
                            dwdb.data_ptr(), part.data_ptr(), st))
n = 26_675_076
p, g_, m, v, vm = (torch.randn(n, device=dev) for _ in range(5))
stepc, bc = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(2, device=dev)
report("adam_amsgrad (26.7 M parameters)", n * 4 * 9,
       lambda: call("aide_adam_amsgrad_dev", p.data_ptr(), g_.data_ptr(), m.data_ptr(), v.abs_().data_ptr(), vm.abs_().data_ptr(), n,
                    1e-4, 0.9, 0.999, 1e-8, stepc.data_ptr(), bc.data_ptr(), 1.0, None, st))
# BatchNorm statistics fold + finalize over the conv kernel's partial rows (4 rows per 8x16-pixel tile), stacked forward
# of 5 groups of `batch` slices
for (hw, Cc) in ((256, 64), (128, 128), (64, 256)):
    rows = 4 * args.batch * ((hw + 7) // 8) * ((hw + 15) // 16)
    G = 5
    parts0 = torch.randn(G, rows, 2, Cc, device=dev)
    parts = parts0.clone()
    gam, bet = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    rm, rv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    tk = torch.zeros(lib.aide_bn_ticket_slots(Cc), dtype=torch.int32, device=dev)
    ss = torch.empty(G, 2, Cc, device=dev)
    mr = torch.empty(G, 2, Cc, device=dev)

    def fold():
        call("aide_bn_finalize_grouped", parts.data_ptr(), rows, G, Cc, float(args.batch * hw * hw), gam.data_ptr(),
             bet.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5, 1, ss.data_ptr(), mr.data_ptr(), tk.data_ptr(), st)
    report(f"bn_stat_fold_finalize [5 x {rows} rows x 2 x {Cc}]", parts.numel() * 4, fold)

# first layer (Cin = 3, fp32 CUDA cores): 27 FMAs per output element, 12 B in + 4*Cout B out per pixel
import aide_b200 as A  # noqa: E402
for N in (5 * args.batch, args.batch):
    hw, cout = 256, 32
    x = ops.Act(N, hw, hw, 3, 0, dev)
    x.planes.normal_()
    w = torch.randn(cout, 3, 3, 3, device=dev) * 0.2
    bias = torch.zeros(cout, device=dev)
    w0, w1, keep = ops.weight_prep(w, 0)
    z = torch.empty((N, hw, hw, cout), dtype=torch.float32, device=dev)
    rows = A.lib.aide_conv3x3_stat_rows(0, 3, cout, N, hw, hw)
    part = torch.empty((rows, 2, cout), dtype=torch.float32, device=dev)
    report(f"conv3x3 first layer 3->32 [{N}x256x256]", N * hw * hw * (12 + 4 * cout),
           lambda: call("aide_conv3x3_fwd", 0, x.p0, x.p1, x.C, 0, 3, w0, w1, bias.data_ptr(), z.data_ptr(), cout, 0, cout,
                        N, hw, hw, part.data_ptr(), st))

if args.json:
    json.dump(dict(peak_gbs=PEAK, rows=rows_out), open(args.json, "w"), indent=1)
