"""Launch the tcgen05 conv3x3 kernel alone on a few fuseunet layer shapes (for ncu --set full captures).

    ncu --set full --clock-control none --import-source on -k regex:conv3x3_fwd_tc -o gpurun_out/conv_full \
        python tools/profile_conv.py --fmt 3 --layers 512,256,64 128,64,256
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aide_b200 as A  # noqa: E402
from aide_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--fmt", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--layers", nargs="+", default=["512,256,64"])          # cin,cout,hw[,BN,MB] (forced tiling)
args = ap.parse_args()
dev = torch.device("cuda:0")
for spec in args.layers:
    vals = [int(v) for v in spec.split(",")]
    cin, cout, hw = vals[:3]
    for k in ("AIDE_CONV_BN", "AIDE_CONV_MB"):
        os.environ.pop(k, None)
    if len(vals) >= 5:
        os.environ["AIDE_CONV_BN"], os.environ["AIDE_CONV_MB"] = str(vals[3]), str(vals[4])
    x = ops.Act(args.batch, hw, hw, cin, args.fmt, dev)
    x.planes.normal_()
    w = torch.randn(cout, cin, 3, 3, device=dev) * (9 * cin) ** -0.5
    bias = torch.zeros(cout, device=dev)
    w0, w1, keep = ops.weight_prep(w, args.fmt)
    z = torch.empty((args.batch, hw, hw, cout), dtype=torch.float32, device=dev)
    rows = A.lib.aide_conv3x3_stat_rows(args.fmt, cin, cout, args.batch, hw, hw)
    part = torch.empty((rows, 2, cout), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(args.reps):
        ops.call("aide_conv3x3_fwd", args.fmt, x.p0, x.p1, x.C, 0, cin, w0, w1, bias.data_ptr(), z.data_ptr(), cout, 0,
                 cout, args.batch, hw, hw, part.data_ptr(), st)
    torch.cuda.synchronize()
    print("ran", spec, flush=True)
