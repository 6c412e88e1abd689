"""One stacked pseudo-label forward (4 views x B) and one train forward + backward of a fuseunet, eager launches, for
ncu captures of the HBM-bound kernels:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"bn_relu|upsample2x|conv1x1|nchw_to_nhwc|zero_insert|adam|loss_" --csv --log-file out.csv \
        python tools/profile_step.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aide_b200 as A  # noqa: E402

B = int(os.environ.get("PROF_B", "8"))
S = int(os.environ.get("PROF_S", "256"))
dev = torch.device("cuda:0")
torch.manual_seed(2)
net = A.fuseunet(num_classes=2, mode="parity").to(dev).train()
g = torch.Generator(device="cuda").manual_seed(1)
views = [tuple(torch.randn(B, 3, S, S, device=dev, generator=g) for _ in range(2)) for _ in range(4)]
x = tuple(torch.randn(B, 3, S, S, device=dev, generator=g) for _ in range(2))
t = (torch.rand(B, S, S, device=dev, generator=g) < 0.08).long()
with torch.no_grad():
    net._engine_forward_grouped(views)
y = net(*x)
A.CEMDiceLoss()(y, t).backward()
torch.cuda.synchronize()
print("done", flush=True)
