"""GPU probe: whole-network logits / gradient error of the engine against the CPU oracle in fp32 AND fp64.

Why: the fp32 reference itself deviates from exact arithmetic by percent-level amounts in some weight gradients
(ReLU masks flip for pre-activations within rounding error of zero; each flip moves one term of a dW sum), so
"gradient parity" has to be read against that band.  Prints, per tensor, engine-vs-fp64 and oracle32-vs-fp64."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aide_b200 as A  # noqa: E402
from oracle import aide_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 8)


def rel(a, b):
    return ((a.detach().cpu().double() - b.detach().cpu().double()).abs().max()
            / b.detach().cpu().double().abs().max().clamp_min(1e-300)).item()


def prebn_bias(n):
    return n.endswith(("block.conv1.bias", "block.conv2.bias", "bilinear_up.1.bias"))


def oracle_grads(b, h, w, dt):
    (x1, x2), t1, t2, _ = O.synthetic_batch(b, h, w, seed=1234)
    torch.manual_seed(2)
    p0 = O.init_fuseunet(2)
    p = {k: ((v.detach().to(dt).requires_grad_() if not O.is_buffer(k) else v.detach().to(dt))
             if v.is_floating_point() else v.clone()) for k, v in p0.items()}
    y = O.fuseunet_forward(p, x1.to(dt), x2.to(dt), training=True)
    loss = O.ce_dice_mean(y, t2)
    names = [k for k in p if not O.is_buffer(k)]
    return y.detach(), dict(zip(names, torch.autograd.grad(loss, [p[k] for k in names])))


shapes = [(2, 32, 32), (4, 64, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "full":
    shapes.append((4, 256, 256))
for (b, h, w) in shapes:
    t0 = time.time()
    y32, g32 = oracle_grads(b, h, w, torch.float32)
    y64, g64 = oracle_grads(b, h, w, torch.float64) if h <= 64 else (y32, g32)
    print(f"=== B={b} {h}x{w}  (oracle {time.time() - t0:.1f}s)  oracle32-vs-64 logits {rel(y32, y64):.2e}", flush=True)
    (x1, x2), t1, t2, _ = O.synthetic_batch(b, h, w, seed=1234)
    for mode in ("exact", "parity", "fast"):
        torch.manual_seed(2)
        net = A.fuseunet(num_classes=2, mode=mode).to(dev).train()
        y = net(x1.to(dev), x2.to(dev))
        A.CEMDiceLoss([1., 1.], [1., 1.], [1., 1.])(y, t2.to(dev)).backward()
        torch.cuda.synchronize()
        flips = int((y.argmax(1).cpu() != y64.argmax(1)).sum())
        rows = []
        for n, prm in net.named_parameters():
            if prebn_bias(n):
                continue
            rows.append((rel(prm.grad, g64[n]), rel(g32[n], g64[n]), rel(prm.grad, g32[n]), n))
        worst = sorted(rows, reverse=True)[:4]
        nbad = sum(1 for r in rows if r[0] > 1e-3)
        print(f"  {mode:6s} logits vs64 {rel(y, y64):.2e} vs32 {rel(y, y32):.2e} argmax flips {flips}; "
              f"grads: {nbad}/{len(rows)} tensors > 1e-3 vs64; worst (eng-vs-64, orc32-vs-64, eng-vs-32):", flush=True)
        for r in worst:
            print(f"      {r[0]:.2e} {r[1]:.2e} {r[2]:.2e}  {r[3]}")
        del net
