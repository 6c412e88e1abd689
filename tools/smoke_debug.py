"""Gradient errors of the smoke() configuration (B=2, 32x32, n_clean=1) per parameter, under kernel-selection switches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aide_b200 as A  # noqa: E402
from oracle import aide_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
S = int(os.environ.get("DBG_SIZE", "32"))
B = int(os.environ.get("DBG_B", "2"))
(x1, x2), t1, t2, augs = O.synthetic_batch(B, S, S, seed=1234, n_aug=2)
torch.manual_seed(2)
p1 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
p2 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
r = O.aide_step(O.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, 0.25, n_clean=1)
d = lambda t: t.to(dev)
rel = lambda a, b: ((a.detach().cpu() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-30)).item()


def run(mode, env):
    for k in ("AIDE_WGRAD_HALO", "AIDE_CONV_STACK", "AIDE_CONV_HALO"):
        os.environ.pop(k, None)
    os.environ.update(env)
    torch.manual_seed(2)
    n1 = A.fuseunet(num_classes=2, mode=mode).to(dev).train()
    n2 = A.fuseunet(num_classes=2, mode=mode).to(dev).train()
    q1, w1 = A.pseudo_label([n1(d(a), d(b)).detach() for a, b in augs])
    q2, w2 = A.pseudo_label([n2(d(a), d(b)).detach() for a, b in augs])
    o1, o2 = n1(d(x1), d(x2)), n2(d(x1), d(x2))
    m = A.coteach_step(o1, o2, d(t1), d(t2), q1, w1, q2, w2, 0.25, n_clean=1)
    m["loss1"].backward(retain_graph=True)
    m["loss2"].backward()
    torch.cuda.synchronize()
    errs = []
    for name, p in n1.named_parameters():
        if name.endswith("bias") and ("conv" in name or "bilinear_up.1" in name):
            continue                                    # pre-BN conv biases: analytically zero gradients
        errs.append((rel(p.grad, r["grads1"][name]), name))
    errs.sort(reverse=True)
    names = [n for _, n in errs]
    fa = torch.cat([dict(n1.named_parameters())[n].grad.detach().cpu().double().flatten() for n in names])
    fb = torch.cat([r["grads1"][n].detach().double().flatten() for n in names])
    cos = torch.nn.functional.cosine_similarity(fa, fb, dim=0).item()
    lw = rel(n1.last_conv1.weight.grad, r["grads1"]["last_conv1.weight"])
    print(f"   1-cos {1 - cos:.2e}  last_conv1.weight {lw:.2e}  loss1 {abs(m['loss1'].item() - r['loss1'].item()):.1e} "
          f"idx_equal {torch.equal(m['indx1'].cpu(), r['indx1']) and torch.equal(m['indx2'].cpu(), r['indx2'])}")
    print(f"mode={mode} env={env}: logits {rel(o1, r['out1']):.2e}  worst grads: " +
          ", ".join(f"{n}:{e:.1e}" for e, n in errs[:4]), flush=True)


run("parity", {})
run("parity_tf32", {})
run("exact", {})
run("fast", {})
