"""Free-running trajectory comparison (second finding; section 8c "100-step check").

Runs the AIDE step for N steps from the same initial weights and the same batches three ways, each with its own
Adam-amsgrad state and NO re-synchronisation:
    A  the oracle (CPU, all threads)
    B  the oracle again with ONE thread  -> the reference's own noise band (same code, other fp32 summation order)
    C  the B200 engine (parity mode)
and reports, per step, |Dice_fn/B| differences and whether the small-loss index sets agree.  The engine is expected to
sit inside the band the oracle has against itself; the pass/fail parity bar (1e-4, bit-exact index sets) is evaluated
teacher-forced (tests/test_gpu_network.py::test_teacher_forced_training_steps), because free-running trajectories are
chaotic even for the reference against itself.

    python tools/free_running.py [--steps 60] [--size 64] [--batch 4]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import aide_oracle as O  # noqa: E402  (test infrastructure)

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--batch", type=int, default=4)
args = ap.parse_args()
B, S, N = args.batch, args.size, args.steps


def batches():
    for k in range(N):
        yield O.synthetic_batch(B, S, S, seed=5000 + k, n_aug=4)


def run_oracle(threads):
    torch.set_num_threads(threads)
    torch.manual_seed(2)
    p1 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
    p2 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
    s1, s2, out = {}, {}, []
    for k, ((x1, x2), t1, t2, augs) in enumerate(batches(), 1):
        r = O.aide_step(O.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, 0.25)
        O.adam_amsgrad_step(p1, r["grads1"], s1, k)
        O.adam_amsgrad_step(p2, r["grads2"], s2, k)
        out.append((r["dice1"].item() / B, r["dice2"].item() / B, r["indx1"].tolist(), r["indx2"].tolist(),
                    r["loss1"].item(), r["loss2"].item()))
    return out


def run_engine():
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=True)
    d = lambda t: t.to(dev)
    out = []
    for (x1, x2), t1, t2, augs in batches():
        m = tr.step((d(x1), d(x2)), d(t1), d(t2), [(d(a), d(b)) for a, b in augs], 0.25)
        out.append((m["dice1"].item() / B, m["dice2"].item() / B, m["indx1"].tolist(), m["indx2"].tolist(),
                    m["loss1"].item(), m["loss2"].item()))
    return out


def compare(a, b, what):
    dd = [max(abs(x[0] - y[0]), abs(x[1] - y[1])) for x, y in zip(a, b)]
    mism = [k for k, (x, y) in enumerate(zip(a, b)) if x[2] != y[2] or x[3] != y[3]]
    dl = [max(abs(x[4] - y[4]), abs(x[5] - y[5])) for x, y in zip(a, b)]
    first = next((k for k, v in enumerate(dd) if v > 0), None)
    print(f"{what}: max |dDice_fn/B| {max(dd):.2e} (first non-zero at step {first}), index sets differ at {len(mism)} of "
          f"{len(a)} steps (first: {mism[0] if mism else None}), max |dloss| {max(dl):.2e}", flush=True)


cores = os.cpu_count() or 1
A_ = run_oracle(cores)
B_ = run_oracle(1)
compare(A_, B_, f"oracle {cores} threads vs oracle 1 thread  (reference self-noise band)")
if torch.cuda.is_available():
    C_ = run_engine()
    compare(C_, A_, f"engine (parity, B200) vs oracle {cores} threads")
    compare(C_, B_, "engine (parity, B200) vs oracle 1 thread  ")
print(f"config: {N} steps, batch {B}, {S}x{S}, two fuseunets, Adam-amsgrad lr 1e-4, rate 0.25")
