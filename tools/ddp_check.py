"""torchrun --nproc-per-node 2 tools/ddp_check.py -- data-parallel semantics of AideTrainer on 2 GPUs.

After one step on different local batches every rank must hold the same weights, and they must equal what a
single process obtains by averaging the two shards' gradients (BatchNorm statistics and the small-loss selection
are local per rank)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aide_b200.trainer import AideTrainer  # noqa: E402
from oracle import aide_oracle as O  # noqa: E402  (test infrastructure: synthetic batches only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = 3, 32
d = lambda t: t.to(dev)


def shard(r):
    (x1, x2), t1, t2, augs = O.synthetic_batch(B, S, S, seed=300 + r, n_aug=2)
    return (d(x1), d(x2)), d(t1), d(t2), [(d(a), d(b)) for a, b in augs]


# (0) the peer-memory all-reduce kernel alone: random data, ragged ranges, against the fixed-order sum it promises
from aide_b200.comm import PeerBuffers  # noqa: E402

pb = PeerBuffers(None, dev, [1 << 20, 4099 * 4], blocks=8)
g = torch.Generator().manual_seed(5)
data = [torch.randn(world, n, generator=g) for n in pb.sizes]
for rep in range(3):
    for i, (lo, hi) in ((0, (0, 1 << 20)), (1, (0, 4099 * 4)), (0, (1024, 70000)), (1, (8, 40))):
        pb.tensors[i].copy_(data[i][rank] * (rep + 1))
        torch.cuda.synchronize()
        dist.barrier()
        pb.all_reduce(i, lo, hi, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want = data[i][rank].clone() * (rep + 1)
        acc = data[i][0][lo:hi] * (rep + 1)
        for r in range(1, world):
            acc = acc + data[i][r][lo:hi] * (rep + 1)
        want[lo:hi] = acc
        assert torch.equal(pb.tensors[i].cpu(), want), (rep, i, lo, hi, (pb.tensors[i].cpu() - want).abs().max())
        dist.barrier()
print(f"[rank {rank}] aide_allreduce_p2p ok (bit-exact, rank-order sums)", flush=True)
pb.close()
# bandwidth of the collective alone at the size of one network's gradient (26.7 M floats), against NCCL
n_grad = 26675076 + (-26675076) % 4
for blocks in (8, 16, 32, 64):
    pb = PeerBuffers(None, dev, [n_grad], blocks=blocks)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        pb.all_reduce(0, 0, n_grad, st)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pb.all_reduce(0, 0, n_grad, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"aide_allreduce_p2p {n_grad * 4 / 1e6:.0f} MB, {blocks} CTAs: {ms:.3f} ms  algbw {n_grad * 4 / ms / 1e6:.0f} GB/s  "
              f"busbw {n_grad * 4 / ms / 1e6 * 2 * (world - 1) / world:.0f} GB/s", flush=True)
    pb.close()
t = torch.zeros(n_grad, device=dev)
for _ in range(3):
    dist.all_reduce(t)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    dist.all_reduce(t)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    ms = e0.elapsed_time(e1) / 10
    print(f"nccl all_reduce {n_grad * 4 / 1e6:.0f} MB: {ms:.3f} ms  algbw {n_grad * 4 / ms / 1e6:.0f} GB/s", flush=True)
del t
if os.environ.get("DDP_CHECK_KERNEL_ONLY", "0") == "1":       # any world size: the collective alone
    dist.barrier()
    if rank == 0:
        print("DDP_CHECK_KERNEL_OK", flush=True)
    sys.stdout.flush()
    os._exit(0)

for graph, comm in ((False, "p2p"), (True, "p2p"), (False, "nccl"), (True, "nccl")):
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph, comm=comm)
    tr.broadcast_parameters(0)
    p0 = tr.opt1.flat.clone()
    x, t1, t2, augs = shard(rank)
    for _ in range(2 if graph else 1):
        m = tr.step(x, t1, t2, augs, 0.25)
    torch.cuda.synchronize()
    # (1) replicas stay identical
    mine = tr.opt1.flat.clone()
    other = mine.clone()
    dist.broadcast(other, 0)
    assert torch.equal(mine, other), "ranks diverged"
    # (2) rank 0 reproduces the update from the two shards' gradients computed locally (world = 1 trainers)
    if rank == 0 and not graph:
        grads = []
        for r in range(world):
            ref = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=False, data_parallel=False)
            xs, a, b, au = shard(r)
            ref._step_eager(xs, a, b, au, 0.25)
            grads.append(ref.net1.last_grad_flat.clone())
        gmean = (grads[0] + grads[1]) / 2
        upd = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=False, data_parallel=False)
        upd.opt1.step(gmean)
        diff = (upd.opt1.flat - mine).abs().max().item()
        assert diff <= 2.1e-4, diff        # identical up to the all-reduce summation order on sign-like first updates
        frac_same = (upd.opt1.flat == mine).float().mean().item()
        assert frac_same > 0.99, frac_same
        assert not torch.equal(mine, p0)
    dist.barrier()
    print(f"[rank {rank}] graph={graph} comm={comm} ok", flush=True)
# (3) nn.DataParallel semantics (global_select=True): the per-image losses of all ranks are gathered, the small-loss
# sort and the clean / rest split are global, gradients add up.  Reference emulation with the oracle: every shard runs
# its own forward (replica-local BatchNorm statistics, as DataParallel does), outputs are concatenated and the loss of
# trainchaos_proposed_30cases1labeled.py:303-321 is evaluated on the global batch.
tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=False, global_select=True)
tr.broadcast_parameters(0)
x, t1, t2, augs = shard(rank)
m = tr.step(x, t1, t2, augs, 0.25)
torch.cuda.synchronize()
flat = tr.net1.last_grad_flat.clone()
other = flat.clone()
dist.broadcast(other, 0)
assert torch.equal(flat, other), "all-reduced gradients differ between ranks"
if rank == 0:
    torch.manual_seed(2)
    p1 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
    p2 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
    outs1, outs2, q1s, w1s, q2s, w2s, t1s, t2s = [], [], [], [], [], [], [], []
    for r in range(world):
        (x1, x2), a, b, au = O.synthetic_batch(B, S, S, seed=300 + r, n_aug=2)
        with torch.no_grad():
            a1 = [O.fuseunet_forward(p1, *v, training=True) for v in au]
            a2 = [O.fuseunet_forward(p2, *v, training=True) for v in au]
        q1, w1 = O.pseudo_label(a1)
        q2, w2 = O.pseudo_label(a2)
        outs1.append(O.fuseunet_forward(p1, x1, x2, training=True))
        outs2.append(O.fuseunet_forward(p2, x1, x2, training=True))
        for lst, v in ((q1s, q1), (w1s, w1), (q2s, q2), (w2s, w2), (t1s, a), (t2s, b)):
            lst.append(v)
    cat = lambda l: torch.cat(l, 0)
    res = O.coteach_losses(cat(outs1), cat(outs2), cat(t1s), cat(t2s), cat(q1s), cat(w1s), cat(q2s), cat(w2s), 0.25,
                           n_clean=2)
    names = [k for k in p1 if not O.is_buffer(k)]
    g1 = dict(zip(names, torch.autograd.grad(res["loss1"], [p1[k] for k in names])))
    assert torch.equal(m["indx1"].cpu(), res["indx1"]) and torch.equal(m["indx2"].cpu(), res["indx2"]), "global argsort"
    assert abs(m["loss1"].item() - res["loss1"].item()) < 2e-5 * max(1.0, abs(res["loss1"].item())), (m["loss1"].item(), res["loss1"].item())
    gl = tr.net1._glayout
    e = gl.view(flat, "last_conv1.weight").cpu()
    rel = ((e - g1["last_conv1.weight"]).abs().max() / g1["last_conv1.weight"].abs().max()).item()
    assert rel < 1e-3, rel
    live = [k for k in names if not k.endswith(("block.conv1.bias", "block.conv2.bias", "bilinear_up.1.bias"))]
    fa = torch.cat([gl.view(flat, k).cpu().double().flatten() for k in live])
    fb = torch.cat([g1[k].double().flatten() for k in live])
    gap = 1.0 - torch.nn.functional.cosine_similarity(fa, fb, dim=0).item()
    assert gap < 1e-3, gap
    print(f"[rank 0] global_select ok: loss1 {m['loss1'].item():.6f} vs {res['loss1'].item():.6f}, last_conv1 grad rel {rel:.1e}, "
          f"1-cos {gap:.1e}", flush=True)
dist.barrier()
torch.cuda.synchronize()
if rank == 0:
    print("DDP_CHECK_OK", flush=True)
# no destroy_process_group(): captured CUDA graphs still reference the communicator (see bench.py)
sys.stdout.flush()
os._exit(0)
