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


for graph in (False, True):
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph)
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
            ref = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=False, process_group=None)
            ref.world = 1
            xs, a, b, au = shard(r)
            ref._step_eager(xs, a, b, au, 0.25)
            grads.append(ref.net1.last_grad_flat.clone())
        gmean = (grads[0] + grads[1]) / 2
        upd = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=False)
        upd.world = 1
        upd.opt1.step(gmean)
        diff = (upd.opt1.flat - mine).abs().max().item()
        assert diff <= 2.1e-4, diff        # identical up to the all-reduce summation order on sign-like first updates
        frac_same = (upd.opt1.flat == mine).float().mean().item()
        assert frac_same > 0.99, frac_same
        assert not torch.equal(mine, p0)
    dist.barrier()
    print(f"[rank {rank}] graph={graph} ok", flush=True)
torch.cuda.synchronize()
if rank == 0:
    print("DDP_CHECK_OK", flush=True)
# no destroy_process_group(): captured CUDA graphs still reference the communicator (see bench.py)
sys.stdout.flush()
os._exit(0)
