"""CPU, world_size 2, gloo: the host-side data-parallel logic of the AIDE step.

The engine's kernels need a GPU, so the per-rank gradients here come from the oracle; what is under test is the
host plumbing the trainer uses around them: the flat fp32 gradient layout (engine.GradLayout -- the unit of the
all-reduce and of the fused Adam), one all-reduce per net with the 1/world scale, parameter broadcast from rank 0,
rank-local BatchNorm statistics and rank-local small-loss selection.  Identity checked: the 2-rank result equals
the mean of the two shards' gradients computed in one process."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard(oracle, rank, B=3, S=32):
    return oracle.synthetic_batch(B, S, S, seed=300 + rank, n_aug=1)


def _flat_grads(glayout, grads):
    flat = torch.zeros(glayout.total, dtype=torch.float32)
    for name, g in grads.items():
        glayout.view(flat, name).copy_(g)
    return flat


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from aide_b200 import engine as E
    from oracle import aide_oracle as O
    plan = E.plan_fuseunet(2)
    glayout = E.GradLayout(plan)
    # rank 0 owns the initial weights; the others start from garbage and must receive them (broadcast_parameters)
    torch.manual_seed(2 if rank == 0 else 77)
    p1 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
    p2 = O.clone_params(O.init_fuseunet(2), requires_grad=True)
    for p in (p1, p2):
        for k, v in p.items():
            dist.broadcast(v.data, 0)
    (x1, x2), t1, t2, augs = _shard(O, rank)
    r = O.aide_step(O.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, 0.25, n_clean=1)
    flats = []
    for grads in (r["grads1"], r["grads2"]):          # one all-reduce per net over the flat buffer, then 1/world
        flat = _flat_grads(glayout, grads)
        dist.all_reduce(flat)
        flat.mul_(1.0 / world)
        flats.append(flat)
    torch.save(dict(flat1=flats[0], flat2=flats[1], indx1=r["indx1"], indx2=r["indx2"],
                    rm=p1["modal1_downblock1.block.bn1.running_mean"].clone(),
                    w0=p1["modal1_downblock1.block.conv1.weight"].detach().clone()),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_step_equals_mean_of_shard_gradients(tmp_path, oracle):
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / f"rank{r}.pt") for r in (0, 1))
    # replicas hold identical weights after the broadcast and identical averaged gradients after the all-reduce
    assert torch.equal(r0["w0"], r1["w0"])
    assert torch.equal(r0["flat1"], r1["flat1"]) and torch.equal(r0["flat2"], r1["flat2"])
    # BatchNorm running statistics and the small-loss ordering are rank-local (different shards -> different values)
    assert not torch.equal(r0["rm"], r1["rm"])
    # single-process reference: same initial weights, each shard on its own copy, gradients averaged
    sys.path.insert(0, ROOT)
    from aide_b200 import engine as E
    glayout = E.GradLayout(E.plan_fuseunet(2))
    flats = []
    for rank in (0, 1):
        torch.manual_seed(2)
        p1 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
        p2 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
        (x1, x2), t1, t2, augs = _shard(oracle, rank)
        r = oracle.aide_step(oracle.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, 0.25, n_clean=1)
        flats.append((_flat_grads(glayout, r["grads1"]), _flat_grads(glayout, r["grads2"])))
        assert torch.equal(r["indx1"], (r0, r1)[rank]["indx1"]) and torch.equal(r["indx2"], (r0, r1)[rank]["indx2"])
    for k, got in ((0, r0["flat1"]), (1, r0["flat2"])):
        want = (flats[0][k] + flats[1][k]) / 2
        scale = want.abs().max().item()
        assert (got - want).abs().max().item() <= 1e-6 * scale + 1e-9      # thread-count-dependent fp32 summation order
    # the flat layout covers every parameter exactly once (4-element alignment gaps stay zero)
    n_params = sum(v.numel() for k, v in oracle.init_fuseunet(2).items() if not oracle.is_buffer(k))
    assert n_params == 26675074 and glayout.total >= n_params
    assert int((r0["flat1"] != 0).sum()) <= n_params
