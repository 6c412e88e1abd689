"""GPU: the fused AIDE training step (aide_b200.trainer.AideTrainer) against the oracle's restatement of
train_files/trainchaos_proposed_30cases1labeled.py:263-325, eager launches vs the captured CUDA graph."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def batch(oracle, B, S, seed):
    (x1, x2), t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=seed, n_aug=4)
    return (x1, x2), t1, t2, augs


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_step_vs_oracle(oracle, graph):
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 64
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    st1, st2 = {}, {}
    d = lambda t: t.to(dev)
    for step in (1, 2):
        x, t1, t2, augs = batch(oracle, B, S, 500 + step)
        r = oracle.aide_step(oracle.fuseunet_forward, p1, p2, x, augs, t1, t2, 0.25)
        m = tr.step(tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 0.25)
        tol = 2e-5 if step == 1 else 2e-3          # step 2 starts from Adam-updated weights (sign-like first update)
        assert abs(m["loss1"].item() - r["loss1"].item()) < tol * max(1, abs(r["loss1"].item())), step
        assert abs(m["loss2"].item() - r["loss2"].item()) < tol * max(1, abs(r["loss2"].item())), step
        if step == 1:
            assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
            assert abs(m["dice1"].item() - r["dice1"].item()) < 4e-3
        oracle.adam_amsgrad_step(p1, r["grads1"], st1, step)
        oracle.adam_amsgrad_step(p2, r["grads2"], st2, step)
    # parameters after two Adam-amsgrad updates: the first update is lr * sign(g), so elements whose gradient is at
    # rounding level may differ by 2 * lr; everything else follows the oracle
    sd = tr.net1.state_dict()
    for k in ("last_conv1.weight", "up_block4.block.bn2.weight", "modal1_downblock3.block.conv1.weight"):
        diff = (sd[k].cpu() - p1[k].detach()).abs()
        assert diff.max().item() <= 4.1e-4 and diff.median().item() < 2e-5, (k, diff.max().item(), diff.median().item())
    assert int(sd["modal1_downblock1.block.bn1.num_batches_tracked"]) == 10          # 5 train-mode forwards per step
    assert tr.opt1.step_dev.item() == 2 and tr.steps == 2


def test_graph_replay_equals_eager_launches(oracle):
    """Same seeds, same batches: the captured graph and the eager launch sequence are the same kernels in the same
    order on deterministic reductions, so the trained weights must be bit-identical."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 3, 32
    tr_e = AideTrainer("fuseunet", mode="parity", device=dev, seed=5, cuda_graph=False)
    tr_g = AideTrainer("fuseunet", mode="parity", device=dev, seed=5, cuda_graph=True)
    assert torch.equal(tr_e.opt1.flat, tr_g.opt1.flat)
    d = lambda t: t.to(dev)
    for step in range(3):
        x, t1, t2, augs = batch(oracle, B, S, 700 + step)
        args = (tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 0.5)
        me = tr_e.step(*args)
        mg = tr_g.step(*args)
        assert torch.equal(me["loss1"], mg["loss1"]) and torch.equal(me["loss2"], mg["loss2"]), step
    for a, b in ((tr_e.opt1, tr_g.opt1), (tr_e.opt2, tr_g.opt2)):
        assert torch.equal(a.flat, b.flat) and torch.equal(a.vmax, b.vmax)
    for (ka, va), (kb, vb) in zip(tr_e.net2.state_dict().items(), tr_g.net2.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    # host buffers go straight into the graph's static inputs
    x, t1, t2, augs = batch(oracle, B, S, 800)
    host = dict(x=tuple(t.pin_memory() for t in x), t1=t1.pin_memory(), t2=t2.pin_memory(),
                augs=[tuple(t.pin_memory() for t in a) for a in augs])
    vals, h2d, d2h = tr_g.step_from_host(host, 0.5)
    ve = tr_e.step(tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 0.5)
    assert vals["loss1"] == ve["loss1"].item() and h2d == sum(t.numel() * t.element_size() for t in
                                                               list(x) + [t1, t2] + [u for a in augs for u in a])


def test_kidney_flavour_and_unet(oracle):
    """Single-modal UNet pair, eval-mode augmented forwards, sharpen = pow(1/T) (trainkidney_proposed_mask1.py:267-333)."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 32
    tr = AideTrainer("unet", mode="parity", device=dev, seed=2, flavour="kidney", temperature=2.0, cuda_graph=False)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_unet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_unet(2), requires_grad=True)
    (x1, _), t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=42, n_aug=4)
    r = oracle.aide_step(oracle.unet_forward, p1, p2, (x1,), [(a[0],) for a in augs], t1, t2, 0.25,
                         temperature=2.0, flavour="kidney")
    d = lambda t: t.to(dev)
    m = tr.step(d(x1), d(t1), d(t2), [d(a[0]) for a in augs], 0.25)
    assert abs(m["loss1"].item() - r["loss1"].item()) < 5e-5 and abs(m["loss2"].item() - r["loss2"].item()) < 5e-5
    assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
    assert int(tr.net1.state_dict()["down_block1.block.bn1.num_batches_tracked"]) == 1     # eval-mode aug forwards


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_equals_mean_of_shard_gradients(tmp_path):
    """DDP semantics on 2 GPUs -- the peer-memory all-reduce kernel bit for bit against the rank-order sum,
    replicas identical after eager / graph steps with both collectives, the update equal to the mean of the shard
    gradients, DataParallel's global selection against the oracle: a subprocess under torchrun, see tools/ddp_check.py."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(root, "tools", "ddp_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DDP_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_step_with_reverse_augmentation(oracle, graph):
    """The step with flipped / rotated views: the engine un-augments the views' logits on the GPU (aide_reverse_aug),
    the oracle through PIL as the reference does (trainchaos_proposed_30cases1labeled.py:271-272 -> :81-95)."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 64
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    x, t1, t2, augs = batch(oracle, B, S, 900)
    d = lambda t: t.to(dev)
    saved = [t.clone() for t in tr._state_tensors()]
    for rep, degs in enumerate(([17.5, -42.0, 0.0, 60.0], [-5.25, 33.0, 90.0, -60.0])):
        augset = {"augno": [4, 4, 3, 4]}
        for v in range(1, 5):
            augset[f"degree{v}"] = [g * (1 if v % 2 else -1) + v for g in degs]
            augset[f"hflip{v}"] = [(b + v + rep) % 2 for b in range(B)]
        r = oracle.aide_step(oracle.fuseunet_forward, oracle.clone_params(p1, True), oracle.clone_params(p2, True), x, augs,
                             t1, t2, 0.25, augset=augset)
        # same trainer, state rewound: the second repetition REPLAYS the captured graph with other augmentation
        # parameters (they live in static device buffers the captured kernels read)
        with torch.no_grad():
            for t, s_ in zip(tr._state_tensors(), saved):
                t.copy_(s_)
        m = tr.step(tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 0.25, augset=augset)
        assert abs(m["loss1"].item() - r["loss1"].item()) < 2e-5 * max(1, abs(r["loss1"].item())), rep
        assert abs(m["loss2"].item() - r["loss2"].item()) < 2e-5 * max(1, abs(r["loss2"].item())), rep
        assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"]), rep


@pytest.mark.parametrize("kind,flavour", [("unetsa", "kidney"), ("fuseunetsa", "chaos")])
def test_trainer_step_attention_variants(oracle, kind, flavour):
    """The AIDE step with the attention networks a script can select (trainkidney_proposed_mask1.py:73-80: UNetsa)."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 64
    tr = AideTrainer(kind, mode="parity", device=dev, seed=2, flavour=flavour, cuda_graph=True)
    torch.manual_seed(2)
    unet = kind == "unetsa"
    init = (lambda: oracle.init_unetsa(2)) if unet else (lambda: oracle.init_fuseunetsa(2))
    fwd = oracle.unetsa_forward if unet else oracle.fuseunetsa_forward
    p1 = oracle.clone_params(init(), requires_grad=True)
    p2 = oracle.clone_params(init(), requires_grad=True)
    (x1, x2), t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=321, n_aug=4)
    ins = (lambda a: (a[0],)) if unet else (lambda a: a)
    r = oracle.aide_step(fwd, p1, p2, ins((x1, x2)), [ins(a) for a in augs], t1, t2, 0.25, flavour=flavour)
    d = lambda t: t.to(dev)
    xin = d(x1) if unet else (d(x1), d(x2))
    ain = [d(a[0]) for a in augs] if unet else [tuple(d(t) for t in a) for a in augs]
    m = tr.step(xin, d(t1), d(t2), ain, 0.25)
    assert abs(m["loss1"].item() - r["loss1"].item()) < 5e-5 and abs(m["loss2"].item() - r["loss2"].item()) < 5e-5
    assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
    oracle.adam_amsgrad_step(p1, r["grads1"], {}, 1)
    sd = tr.net1.state_dict()
    key = "sa3.conv2.weight" if unet else "modal2_sa3.conv2.weight"
    diff = (sd[key].cpu() - p1[key].detach()).abs()
    assert diff.max().item() <= 2.1e-4 and diff.median().item() < 2e-5, (diff.max().item(), diff.median().item())


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_step_learned_bilinear_decoder(oracle, graph):
    """learned_bilinear=True (ConvTranspose2d(k=2,s=2) up path, netblocks.py:11-14) inside the fused step: the transposed
    weights keep their [Cin,Cout,2,2] slots in the flat parameter / gradient buffers (the 3x3 stand-in's gradient is
    re-indexed into them), so flat Adam, the bucket plan and the all-reduce need no special case.  One step against the
    oracle: losses, index sets, the transposed-weight gradient, the weights after Adam."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 64
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph, net_kwargs=dict(learned_bilinear=True))
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_fuseunet(2, learned_bilinear=True), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_fuseunet(2, learned_bilinear=True), requires_grad=True)
    assert list(tr.net1.state_dict().keys()) == list(p1.keys())
    d = lambda t: t.to(dev)
    x, t1, t2, augs = batch(oracle, B, S, 640)
    r = oracle.aide_step(oracle.fuseunet_forward, p1, p2, x, augs, t1, t2, 0.25)
    m = tr.step(tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 0.25)
    assert abs(m["loss1"].item() - r["loss1"].item()) < 2e-5 * max(1, abs(r["loss1"].item()))
    assert abs(m["loss2"].item() - r["loss2"].item()) < 2e-5 * max(1, abs(r["loss2"].item()))
    assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
    name = "up_block1.bilinear_up.0.weight"
    g_ref = r["grads1"][name]
    g = tr.net1._glayout.view(tr.net1.last_grad_flat, name).cpu()
    assert g.shape == g_ref.shape == (1024, 512, 2, 2)
    # the deepest decoder level of a 64x64 input is a 4x4 map under train-mode BatchNorm: its gradient is the most
    # rounding-sensitive of the net (cf. oracle_sensitivity_band); a wrong re-indexing would not correlate at all
    assert ((g - g_ref).abs().max() / g_ref.abs().max()).item() < 1e-2
    cos = torch.nn.functional.cosine_similarity(g.double().flatten(), g_ref.double().flatten(), dim=0).item()
    assert 1.0 - cos < 1e-4, cos
    st1 = {}
    oracle.adam_amsgrad_step(p1, r["grads1"], st1, 1)
    w = tr.net1.state_dict()[name].cpu()
    diff = (w - p1[name].detach()).abs()
    assert diff.max().item() <= 2.1e-4 and diff.median().item() < 2e-5, (diff.max().item(), diff.median().item())


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_checkpoint_resume_is_bit_exact(oracle, graph, tmp_path):
    """Checkpoint / resume: two steps, state_dict() through torch.save / torch.load, a third step -- a fresh trainer that
    loads the checkpoint and runs the same third step must land on the same bits (weights, BatchNorm buffers, Adam state,
    losses).  The optimiser part has the entries of torch.optim.Adam(amsgrad=True).state_dict(); the per-network files of
    the reference (trainchaos_proposed_30cases1labeled.py:504-526) load into the modules."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 64
    d = lambda t: t.to(dev)
    dev_batch = lambda b: (tuple(d(t) for t in b[0]), d(b[1]), d(b[2]), [tuple(d(t) for t in a) for a in b[3]])
    batches = [dev_batch(batch(oracle, B, S, 900 + i)) for i in range(3)]
    a = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph)
    for i in range(2):
        a.step(*batches[i], 0.25)
    path = str(tmp_path / "trainer.pt")
    torch.save(a.state_dict(), path)
    a.save_reference_checkpoints(str(tmp_path / "n1.pkl"), str(tmp_path / "n2.pkl"), epoch=3, loss1=0.5, loss2=0.6)
    ma = a.step(*batches[2], 0.25)
    la = (ma["loss1"].item(), ma["loss2"].item())
    b = AideTrainer("fuseunet", mode="parity", device=dev, seed=7, cuda_graph=graph)      # different initial weights
    b.load_state_dict(torch.load(path, map_location=dev))
    mb = b.step(*batches[2], 0.25)
    assert (mb["loss1"].item(), mb["loss2"].item()) == la
    for (ka, va), (kb, vb) in zip(a.net1.state_dict().items(), b.net1.state_dict().items()):
        assert torch.equal(va, vb), ka
    for x, y in ((a.opt2.m, b.opt2.m), (a.opt2.v, b.opt2.v), (a.opt2.vmax, b.opt2.vmax), (a.opt2.flat, b.opt2.flat)):
        assert torch.equal(x, y)
    assert int(b.opt1.step_dev.item()) == 3 and b.steps == 3
    # the optimiser part has torch.optim.Adam's per-parameter entries (step, exp_avg, exp_avg_sq, max_exp_avg_sq), numbered
    # in the flat buffer's parameter order
    import aide_b200 as A
    ref_net = A.fuseunet(num_classes=2, mode="parity").to(dev)
    ck = torch.load(path, map_location=dev)
    order = sorted(a.net1._glayout.off.items(), key=lambda kv: kv[1][0])
    assert len(ck["opt1"]["state"]) == len(order) == len(list(ref_net.parameters()))
    assert set(ck["opt1"]["state"][0]) == {"step", "exp_avg", "exp_avg_sq", "max_exp_avg_sq"}
    named = dict(ref_net.named_parameters())
    for i, (name, _) in enumerate(order):
        assert ck["opt1"]["state"][i]["exp_avg"].shape == named[name].shape
    # the reference-format file loads into a module and reproduces net1 after two steps
    ref_net.load_state_dict(torch.load(str(tmp_path / "n1.pkl"), map_location=dev)["net"])
    assert torch.equal(ref_net.state_dict()["last_conv1.weight"], ck["net1"]["last_conv1.weight"])


@pytest.mark.parametrize("graph", [False, True])
def test_step_from_host_prefetch_gives_the_same_steps(oracle, graph):
    """step_from_host(batch, prefetch=next) starts the next batch's H2D copy on a copy stream while the step computes; the
    results must be those of the plain call sequence, also when the following call does NOT use the prefetched batch."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 64
    hosts = []
    for i in range(4):
        x, t1, t2, augs = batch(oracle, B, S, 950 + i)
        hosts.append(dict(x=tuple(t.pin_memory() for t in x), t1=t1.pin_memory(), t2=t2.pin_memory(),
                          augs=[tuple(t.pin_memory() for t in a) for a in augs]))
    a = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph)
    b = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=graph)
    order = [0, 1, 2, 0, 3]
    nxt = [1, 2, 3, 3, None]                      # call 2 prefetches batch 3 but call 3 then runs batch 0: staged copy unused
    for i, k in enumerate(order):
        va, _, _ = a.step_from_host(hosts[k], 0.25)
        vb, h2d, d2h = b.step_from_host(hosts[k], 0.25, prefetch=None if nxt[i] is None else hosts[nxt[i]])
        assert va == vb, (i, va, vb)
    assert torch.equal(a.opt1.flat, b.opt1.flat) and torch.equal(a.opt2.flat, b.opt2.flat)
