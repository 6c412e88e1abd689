"""GPU: parity at the sizes BASELINE.json's configs name (driver-run, `-m gpu`).

    C2  fuseunet forward + backward, batch 4, 256x256           (configs[1])
    C3  AIDE proposed step, two fuseunets, batch 8, 256x256     (configs[2]) through AideTrainer (captured CUDA graph)
    C5  two single-modal UNets, 320x320, kidney flavour         (configs[4] shape; per-GPU part)
plus the trainer behaviours a training script relies on between steps (schedulers, the warm-up rate, evaluation with
`graphed_eval` after graph-replayed steps).

The CPU side is the oracle (oracle/aide_oracle.py, pinned bit-for-bit to the reference by tests/golden/make_golden.py)
run live on the box's host cores with the same seeded inputs; where the golden file holds reference values for the
case (config 2) they are checked too.  Bars as in SURVEY.md 8c: logits max|d|/max|ref| <= 1e-3 (asserted at 2e-4),
per-image losses rel 2e-5, small-loss index sets equal, Dice_fn/B within 1e-4; gradients by the shortest-path tensor
and the direction of the whole gradient (per-tensor values are ill-conditioned in the fp32 reference itself, see
tests/test_gpu_network.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def relmax(a, b):
    return ((a.detach().cpu().double() - b.detach().cpu().double()).abs().max()
            / b.detach().cpu().double().abs().max().clamp_min(1e-30)).item()


def is_prebn_bias(name):
    return name.endswith(("block.conv1.bias", "block.conv2.bias", "bilinear_up.1.bias"))


def cosine_gap(ga, gb, names):
    fa = torch.cat([ga[k].detach().cpu().double().flatten() for k in names])
    fb = torch.cat([gb[k].detach().cpu().double().flatten() for k in names])
    return 1.0 - torch.nn.functional.cosine_similarity(fa, fb, dim=0).item()


@pytest.fixture()
def all_threads():
    n = torch.get_num_threads()
    torch.set_num_threads(os.cpu_count() or 1)
    yield
    torch.set_num_threads(n)


def test_config2_fuseunet_forward_backward_b4_256(golden, oracle, all_threads):
    """configs[1]: one fuseunet forward + backward at batch 4, 256x256 (CEMDiceLoss), engine vs reference values."""
    import aide_b200 as A
    dev = torch.device("cuda:0")
    g = golden["ka256"]
    gen = torch.Generator().manual_seed(1234)
    x1 = torch.randn(4, 3, 256, 256, generator=gen)
    x2 = torch.randn(4, 3, 256, 256, generator=gen)
    t = (torch.rand(4, 256, 256, generator=gen) < 0.08).long()
    torch.manual_seed(2)
    net = A.fuseunet(num_classes=2, mode="parity").to(dev).train()
    y = net(x1.to(dev), x2.to(dev))
    loss = A.CEMDiceLoss([1., 1.], [1., 1.], [1., 1.])(y, t.to(dev))
    loss.backward()
    li = A.CEMDiceLossImage([1., 1.], [1., 1.], [1., 1.])(y.detach(), t.to(dev))
    # frozen from the unmodified reference
    assert relmax(y[:, :, ::8, ::8], g["fuse_logits_sub"]) < 2e-4
    assert torch.allclose(li.cpu(), g["fuse_loss_img"], rtol=2e-5) and li.sort()[1].tolist() == g["fuse_sort"].tolist()
    assert abs(loss.item() - g["fuse_loss_mean"]) < 2e-5
    e_last = relmax(net.last_conv1.weight.grad, g["fuse_grad_last_w"])
    assert e_last < 1e-3, e_last
    assert relmax(net.last_conv1.bias.grad, g["fuse_grad_last_b"]) < 1e-4
    # live oracle on the same inputs: full logits, every gradient.  Per-tensor gradient values are ill-conditioned in
    # the fp32 reference itself (dW sums cancel heavily because train-mode BatchNorm makes every dz channel mean-free;
    # ReLU / max-pool masks flip at rounding level): the bar is the oracle's OWN one-ulp sensitivity band measured on
    # these inputs (as in tests/test_gpu_network.py), plus the well-conditioned shortest-path tensors above.
    from test_gpu_network import assert_grads_inside_band, oracle_sensitivity_band
    names = [k for k in oracle.init_fuseunet(2) if not oracle.is_buffer(k)]
    live = [k for k in names if not is_prebn_bias(k)]

    def grad_fn(xin):
        torch.manual_seed(2)
        pp = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
        yy = oracle.fuseunet_forward(pp, *xin, training=True)
        grad_fn.last = (pp, yy.detach())
        return dict(zip(live, torch.autograd.grad(oracle.ce_dice_mean(yy, t), [pp[k] for k in live])))

    go, band, band_cos = oracle_sensitivity_band(grad_fn, (x1, x2), live, n_draws=3)
    torch.manual_seed(2)
    p = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    yo = oracle.fuseunet_forward(p, x1, x2, training=True)
    assert relmax(go["last_conv1.weight"], g["fuse_grad_last_w"]) < 1e-5          # the live oracle IS the reference
    e_logit = relmax(y, yo)
    assert e_logit < 2e-4, e_logit
    eng = {k: v.grad for k, v in net.named_parameters()}
    for k in names:
        if is_prebn_bias(k):
            assert eng[k].abs().max().item() < 1e-5, k                            # analytically zero
    print(f"C2 fuseunet B=4 256x256: logits {e_logit:.2e}, last_conv1.weight grad {e_last:.2e}")
    assert_grads_inside_band({k: eng[k] for k in live}, go, band, band_cos, live, "C2 fuse/parity/256")
    # BatchNorm buffers after ONE train-mode forward
    sd = net.state_dict()
    for k, v in p.items():
        if k.endswith(("running_mean", "running_var")):
            assert relmax(sd[k], v) < 1e-4, k


def test_config3_aide_step_b8_256_through_trainer(oracle, all_threads):
    """configs[2]: the full AIDE proposed step at batch 8, 256x256 -- the benchmark's workload -- through
    AideTrainer's captured CUDA graph, against the oracle's restatement of
    train_files/trainchaos_proposed_30cases1labeled.py:263-325."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 8, 256
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=True)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    x, t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=1234, n_aug=4)
    r = oracle.aide_step(oracle.fuseunet_forward, p1, p2, x, augs, t1, t2, 0.25)
    d = lambda t: t.to(dev)
    m = tr.step(tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 0.25)
    e1, e2 = relmax(m["out1"], r["out1"]), relmax(m["out2"], r["out2"])
    dd = max(abs(m["dice1"].item() - r["dice1"].item()), abs(m["dice2"].item() - r["dice2"].item())) / B
    print(f"C3 AIDE step B=8 256x256: logits {e1:.2e} / {e2:.2e}, |dDice_fn/B| {dd:.2e}, "
          f"loss1 {m['loss1'].item():.6f} vs {r['loss1'].item():.6f}, loss2 {m['loss2'].item():.6f} vs {r['loss2'].item():.6f}")
    assert e1 < 2e-4 and e2 < 2e-4
    assert torch.allclose(m["pre1"].cpu(), r["pre1"].detach(), rtol=2e-5)
    assert torch.allclose(m["pre2"].cpu(), r["pre2"].detach(), rtol=2e-5)
    assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
    assert abs(m["loss1"].item() - r["loss1"].item()) < 2e-5 * max(1.0, abs(r["loss1"].item()))
    assert abs(m["loss2"].item() - r["loss2"].item()) < 2e-5 * max(1.0, abs(r["loss2"].item()))
    assert dd < 1e-4, dd
    from aide_b200 import lib
    assert lib.aide_f16_saturated(1) == 0          # no operand left the fp16 range of the split-precision planes
    # the update that followed: Adam-amsgrad's first step is lr * sign(g) -> elements at rounding level differ by 2 lr
    oracle.adam_amsgrad_step(p1, r["grads1"], {}, 1)
    sd = tr.net1.state_dict()
    for k in ("last_conv1.weight", "up_block4.block.bn2.weight", "up_block1.block.conv1.weight"):
        diff = (sd[k].cpu() - p1[k].detach()).abs()
        assert diff.max().item() <= 2.1e-4 and diff.median().item() < 2e-5, (k, diff.max().item())
    assert int(sd["modal1_downblock1.block.bn1.num_batches_tracked"]) == 5


def test_config5_unet_320_kidney_step(oracle, all_threads):
    """configs[4] shape on one GPU: two single-modal UNets at 320x320, eval-mode pseudo-label forwards, sharpen
    pow(1/T) (trainkidney_proposed_mask1.py:267-333), batch 4."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 4, 320
    tr = AideTrainer("unet", mode="parity", device=dev, seed=2, flavour="kidney", temperature=0.5, cuda_graph=True)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_unet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_unet(2), requires_grad=True)
    (x1, _), t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=77, n_aug=4)
    r = oracle.aide_step(oracle.unet_forward, p1, p2, (x1,), [(a[0],) for a in augs], t1, t2, 0.25,
                         temperature=0.5, flavour="kidney")
    d = lambda t: t.to(dev)
    m = tr.step(d(x1), d(t1), d(t2), [d(a[0]) for a in augs], 0.25)
    e1, e2 = relmax(m["out1"], r["out1"]), relmax(m["out2"], r["out2"])
    dd = max(abs(m["dice1"].item() - r["dice1"].item()), abs(m["dice2"].item() - r["dice2"].item())) / B
    print(f"C5 UNet pair B=4 320x320 (kidney flavour): logits {e1:.2e} / {e2:.2e}, |dDice_fn/B| {dd:.2e}")
    assert e1 < 2e-4 and e2 < 2e-4
    assert torch.allclose(m["pre1"].cpu(), r["pre1"].detach(), rtol=2e-5)
    assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"])
    assert abs(m["loss1"].item() - r["loss1"].item()) < 5e-5 and abs(m["loss2"].item() - r["loss2"].item()) < 5e-5
    assert dd < 1e-4 * (256.0 / S) ** 2 * 4, dd       # thresholded count: one tie pixel moves Dice_fn/B by ~2e-6 * (256/S)^2
    assert int(tr.net1.state_dict()["down_block1.block.bn1.num_batches_tracked"]) == 1


def test_graphed_eval_sees_weights_updated_by_graph_replay(oracle):
    """After graph-replayed training steps (Adam runs inside the replay, behind autograd's back) an evaluation
    forward -- eager or net.graphed_eval() -- must use the UPDATED convolution weights (the per-epoch evaluation /
    pseudo-label rewrite loop, trainchaos_proposed_30cases1labeled.py:373-496)."""
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 3, 64
    tr = AideTrainer("fuseunet", mode="parity", device=dev, seed=2, cuda_graph=True, lr=1e-2)
    d = lambda t: t.to(dev)
    xe = tuple(torch.randn(1, 3, S, S, generator=torch.Generator().manual_seed(3)).to(dev) for _ in range(2))
    tr.net1.eval()
    f = tr.net1.graphed_eval(*xe)
    with torch.no_grad():
        y0 = f(*xe).clone()
    tr.net1.train()
    for step in range(3):                      # step 0 captures, steps 1-2 replay at a fixed rate
        x, t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=600 + step, n_aug=4)
        tr.step(tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], 1.0)
    tr.net1.eval()
    with torch.no_grad():
        y_graph = f(*xe).clone()
        # an independent module holding the trained state: fresh weight planes by construction
        import aide_b200 as A
        ref = A.fuseunet(num_classes=2, mode="parity").to(dev).eval()
        ref.load_state_dict(tr.net1.state_dict())
        y_ref = ref(*xe)
        y_eager = tr.net1(*xe)
    assert relmax(y_graph, y0) > 1e-3                  # lr 1e-2 for three steps: the weights did move
    assert torch.equal(y_eager, y_ref)
    assert torch.equal(y_graph, y_ref)
    tr.net1.train()


def test_rate_and_lr_change_under_one_captured_graph(oracle):
    """The warm-up rate (trainchaos_proposed_30cases1labeled.py:248) and the scheduler's learning rate (:236-241,
    :589-590) are device scalars of the captured step: changing them must neither re-capture nor be ignored.
    Bit-equality against the eager trainer fed the same sequence."""
    from torch.optim.lr_scheduler import StepLR
    from aide_b200.trainer import AideTrainer
    dev = torch.device("cuda:0")
    B, S = 3, 32
    tr_e = AideTrainer("fuseunet", mode="parity", device=dev, seed=5, cuda_graph=False)
    tr_g = AideTrainer("fuseunet", mode="parity", device=dev, seed=5, cuda_graph=True)
    scheds = [StepLR(o, step_size=1, gamma=0.5) for o in (tr_e.opt1, tr_e.opt2, tr_g.opt1, tr_g.opt2)]
    d = lambda t: t.to(dev)
    for step, rate in enumerate((0.0, 0.25, 0.25, 1.0)):
        x, t1, t2, augs = oracle.synthetic_batch(B, S, S, seed=300 + step, n_aug=4)
        args = (tuple(d(t) for t in x), d(t1), d(t2), [tuple(d(t) for t in a) for a in augs], rate)
        me, mg = tr_e.step(*args), tr_g.step(*args)
        assert torch.equal(me["loss1"], mg["loss1"]) and torch.equal(me["loss2"], mg["loss2"]), step
        for s_ in scheds:
            s_.step()
    assert len(tr_g._graphs) == 1
    assert abs(tr_g.opt1.lr - 1e-4 * 0.5 ** 4) < 1e-12 and tr_g.opt1.lr_dev.item() == pytest.approx(1e-4 * 0.5 ** 3)
    assert torch.equal(tr_e.opt1.flat, tr_g.opt1.flat) and torch.equal(tr_e.opt2.flat, tr_g.opt2.flat)
    # a learning rate of zero freezes the weights: proof that the kernel reads the device scalar
    before = tr_g.opt1.flat.clone()
    tr_g.set_lr(0.0)
    tr_g.step(*args)
    assert torch.equal(before, tr_g.opt1.flat)
