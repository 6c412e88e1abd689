"""GPU: whole-network parity of the CUDA engine against the CPU oracle and the golden vectors.

Bars (BASELINE.json north_star / SURVEY.md 8c): logits max|d|/max|ref| <= 1e-3 -- asserted ~10x tighter here
(LOGIT_TOL) -- argmax masks equal except at numerical ties (pixels whose oracle margin is below the logit error;
counted and bounded), small-loss index sets equal, per-image losses rel <= 2e-5 (parity mode).

Gradients: the fp32 REFERENCE ITSELF is ill-conditioned here.  ReLU / max-pool masks flip for pre-activations within
rounding error of zero, and one flip moves a whole term of a dW sum: perturbing the oracle's inputs by 1e-7
relative (one ulp) moves 93 of 97 gradient tensors by more than 1e-3 of their max and single tensors by up to 25 %,
while the logits move by 1e-5 (DESIGN.md "Gradient conditioning"; tools/grad_probe.py).  A per-tensor 1e-3 bar is
therefore unattainable for ANY implementation that is not bit-identical to oneDNN's summation order, the oracle with
8 threads vs 1 thread included.  The gradient tests measure the oracle's own one-ulp sensitivity band on the same
inputs and require the engine to stay inside it (max, median and whole-gradient cosine); every backward KERNEL is
separately held to <= 3e-5 against torch fp32 in test_gpu_kernels.py.  The conv biases that precede a train-mode
BatchNorm have a true gradient of zero (third trap): absolute tolerance."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOGIT_TOL = {"exact": 1e-4, "parity": 2e-4, "parity_tf32": 2e-4, "fast": 2.5e-1}    # measured: exact 2e-5, parity 5e-5..6e-5


def relmax(a, b):
    return ((a.detach().cpu().double() - b.detach().cpu().double()).abs().max()
            / b.detach().cpu().double().abs().max().clamp_min(1e-30)).item()


def is_prebn_bias(name):
    return name.endswith(("block.conv1.bias", "block.conv2.bias", "bilinear_up.1.bias"))


def grad_deviation(ga, gb, names):
    """per-tensor max|a-b|/max|b|, and 1 - cosine of the concatenated gradients."""
    per = {k: relmax(ga[k], gb[k]) for k in names}
    fa = torch.cat([ga[k].detach().cpu().double().flatten() for k in names])
    fb = torch.cat([gb[k].detach().cpu().double().flatten() for k in names])
    cos = torch.nn.functional.cosine_similarity(fa, fb, dim=0).item()
    return per, 1.0 - cos


def oracle_sensitivity_band(grad_fn, xs, names, n_draws=4, eps=1e-7):
    """The oracle against itself when its inputs are perturbed by `eps` relative (one fp32 ulp)."""
    g0 = grad_fn(xs)
    worst = {k: 0.0 for k in names}
    worst_cos = 0.0
    for i in range(n_draws):
        gen = torch.Generator().manual_seed(900 + i)
        xp = tuple(x * (1 + eps * torch.randn(x.shape, generator=gen)) for x in xs)
        per, c = grad_deviation(grad_fn(xp), g0, names)
        worst = {k: max(worst[k], per[k]) for k in names}
        worst_cos = max(worst_cos, c)
    return g0, worst, worst_cos


def assert_grads_inside_band(engine_grads, g0, band, band_cos, names, what):
    per, c = grad_deviation(engine_grads, g0, names)
    bmax, bmed = max(band.values()), float(np.median(list(band.values())))
    emax, emed = max(per.values()), float(np.median(list(per.values())))
    print(f"{what}: engine-vs-oracle grads max {emax:.2e} median {emed:.2e} 1-cos {c:.2e} | "
          f"oracle one-ulp band max {bmax:.2e} median {bmed:.2e} 1-cos {band_cos:.2e}")
    assert emax < max(1e-3, 4.0 * bmax), (what, emax, bmax)      # single flips: heavy-tailed, 4 draws only
    assert emed < max(1e-3, 4.0 * bmed), (what, emed, bmed)
    assert c < max(1e-6, 8.0 * band_cos), (what, c, band_cos)


def build(kind, mode, dev):
    import aide_b200
    torch.manual_seed(2)
    net = (aide_b200.fuseunet if kind == "fuse" else aide_b200.UNet)(num_classes=2, mode=mode)
    return net.to(dev)


def oracle_params(oracle, kind):
    torch.manual_seed(2)
    return oracle.clone_params(oracle.init_fuseunet(2) if kind == "fuse" else oracle.init_unet(2), requires_grad=True)


def fwd_oracle(oracle, kind, p, xs, training=True):
    return oracle.fuseunet_forward(p, *xs, training=training) if kind == "fuse" else oracle.unet_forward(p, xs[0], training=training)


@pytest.mark.parametrize("mode", ["exact", "parity", "parity_tf32"])
@pytest.mark.parametrize("kind", ["fuse", "unet"])
@pytest.mark.parametrize("tag,shape", [("s32", (2, 32, 32)), ("s48x64", (3, 48, 64))])
def test_forward_backward_vs_golden_and_oracle(golden, oracle, mode, kind, tag, shape):
    import aide_b200
    dev = torch.device("cuda:0")
    b, h, w = shape
    (x1, x2), t1, t2, _ = oracle.synthetic_batch(b, h, w, seed=1234)
    xs = (x1, x2) if kind == "fuse" else (x1,)
    g = golden[tag][kind]
    net = build(kind, mode, dev)
    net.train()
    y = net(*[x.to(dev) for x in xs])
    assert y.shape == (b, 2, h, w) and y.requires_grad
    assert relmax(y, g["logits"]) < LOGIT_TOL[mode]
    # argmax: a pixel may differ from the fp32 reference only where it is a numerical tie, i.e. where the margin of the
    # SAME forward evaluated in fp64 lies inside the reference's own fp32-vs-fp64 noise band on that margin
    flips = y.argmax(1).cpu() != g["logits"].argmax(1)
    if mode != "fast":
        p64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in oracle_params(oracle, kind).items()}
        with torch.no_grad():
            y64 = fwd_oracle(oracle, kind, {k: v.detach() for k, v in p64.items()}, tuple(x.double() for x in xs))
        d64 = y64[:, 1] - y64[:, 0]
        band = ((g["logits"][:, 1] - g["logits"][:, 0]).double() - d64).abs().max()
        assert (flips & (d64.abs() > band)).sum() == 0 and int(flips.sum()) <= 2, (int(flips.sum()), band.item())
    # loss + gradients
    p = oracle_params(oracle, kind)
    yo = fwd_oracle(oracle, kind, p, xs)
    if kind == "fuse":
        loss = aide_b200.CEMDiceLoss([1., 1.], [1., 1.], [1., 1.])(y, t2.to(dev))
        loss_o = oracle.ce_dice_mean(yo, t2)
        assert abs(loss.item() - g["loss_mean"]) < 2e-5
        li = aide_b200.CEMDiceLossImage([1., 1.], [1., 1.], [1., 1.])(y.detach(), t2.to(dev))
        assert torch.allclose(li.cpu(), g["loss_img"], rtol=2e-5) and torch.equal(li.sort()[1].cpu(), g["sort_idx"])
        assert abs(aide_b200.Dice_fn(y, t2.to(dev)).item() - g["dice_fn"]) < 1e-4
    else:
        loss = aide_b200.DiceLoss()(y, t1.to(dev))
        loss_o = oracle.dice_loss_mean(yo, t1)
        assert abs(loss.item() - g["dice_loss"]) < 2e-5
    loss.backward()
    names = [k for k in p if not oracle.is_buffer(k) and not is_prebn_bias(k)]

    def grad_fn(xin):
        pp = oracle_params(oracle, kind)
        yy = fwd_oracle(oracle, kind, pp, xin)
        ll = oracle.ce_dice_mean(yy, t2) if kind == "fuse" else oracle.dice_loss_mean(yy, t1)
        return dict(zip(names, torch.autograd.grad(ll, [pp[k] for k in names])))

    go, band, band_cos = oracle_sensitivity_band(grad_fn, xs, names)
    assert relmax(go["last_conv1.weight"], g["grad_last_w"]) < 1e-6          # the live oracle IS the frozen reference
    eng = {}
    for name, prm in net.named_parameters():
        assert prm.grad is not None, name
        if is_prebn_bias(name):
            assert prm.grad.abs().max().item() < 1e-5, name      # analytically zero
            continue
        eng[name] = prm.grad
    assert_grads_inside_band(eng, go, band, band_cos, names, f"{kind}/{mode}/{tag}")
    assert relmax(eng["last_conv1.weight"], go["last_conv1.weight"]) < 2e-3   # shortest path, least sensitive
    # BatchNorm buffers (running stats, counters) follow the module semantics
    sd = net.state_dict()
    for k, v in p.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert relmax(sd[k], v) < 1e-4, k
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v) == 1
    # eval-mode forward uses the running statistics
    net.eval()
    with torch.no_grad():
        ye = net(*[x.to(dev) for x in xs])
    yeo = fwd_oracle(oracle, kind, {k: v.detach() for k, v in p.items()}, xs, training=False)
    assert relmax(ye, yeo) < 1e-3
    if kind == "fuse":
        assert relmax(ye, g["logits_eval"]) < 1e-3


def test_state_dict_roundtrip_with_reference_layout(oracle):
    """A reference checkpoint ({'net': state_dict}) loads into the engine module and reproduces the oracle."""
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    p = oracle.init_fuseunet(2)
    for k in p:                                   # non-trivial BN buffers / affine params
        if k.endswith(("running_mean", "bn1.bias", "bn2.bias")):
            p[k] = torch.randn_like(p[k]) * 0.1
        if k.endswith("running_var"):
            p[k] = torch.rand_like(p[k]) + 0.5
    net = build("fuse", "parity", dev)
    net.load_state_dict({k: v.clone() for k, v in p.items()})
    net.eval()
    (x1, x2), *_ = oracle.synthetic_batch(2, 32, 32, seed=3)
    with torch.no_grad():
        y = net(x1.to(dev), x2.to(dev))
    assert relmax(y, oracle.fuseunet_forward(p, x1, x2, training=False)) < 1e-4


def test_aide_step_vs_golden(golden, oracle):
    """Full AIDE step (4 augmented train-mode forwards per net, pseudo labels, co-teaching selection, backward)
    at 64x64, B=4 against the values frozen from the reference flow (trainchaos_proposed...:263-321)."""
    import aide_b200 as A
    dev = torch.device("cuda:0")
    g = golden["aide64"]
    (x1, x2), t1, t2, augs = oracle.synthetic_batch(4, 64, 64, seed=1234, n_aug=4)
    torch.manual_seed(2)
    n1 = A.fuseunet(num_classes=2, mode="parity").to(dev)
    n2 = A.fuseunet(num_classes=2, mode="parity").to(dev)
    n1.train(); n2.train()
    d = lambda t: t.to(dev)
    a1 = [n1(d(a), d(b)).detach() for a, b in augs]
    a2 = [n2(d(a), d(b)).detach() for a, b in augs]
    q1, w1 = A.pseudo_label(a1, 1.0)
    q2, w2 = A.pseudo_label(a2, 1.0)
    assert abs(q1.double().sum().item() - g["q1_sum"]) < 1e-2 and abs(w1.double().sum().item() - g["w1_sum"]) < 5e-2
    # |dq| <= |dlogit| / 4 and |dw| <= 4 |dq|; logits agree to LOGIT_TOL * max|logit| ~ 4e-4
    assert torch.allclose(q2.cpu(), g["q2"], atol=1e-4) and torch.allclose(w2.cpu(), g["w2"], atol=4e-4)
    o1, o2 = n1(d(x1), d(x2)), n2(d(x1), d(x2))
    assert relmax(o1, g["out1"]) < LOGIT_TOL["parity"] and relmax(o2, g["out2"]) < LOGIT_TOL["parity"]
    m = A.coteach_step(o1, o2, d(t1), d(t2), q1, w1, q2, w2, 0.25)
    assert torch.allclose(m["pre1"].cpu(), g["pre1"], rtol=2e-5) and torch.allclose(m["pre2"].cpu(), g["pre2"], rtol=2e-5)
    assert torch.equal(m["indx1"].cpu(), g["indx1"]) and torch.equal(m["indx2"].cpu(), g["indx2"])
    assert abs(m["loss1"].item() - g["loss1"]) < 2e-5 and abs(m["loss2"].item() - g["loss2"]) < 2e-5
    # Dice_fn is a thresholded count: one pixel at a numerical tie moves the batch sum by ~2/(sum p + sum t) ~ 1e-3 here
    assert abs(m["dice1"].item() - g["dice1"]) < 4e-3 and abs(m["dice2"].item() - g["dice2"]) < 4e-3
    m["loss1"].backward(retain_graph=True)
    m["loss2"].backward()
    assert relmax(n1.last_conv1.weight.grad, g["grad1_last_w"]) < 2e-3
    # gradient magnitudes per tensor (ill-conditioned element-wise, see the module docstring): median within 1 %
    devs = [abs(prm.grad.abs().max().item() - g["grad1_absmax"][name]) / g["grad1_absmax"][name]
            for name, prm in n1.named_parameters() if not is_prebn_bias(name)]
    assert float(np.median(devs)) < 1e-2 and max(devs) < 0.5, (float(np.median(devs)), max(devs))
    assert int(n1.modal1_downblock1.block.bn1.num_batches_tracked) == g["nbt"] == 5
    assert relmax(n1.up_block4.block.bn2.running_mean, g["rm_last"]) < 1e-4
    # the drop-in (unfused) formulation of the reference script gives the same loss through autograd indexing
    crit = A.CEMDiceLossImage([1., 1.], torch.tensor([1., 1.]), [1., 1.])
    mse = A.MulticlassMSELoss(reduction="none")
    i2 = m["indx2"]
    l1 = crit(o1[i2[0:2]], d(t2)[i2[0:2]]).mean() + 0.75 * crit(o1[i2[2:]], d(t2)[i2[2:]]).mean() \
        + 10.0 * 0.25 * (w2[i2[2:]] * mse(o1[i2[2:]], q2[i2[2:]])).mean()
    assert abs(l1.item() - g["loss1"]) < 2e-5


def test_known_answers_256_parity_mode(golden):
    """SURVEY.md 8c known answers at BASELINE config-2 size: fuseunet, seed 2, B=4, 256x256, train-mode BN."""
    import aide_b200 as A
    dev = torch.device("cuda:0")
    g = golden["ka256"]
    torch.manual_seed(2)
    f = A.fuseunet(num_classes=2, mode="parity").to(dev)
    u = A.UNet(num_classes=2, mode="parity").to(dev)
    gen = torch.Generator().manual_seed(1234)
    x1 = torch.randn(4, 3, 256, 256, generator=gen)
    x2 = torch.randn(4, 3, 256, 256, generator=gen)
    t = (torch.rand(4, 256, 256, generator=gen) < 0.08).long()
    with torch.no_grad():
        yf = f(x1.to(dev), x2.to(dev))
        yu = u(x1.to(dev))
    assert abs(yf.double().sum().item() - g["fuse_sum"]) < 2.0
    assert abs(yf.abs().max().item() - g["fuse_absmax"]) < 4e-4
    assert relmax(yf[:, :, ::8, ::8], g["fuse_logits_sub"]) < LOGIT_TOL["parity"]
    assert relmax(yu[:, :, ::8, ::8], g["unet_logits_sub"]) < LOGIT_TOL["parity"]
    packed = torch.from_numpy(np.packbits((yf.argmax(1) == 1).cpu().numpy().reshape(-1)))
    flipped = np.unpackbits((packed ^ g["fuse_argmax_packed"]).numpy()).nonzero()[0]
    # Argmax audit (262 144 pixels).  The golden file holds the SAME forward evaluated in fp64 by the oracle: the fp64
    # margin d = z1 - z0 of every pixel with |d| < 64 * band, where band = max |d_fp32 - d_fp64| is the reference's own
    # fp32 rounding noise on the margin (4.1e-5; the fp32 reference itself flips 1 pixel against fp64).  EVERY pixel
    # whose engine argmax differs from the fp32 reference's must be such a numerical tie: |d_fp64| inside the band.
    band = g["fp64_margin_band"]
    cand = {int(i): float(m) for i, m in zip(g["fp64_cand_idx"].tolist(), g["fp64_cand_margin"].tolist())}
    margins = [abs(cand.get(int(i), float("inf"))) for i in flipped]
    e64 = relmax(yf[:, :, ::8, ::8], g["fp64_logits_sub"])
    print(f"256x256 parity: {len(flipped)} argmax flips of 262144 vs the fp32 reference; their fp64 margins "
          f"{['%.2e' % m for m in margins]} (band {band:.2e}; the fp32 reference flips {g['fp32_flips_vs_fp64']} vs fp64); "
          f"engine logits vs fp64 {e64:.2e} (fp32 reference vs fp64 {g['fp64_logit_err_rel']:.2e})")
    assert all(m < band for m in margins), (margins, band)
    assert len(flipped) <= 8, len(flipped)
    li = A.CEMDiceLossImage([1., 1.], [1., 1.], [1., 1.])(yf, t.to(dev))
    assert torch.allclose(li.cpu(), g["fuse_loss_img"], rtol=2e-5)
    assert li.sort()[1].tolist() == [1, 2, 0, 3]
    assert abs(A.DiceLoss()(yu, t.to(dev)).item() - g["unet_dice_loss"]) < 1e-5
    assert abs(A.Dice_fn(yu, t.to(dev)).item() - g["unet_dice_fn"]) < 1e-4


def test_fast_mode_sanity(oracle):
    """Single-pass BF16 operands: not a parity mode (1.7e-1 of max|logit| at 256^2), only sanity."""
    dev = torch.device("cuda:0")
    (x1, x2), t1, t2, _ = oracle.synthetic_batch(2, 64, 64, seed=1234)
    net = build("fuse", "fast", dev).train()
    ref = build("fuse", "parity", dev).train()
    import aide_b200 as A
    y, yr = net(x1.to(dev), x2.to(dev)), ref(x1.to(dev), x2.to(dev))
    assert relmax(y, yr) < LOGIT_TOL["fast"]
    A.CEMDiceLoss()(y, t2.to(dev)).backward()
    A.CEMDiceLoss()(yr, t2.to(dev)).backward()
    cos = torch.nn.functional.cosine_similarity(net.last_conv1.weight.grad.flatten(), ref.last_conv1.weight.grad.flatten(), dim=0)
    assert cos.item() > 0.98
    gf, gr = net.up_block2.block.conv1.weight.grad.flatten(), ref.up_block2.block.conv1.weight.grad.flatten()
    assert torch.nn.functional.cosine_similarity(gf, gr, dim=0).item() > 0.8


def test_teacher_forced_training_steps(oracle):
    """Per-step parity with re-synchronised state (second finding): at every step both
    engines start from the ORACLE's weights / BN buffers, run the AIDE step, and must agree on Dice_fn/B and on the
    small-loss index sets; the oracle then advances with Adam-amsgrad.  The 1e-4 Dice bar of BASELINE.json is
    stated at 256x256, where one pixel at a numerical tie moves Dice_fn/B by ~2e-6; Dice_fn is a thresholded
    count, so at a smaller test size S the same pixel moves it (256/S)^2 times more and the bar scales with it
    (1.6e-3 at 64).  Default: 10 steps at the full 256x256 size, batch 4 (AIDE_TF_STEPS / AIDE_TF_SIZE change it; the
    100-step run is recorded in profiles/)."""
    import aide_b200 as A
    dev = torch.device("cuda:0")
    steps = int(os.environ.get("AIDE_TF_STEPS", "10"))
    size = int(os.environ.get("AIDE_TF_SIZE", "256"))
    B = int(os.environ.get("AIDE_TF_BATCH", "4"))
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    n1 = A.fuseunet(num_classes=2, mode="parity").to(dev).train()
    n2 = A.fuseunet(num_classes=2, mode="parity").to(dev).train()
    st1, st2 = {}, {}
    worst_dice, worst_logit = 0.0, 0.0
    for step in range(1, steps + 1):
        (x1, x2), t1, t2, augs = oracle.synthetic_batch(B, size, size, seed=1000 + step, n_aug=4)
        n1.load_state_dict({k: v.detach().clone() for k, v in p1.items()})
        n2.load_state_dict({k: v.detach().clone() for k, v in p2.items()})
        r = oracle.aide_step(oracle.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, 0.25)
        d = lambda t: t.to(dev)
        a1 = [n1(d(a), d(b)).detach() for a, b in augs]
        a2 = [n2(d(a), d(b)).detach() for a, b in augs]
        q1, w1 = A.pseudo_label(a1)
        q2, w2 = A.pseudo_label(a2)
        o1, o2 = n1(d(x1), d(x2)), n2(d(x1), d(x2))
        m = A.coteach_step(o1, o2, d(t1), d(t2), q1, w1, q2, w2, 0.25)
        worst_logit = max(worst_logit, relmax(o1, r["out1"]), relmax(o2, r["out2"]))
        assert relmax(o1, r["out1"]) < 1e-3 and relmax(o2, r["out2"]) < 1e-3, step
        assert torch.equal(m["indx1"].cpu(), r["indx1"]) and torch.equal(m["indx2"].cpu(), r["indx2"]), step
        dd = max(abs(m["dice1"].item() - r["dice1"].item()), abs(m["dice2"].item() - r["dice2"].item())) / B
        worst_dice = max(worst_dice, dd)
        assert dd < 1e-4 * (256.0 / size) ** 2, (step, dd)
        assert abs(m["loss1"].item() - r["loss1"].item()) < 1e-4 and abs(m["loss2"].item() - r["loss2"].item()) < 1e-4
        oracle.adam_amsgrad_step(p1, r["grads1"], st1, step)
        oracle.adam_amsgrad_step(p2, r["grads2"], st2, step)
    print(f"teacher-forced {steps} steps @ {size}: worst |dDice_fn/B| {worst_dice:.2e}, worst logit rel {worst_logit:.2e}")


@pytest.mark.parametrize("kind,training", [("fuse", True), ("unet", True), ("fuse", False)])
def test_grouped_augmented_forward_equals_sequential_forwards(oracle, kind, training):
    """The AIDE step's 4 augmented forwards per net (trainchaos_proposed_30cases1labeled.py:265-269) run as ONE
    stacked-batch forward with per-view BatchNorm statistics: logits, running statistics and num_batches_tracked must
    match 4 separate forward calls (different tile shapes -> not bit-identical, fp32 summation-order level)."""
    dev = torch.device("cuda:0")
    B, S, G = 3, 64, 4
    a, b = build(kind, "parity", dev), build(kind, "parity", dev)
    b.load_state_dict(a.state_dict())
    a.train(training)
    b.train(training)
    g = torch.Generator().manual_seed(11)
    n_in = 2 if kind == "fuse" else 1
    views = [tuple(torch.randn(B, 3, S, S, generator=g).to(dev) for _ in range(n_in)) for _ in range(G)]
    with torch.no_grad():
        seq = [a(*v) for v in views]
        grp = b._engine_forward_grouped(views)
    torch.cuda.synchronize()
    for s_, g_ in zip(seq, grp):
        assert relmax(g_, s_) < 2e-5
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if k.endswith("num_batches_tracked"):
            assert int(sa[k]) == int(sb[k]) == (G if training else 0)
        elif "running" in k:
            assert relmax(sb[k], sa[k]) < 1e-5, k


@pytest.mark.parametrize("mode", ["exact", "parity"])
@pytest.mark.parametrize("kind", ["fuse", "unet"])
def test_learned_bilinear_decoder_vs_reference_vectors(oracle, kind, mode):
    """learned_bilinear=True: ConvTranspose2d(k=2,s=2) -> BN -> ReLU up path (netblocks.py:11-14, UNet.py:6-9), run
    by the engine as conv3x3 on the zero-inserted input.  Forward vs vectors frozen from the unmodified reference,
    gradients vs the oracle (robust measures at this tiny size, see oracle_sensitivity_band), state_dict layout of
    the reference (bilinear_up.0 = ConvTranspose2d [Cin,Cout,2,2], bilinear_up.1 = BatchNorm)."""
    import os
    import aide_b200
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_lb.pt"), weights_only=False)
    b, h, w = 3, 48, 64
    (x1, x2), t1, t2, _ = oracle.synthetic_batch(b, h, w, seed=1234)
    xs = (x1, x2) if kind == "fuse" else (x1,)
    torch.manual_seed(2)
    net = (aide_b200.fuseunet if kind == "fuse" else aide_b200.UNet)(num_classes=2, learned_bilinear=True, mode=mode).to(dev)
    assert list(net.state_dict().keys()) == g["fuse_keys" if kind == "fuse" else "unet_keys"]
    assert tuple(net.up_block1.bilinear_up[0].weight.shape) == (1024, 512, 2, 2)
    net.train()
    y = net(*[x.to(dev) for x in xs])
    c = g["s48x64"][kind]
    assert relmax(y, c["logits"]) < LOGIT_TOL[mode]
    loss = aide_b200.CEMDiceLoss([1., 1.], [1., 1.], [1., 1.])(y, t2.to(dev))
    assert abs(loss.item() - c["loss_mean"]) < 2e-5
    loss.backward()
    torch.manual_seed(2)
    p = oracle.clone_params((oracle.init_fuseunet if kind == "fuse" else oracle.init_unet)(2, True), requires_grad=True)
    yo = (oracle.fuseunet_forward if kind == "fuse" else oracle.unet_forward)(p, *xs, training=True)
    names = [k for k in p if not oracle.is_buffer(k) and not is_prebn_bias(k) and not k.endswith("bilinear_up.0.bias")]
    go = dict(zip(names, torch.autograd.grad(oracle.ce_dice_mean(yo, t2), [p[k] for k in names])))
    eng = dict(net.named_parameters())
    for k in names:
        assert eng[k].grad is not None and eng[k].grad.shape == go[k].shape, k
    assert relmax(eng["last_conv1.weight"].grad, c["grad_last_w"]) < 2e-3
    fa = torch.cat([eng[k].grad.detach().cpu().double().flatten() for k in names])
    fb = torch.cat([go[k].detach().double().flatten() for k in names])
    assert 1.0 - torch.nn.functional.cosine_similarity(fa, fb, dim=0).item() < 1e-3
    up = [k for k in names if k.endswith("bilinear_up.0.weight")]
    fa = torch.cat([eng[k].grad.detach().cpu().double().flatten() for k in up])
    fb = torch.cat([go[k].detach().double().flatten() for k in up])
    assert 1.0 - torch.nn.functional.cosine_similarity(fa, fb, dim=0).item() < 1e-3       # the transposed-conv weights
    assert relmax(net.up_block2.bilinear_up[1].running_var, c["rv_up"]) < 1e-4
    net.eval()
    with torch.no_grad():
        ye = net(*[x.to(dev) for x in xs])
    assert relmax(ye, c["logits_eval"]) < 1e-3


def test_graphed_eval_forward_equals_eager_and_is_faster(oracle):
    """Single-slice evaluation forwards (trainchaos_proposed_30cases1labeled.py:403-409) replayed as one CUDA graph:
    bit-identical to the eager launch sequence, re-captured when the weights change."""
    import time
    import aide_b200 as A
    dev = torch.device("cuda:0")
    net = build("fuse", "parity", dev).eval()
    g = torch.Generator().manual_seed(5)
    xs = [tuple(torch.randn(1, 3, 256, 256, generator=g).to(dev) for _ in range(2)) for _ in range(3)]
    f = net.graphed_eval(*xs[0])
    with torch.no_grad():
        for x in xs:
            eager = net(*x).clone()
            assert torch.equal(f(*x), eager)
        mask = A.predict_mask(f(*xs[0]))
        assert mask.shape == (1, 256, 256) and mask.dtype == torch.uint8
        # weights change -> the captured operand planes are stale -> re-capture
        net.last_conv1.bias.add_(0.25)
        net.up_block4.block.conv2.weight.mul_(1.5)
        assert torch.equal(f(*xs[1]), net(*xs[1]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            f(*xs[2])
        torch.cuda.synchronize()
        t_graph = (time.perf_counter() - t0) / 20
        t0 = time.perf_counter()
        for _ in range(20):
            net(*xs[2])
        torch.cuda.synchronize()
        t_eager = (time.perf_counter() - t0) / 20
    print(f"single-slice eval forward 256x256: graph {t_graph * 1e3:.2f} ms ({1 / t_graph:.0f} slices/s), "
          f"eager {t_eager * 1e3:.2f} ms")
    assert t_graph < t_eager


@pytest.mark.parametrize("mode", ["exact", "parity"])
@pytest.mark.parametrize("kind", ["fuseunetsa", "fuseunetsaseparate", "unetsa"])
def test_attention_variants_vs_reference_vectors(oracle, kind, mode):
    """fuseunetsa / fuseunetsaseparate / UNetsa (fuseunet.py:93-325, UNet.py:168-208): every encoder block gated by its
    Spatial_Attention (netblocks.py:68-89).  Forward vs vectors frozen from the unmodified reference (golden_sa.pt),
    gradients vs the reference values (attention parameters, shortest path) and the live oracle (direction of the whole
    gradient), reference state_dict layout, eval mode on the running statistics."""
    import aide_b200
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_sa.pt"), weights_only=False)
    b, h, w = 3, 48, 64
    (x1, x2), t1, t2, _ = oracle.synthetic_batch(b, h, w, seed=1234)
    ctor, init, fwd, xs = {
        "fuseunetsa": (aide_b200.fuseunetsa, lambda: oracle.init_fuseunetsa(2), oracle.fuseunetsa_forward, (x1, x2)),
        "fuseunetsaseparate": (aide_b200.fuseunetsaseparate, lambda: oracle.init_fuseunetsa(2, True),
                               oracle.fuseunetsaseparate_forward, (x1, x2)),
        "unetsa": (aide_b200.UNetsa, lambda: oracle.init_unetsa(2), oracle.unetsa_forward, (x1,))}[kind]
    torch.manual_seed(2)
    net = ctor(num_classes=2, mode=mode).to(dev)
    assert list(net.state_dict().keys()) == g[kind + "_keys"]
    net.train()
    y = net(*[x.to(dev) for x in xs])
    c = g["s48x64"][kind]
    e_logit = relmax(y, c["logits"])
    assert e_logit < LOGIT_TOL[mode], e_logit
    loss = aide_b200.CEMDiceLoss([1., 1.], [1., 1.], [1., 1.])(y, t2.to(dev))
    assert abs(loss.item() - c["loss_mean"]) < 2e-5
    loss.backward()
    eng = {k: v.grad for k, v in net.named_parameters()}
    assert relmax(eng["last_conv1.weight"], c["grad_last_w"]) < 2e-3
    # live oracle: whole-gradient direction, and the attention branch's own parameters
    torch.manual_seed(2)
    p = oracle.clone_params(init(), requires_grad=True)
    yo = fwd(p, *xs, training=True)
    names = [k for k in p if not oracle.is_buffer(k) and not is_prebn_bias(k)]
    go = dict(zip(names, torch.autograd.grad(oracle.ce_dice_mean(yo, t2), [p[k] for k in names])))
    for k in names:
        assert eng[k] is not None and eng[k].shape == go[k].shape, k
    fa = torch.cat([eng[k].detach().cpu().double().flatten() for k in names])
    fb = torch.cat([go[k].detach().double().flatten() for k in names])
    gap = 1.0 - torch.nn.functional.cosine_similarity(fa, fb, dim=0).item()
    sa = [k for k in names if "sa" in k.split(".")[0]]
    fa = torch.cat([eng[k].detach().cpu().double().flatten() for k in sa])
    fb = torch.cat([go[k].detach().double().flatten() for k in sa])
    gap_sa = 1.0 - torch.nn.functional.cosine_similarity(fa, fb, dim=0).item()
    print(f"{kind}/{mode}: logits {e_logit:.2e}, 1-cos(all grads) {gap:.2e}, 1-cos(attention grads) {gap_sa:.2e}")
    assert gap < 1e-3 and gap_sa < 1e-3
    sd = net.state_dict()
    for k, v in c["sa_bn_rm"].items():
        assert relmax(sd[k], v) < 1e-4, k
    for k, v in c["sa_bn_rv"].items():
        assert relmax(sd[k], v) < 1e-4, k
    assert int(sd[[k for k in sd if k.endswith("bn.num_batches_tracked") and "sa" in k][0]]) == 1
    net.eval()
    with torch.no_grad():
        ye = net(*[x.to(dev) for x in xs])
    assert relmax(ye, c["logits_eval"]) < 1e-3


def test_unet_width_variants_run(oracle):
    """UNet128 / UNet32 / UNet16 / UNet8 / UNet4 (UNet.py:210-368): the same network at other widths; widths below the
    tcgen05 tile granularity run on the fp32 CUDA-core kernels.  Forward + backward against the oracle's UNet forward
    evaluated on the module's own parameters."""
    import aide_b200
    dev = torch.device("cuda:0")
    (x1, _), t1, _, _ = oracle.synthetic_batch(2, 32, 32, seed=5)
    for cls in (aide_b200.UNet32, aide_b200.UNet16, aide_b200.UNet4):
        torch.manual_seed(3)
        net = cls(num_classes=2).to(dev).train()
        p = oracle.clone_params({k: v.detach().cpu() for k, v in net.state_dict().items()}, requires_grad=True)
        y = net(x1.to(dev))
        yo = oracle.unet_forward(p, x1, training=True)
        assert relmax(y, yo) < 2e-4, cls.__name__
        aide_b200.DiceLoss()(y, t1.to(dev)).backward()
        gw = torch.autograd.grad(oracle.dice_loss_mean(yo, t1), p["last_conv1.weight"])[0]
        assert relmax(net.last_conv1.weight.grad, gw) < 2e-3, cls.__name__


@pytest.mark.parametrize("kind", ["fuse", "unet", "unetsa"])
def test_fused_eval_forward_is_bit_identical_to_unfused(oracle, kind, monkeypatch):
    """Eval mode folds BatchNorm(running statistics) + ReLU (+ max-pool) into the conv epilogue
    (aide_conv3x3_bn_relu_fwd) and derives every layer's scale / shift in one launch; the planes it writes must equal
    the conv -> finalize -> apply sequence bit for bit, and the launch count must drop."""
    import aide_b200
    from aide_b200 import lib
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    net = (aide_b200.UNetsa(num_classes=2, mode="parity") if kind == "unetsa" else build(kind, "parity", dev)).to(dev)
    sd = net.state_dict()
    for k in sd:                                      # non-trivial running statistics / affine parameters
        if k.endswith("running_mean") or (k.endswith(".bias") and "bn" in k):
            sd[k] = torch.randn_like(sd[k]) * 0.1
        if k.endswith("running_var"):
            sd[k] = torch.rand_like(sd[k]) + 0.5
    net.load_state_dict(sd)
    net.eval()
    g = torch.Generator().manual_seed(8)
    xs = tuple(torch.randn(2, 3, 64, 96, generator=g).to(dev) for _ in range(2 if kind == "fuse" else 1))
    with torch.no_grad():
        monkeypatch.setenv("AIDE_B200_EVAL_FUSE", "0")
        n0 = lib.aide_launch_count()
        y_ref = net(*xs).clone()
        n_unfused = lib.aide_launch_count() - n0
        monkeypatch.setenv("AIDE_B200_EVAL_FUSE", "1")
        n0 = lib.aide_launch_count()
        y = net(*xs)
        n_fused = lib.aide_launch_count() - n0
    assert torch.equal(y, y_ref)
    print(f"{kind}: eval forward launches {n_unfused} -> {n_fused}")
    assert n_fused < 0.7 * n_unfused


def test_nograd_module_forward_replays_a_graph_and_follows_weight_updates():
    """The no-grad forwards of an unmodified reference script (8 pseudo-label forwards per step in train mode,
    trainchaos_proposed_30cases1labeled.py:263-272; single-slice evaluation, :373-496) are captured into a CUDA graph on
    their second call.  The replay must be indistinguishable from the eager forward: same logits bit for bit, same
    running statistics / num_batches_tracked, it must follow in-place parameter updates without re-capture, and every
    returned tensor must stay valid when the next forward runs."""
    import aide_b200 as A
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    a = A.fuseunet(num_classes=2, mode="parity").to(dev).train()
    b = A.fuseunet(num_classes=2, mode="parity").to(dev).train()
    b.load_state_dict(a.state_dict())
    b._auto_graph = False
    g = torch.Generator().manual_seed(3)
    xs = [(torch.randn(2, 3, 64, 64, generator=g).to(dev), torch.randn(2, 3, 64, 64, generator=g).to(dev)) for _ in range(6)]
    outs = []
    with torch.no_grad():
        for i, x in enumerate(xs):
            if i == 4:                                     # an optimiser-style in-place update between forwards
                for pa, pb in zip(a.parameters(), b.parameters()):
                    d = 0.01 * torch.randn(pa.shape, generator=g).to(dev)
                    pa.add_(d)
                    pb.add_(d)
            ya, yb = a(*x), b(*x)
            assert torch.equal(ya, yb), i
            outs.append((ya, yb.clone()))
        graphs = [v for v in a._auto_graphs.values() if v]
        assert len(graphs) == 1 and graphs[0].n_kernels > 50           # calls 2.. were replays of one captured forward
        for ya, yb in outs:
            assert torch.equal(ya, yb)                                  # earlier results were not overwritten
        for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
            assert torch.equal(va, vb), ka                              # running statistics, num_batches_tracked
        a.eval(), b.eval()
        for i, x in enumerate(xs[:3]):
            assert torch.equal(a(x[0][:1], x[1][:1]), b(x[0][:1], x[1][:1])), i
        assert len([v for v in a._auto_graphs.values() if v]) == 2
    # a forward WITH grad in between uses the eager path and its own weight planes
    x = xs[0]
    la = a.train()(*x).sum()
    lb = b.train()(*x).sum()
    la.backward(), lb.backward()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa.grad, pb.grad)
