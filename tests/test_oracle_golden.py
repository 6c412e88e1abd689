"""CPU: the oracle (oracle/aide_oracle.py) against the golden vectors frozen from the unmodified
reference (tests/golden/make_golden.py).  Tolerances are tiny but non-zero: golden values were produced
with 8 CPU threads; a different thread count / CPU changes fp32 summation order."""
import numpy as np
import pytest
import torch


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_init_matches_reference(golden, oracle):
    torch.manual_seed(2)
    pf, pu = oracle.init_fuseunet(2), oracle.init_unet(2)
    g = golden["init"]
    assert list(pf.keys()) == g["fuse_keys"] and list(pu.keys()) == g["unet_keys"]
    assert {k: tuple(v.shape) for k, v in pf.items()} == g["fuse_shapes"]
    assert {k: tuple(v.shape) for k, v in pu.items()} == g["unet_shapes"]
    assert pf["modal1_downblock1.block.conv1.weight"].double().sum().item() == g["fuse_first_w_sum"]
    n = lambda p: sum(v.numel() for k, v in p.items() if not oracle.is_buffer(k))
    assert n(pf) == g["fuse_n_params"] == 26675074 and n(pu) == g["unet_n_params"] == 34527106


@pytest.mark.parametrize("tag,shape", [("s32", (2, 32, 32)), ("s48x64", (3, 48, 64))])
def test_small_networks(golden, oracle, tag, shape):
    b, h, w = shape
    (x1, x2), t1, t2, _ = oracle.synthetic_batch(b, h, w, seed=1234)
    g = golden[tag]["fuse"]
    torch.manual_seed(2)
    p = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    y = oracle.fuseunet_forward(p, x1, x2, training=True)
    assert rel(y.detach(), g["logits"]) < 1e-5
    li = oracle.ce_dice_per_image(y, t2)
    assert torch.allclose(li.detach(), g["loss_img"], rtol=1e-5, atol=1e-6)
    assert torch.equal(li.sort()[1], g["sort_idx"])
    lm = oracle.ce_dice_mean(y, t2)
    assert abs(lm.item() - g["loss_mean"]) < 1e-5
    assert abs(oracle.dice_fn(y.detach(), t2).item() - g["dice_fn"]) < 1e-5
    names = [k for k in p if not oracle.is_buffer(k)]
    grads = dict(zip(names, torch.autograd.grad(lm, [p[k] for k in names])))
    assert rel(grads["last_conv1.weight"], g["grad_last_w"]) < 1e-4
    assert rel(grads["modal1_downblock1.block.conv1.weight"], g["grad_first_w"]) < 1e-3
    assert rel(p["modal1_downblock1.block.bn1.running_mean"].detach(), g["rm_first"]) < 1e-5
    pe = {k: v.detach() for k, v in p.items()}
    assert rel(oracle.fuseunet_forward(pe, x1, x2, training=False), g["logits_eval"]) < 1e-4
    # UNet
    gu = golden[tag]["unet"]
    torch.manual_seed(2)
    oracle.init_fuseunet(2)                       # the golden run built fuseunet first (RNG order)
    torch.manual_seed(2)
    pu = oracle.init_unet(2)
    yu = oracle.unet_forward(pu, x1, training=True)
    assert rel(yu, gu["logits"]) < 1e-5
    assert abs(oracle.dice_loss_mean(yu, t1).item() - gu["dice_loss"]) < 1e-5


def test_loss_vectors(golden, oracle):
    g = golden["loss"]
    lg, lg2, tg = g["logits"], g["logits2"], g["targets"]
    assert torch.allclose(oracle.ce_dice_per_image(lg, tg), g["cedice_img"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(oracle.ce_dice_per_image(lg, tg, (0.5, 2.0), (0.3, 1.7)), g["cedice_img_w"], rtol=1e-6, atol=1e-6)
    assert abs(oracle.ce_dice_mean(lg, tg).item() - g["cedice_mean"]) < 1e-6
    assert abs(oracle.dice_loss_mean(lg, tg).item() - g["dice_loss"]) < 1e-6
    assert abs(oracle.dice_fn(lg, tg).item() - g["dice_fn"]) < 1e-6
    assert torch.allclose(oracle.ce_per_pixel(lg, tg), g["ce_none"], rtol=1e-6, atol=1e-6)
    q, wm = oracle.pseudo_label([lg, lg2], 1.0)
    assert torch.allclose(q, g["q"], atol=1e-7) and torch.allclose(wm, g["wmap"], atol=1e-6)
    assert abs(oracle.weighted_mse_mean(lg2, q, wm).item() - g["wmse"]) < 1e-7
    assert torch.allclose(oracle.pseudo_label([lg, lg2], 2.0, "chaos")[0], g["q_T2_chaos"], atol=1e-6)
    assert torch.allclose(oracle.pseudo_label([lg, lg2], 2.0, "kidney")[0], g["q_T2_kidney"], atol=1e-6)


def test_aide_step_small(golden, oracle):
    g = golden["aide64"]
    (x1, x2), t1, t2, augs = oracle.synthetic_batch(4, 64, 64, seed=1234, n_aug=4)
    torch.manual_seed(2)
    p1 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    p2 = oracle.clone_params(oracle.init_fuseunet(2), requires_grad=True)
    r = oracle.aide_step(oracle.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, 0.25)
    assert rel(r["out1"], g["out1"]) < 1e-4 and rel(r["out2"], g["out2"]) < 1e-4
    assert torch.allclose(r["pre1"].detach(), g["pre1"], rtol=1e-5) and torch.allclose(r["pre2"].detach(), g["pre2"], rtol=1e-5)
    assert torch.equal(r["indx1"], g["indx1"]) and torch.equal(r["indx2"], g["indx2"])
    assert abs(r["loss1"].item() - g["loss1"]) < 1e-5 and abs(r["loss2"].item() - g["loss2"]) < 1e-5
    assert abs(r["dice1"].item() - g["dice1"]) < 1e-4
    assert rel(r["grads1"]["last_conv1.weight"], g["grad1_last_w"]) < 1e-3
    assert int(p1["modal1_downblock1.block.bn1.num_batches_tracked"]) == g["nbt"] == 5
    assert rel(p1["up_block4.block.bn2.running_mean"].detach(), g["rm_last"]) < 1e-4


def test_known_answers_256(golden, oracle):
    """SURVEY.md section 8c known answers (fuseunet / UNet, seed 2, B=4, 256x256)."""
    g = golden["ka256"]
    torch.manual_seed(2)
    pf, pu = oracle.init_fuseunet(2), oracle.init_unet(2)
    gen = torch.Generator().manual_seed(1234)
    x1 = torch.randn(4, 3, 256, 256, generator=gen)
    x2 = torch.randn(4, 3, 256, 256, generator=gen)
    t = (torch.rand(4, 256, 256, generator=gen) < 0.08).long()
    assert int(t.sum()) == g["t_sum"] == 21057
    with torch.no_grad():
        yf = oracle.fuseunet_forward(pf, x1, x2, True)
    assert abs(yf.double().sum().item() - g["fuse_sum"]) < 0.5           # 85922.62
    assert abs(yf.abs().max().item() - g["fuse_absmax"]) < 1e-4          # 1.979134
    assert rel(yf[:, :, ::8, ::8], g["fuse_logits_sub"]) < 1e-4
    li = oracle.ce_dice_per_image(yf, t)
    assert torch.allclose(li, g["fuse_loss_img"], rtol=1e-5)             # [1.583495, 1.581936, 1.582125, 1.583882]
    assert li.sort()[1].tolist() == g["fuse_sort"].tolist() == [1, 2, 0, 3]
    packed = torch.from_numpy(np.packbits((yf.argmax(1) == 1).numpy().reshape(-1)))
    flips = int(np.unpackbits((packed ^ g["fuse_argmax_packed"]).numpy()).sum())
    assert flips <= 2, flips                                              # numerical ties only (min margin 1.5e-6)


def test_reverseaug_restatement_matches_pil_bit_for_bit(oracle):
    """The numpy restatement of Pillow's flip + rotate(BILINEAR) (oracle.reverseaug_plane) against the reference's own
    code path through PIL (oracle.reverseaug_pil = trainchaos_proposed_30cases1labeled.py:81-95), including PIL's
    exact fast paths; the host-side matrix builder of the product must agree with the oracle's."""
    import random
    import numpy as np
    from aide_b200.augment import rotate_matrix
    rnd = random.Random(0)
    rs = np.random.RandomState(0)
    for (H, W) in [(32, 32), (48, 64), (17, 23)]:
        for trial in range(24):
            deg = rnd.choice([0, 90, 180, 270, -90, 360]) if trial < 6 else rnd.uniform(-60, 60)
            flip = rnd.random() < 0.5
            x = rs.randn(H, W).astype(np.float32)
            augset = {"augno": [1], "hflip1": [int(flip)], "degree1": [deg]}
            ref = oracle.reverseaug_pil(augset, [torch.from_numpy(x)[None, None].clone()], 1)[0][0, 0].numpy()
            got = oracle.reverseaug_plane(np.ascontiguousarray(x), flip, 0 - deg)
            assert np.array_equal(ref, got), (H, W, deg, flip)
            assert rotate_matrix(0 - deg, W, H) == oracle.rotate_matrix(0 - deg, W, H)


@pytest.mark.parametrize("tag,shape", [("s32", (2, 32, 32)), ("s48x64", (3, 48, 64))])
def test_learned_bilinear_networks(oracle, tag, shape):
    """learned_bilinear=True (ConvTranspose2d(k=2,s=2) -> BN -> ReLU up path, netblocks.py:11-14 / UNet.py:6-9): the
    oracle against vectors frozen from the unmodified reference (tests/golden/make_golden.py lb)."""
    import os
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "golden_lb.pt"), weights_only=False)
    b, h, w = shape
    (x1, x2), t1, t2, _ = oracle.synthetic_batch(b, h, w, seed=1234)
    for kind, init, fwd, xs, keys in (("fuse", oracle.init_fuseunet, oracle.fuseunet_forward, (x1, x2), "fuse_keys"),
                                      ("unet", oracle.init_unet, oracle.unet_forward, (x1,), "unet_keys")):
        torch.manual_seed(2)
        p = oracle.clone_params(init(2, True), requires_grad=True)
        assert list(p.keys()) == g[keys]
        y = fwd(p, *xs, training=True)
        c = g[tag][kind]
        assert torch.allclose(y, c["logits"], rtol=0, atol=2e-5 * c["logits"].abs().max().item())
        loss = oracle.ce_dice_mean(y, t2)
        assert abs(loss.item() - c["loss_mean"]) < 2e-6
        names = ["last_conv1.weight", "up_block4.bilinear_up.0.weight", "up_block4.bilinear_up.0.bias",
                 "up_block1.bilinear_up.0.weight"]
        gl, g4w, g4b, g1w = torch.autograd.grad(loss, [p[k] for k in names])
        for got, want in ((gl, c["grad_last_w"]), (g4w, c["grad_up4_w"]), (g4b, c["grad_up4_b"]),
                          (g1w[::16, ::16], c["grad_up1_w_sub"])):
            assert (got - want).abs().max().item() <= 2e-3 * want.abs().max().item() + 1e-9
        assert torch.allclose(p["up_block2.bilinear_up.1.running_var"], c["rv_up"], rtol=1e-5)
        with torch.no_grad():
            ye = fwd({k: v.detach() for k, v in p.items()}, *xs, training=False)
        assert torch.allclose(ye, c["logits_eval"], rtol=0, atol=1e-4 * c["logits_eval"].abs().max().item())


def test_forward_augmentation_restatement_matches_pil():
    """oracle.rotate_u8 (numpy restatement of Pillow's 8-bit bilinear rotate) against PIL itself, bit for bit -- the pin
    of the GPU forward-augmentation kernel (datasetchaos_proposed/transform.py:81-106)."""
    import random
    import numpy as np
    from PIL import Image
    from oracle import aide_oracle as O
    rng = np.random.default_rng(0)
    for trial in range(24):
        h, w = (64, 64) if trial % 2 else (48, 80)
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        angle = random.Random(trial).random() * 120 - 60 if trial < 20 else [0.0, 180.0, 90.0, 270.0][trial - 20]
        ref = np.array(Image.fromarray(img).rotate(angle, Image.BILINEAR))
        assert np.array_equal(O.rotate_u8(img, angle), ref), (trial, angle)
