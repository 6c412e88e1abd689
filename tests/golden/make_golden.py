"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py            # needs /root/reference

1. imports the reference modules (models_twomodalinputs.fuseunet, models_singlemodalinput.UNet,
   utils.*) with a 2-line matplotlib stub (utils/metrics2d.py:6 imports matplotlib, not installed),
2. asserts that oracle/aide_oracle.py reproduces them BIT-FOR-BIT (same torch build, same thread
   count) -- this is what pins the oracle,
3. freezes known-answer vectors so that the pin travels to machines without /root/reference.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("AIDE_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)

from models_twomodalinputs import fuseunet as RefFuse          # noqa: E402
from models_singlemodalinput import UNet as RefUNet            # noqa: E402
import utils as refutils                                        # noqa: E402

from oracle import aide_oracle as O                             # noqa: E402


def same(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.equal(a, b), f"{what}: oracle != reference (max diff {(a - b).abs().max().item():.3e})"


def check_state(mod, p, what):
    sd = mod.state_dict()
    assert list(sd.keys()) == list(p.keys()), what + ": state_dict keys/order differ"
    for k in sd:
        same(sd[k], p[k], f"{what}.{k}")


def grads_of(loss, tensors):
    return torch.autograd.grad(loss, tensors)


def main():
    out = {}
    torch.set_num_threads(8)
    out["meta"] = dict(torch=torch.__version__, threads=torch.get_num_threads(),
                       reference="lich0031/AIDE @ fcf5d41")

    # ---------------------------------------------------------------- init parity (seed 2)
    torch.manual_seed(2)
    rf, ru = RefFuse(num_classes=2), RefUNet(num_classes=2)
    torch.manual_seed(2)
    pf, pu = O.init_fuseunet(2), O.init_unet(2)
    check_state(rf, pf, "fuseunet")
    check_state(ru, pu, "UNet")
    out["init"] = dict(
        fuse_first_w_sum=rf.modal1_downblock1.block.conv1.weight.double().sum().item(),
        fuse_n_params=sum(p.numel() for p in rf.parameters()),
        unet_n_params=sum(p.numel() for p in ru.parameters()),
        fuse_keys=list(rf.state_dict().keys()), unet_keys=list(ru.state_dict().keys()),
        fuse_shapes={k: tuple(v.shape) for k, v in rf.state_dict().items()},
        unet_shapes={k: tuple(v.shape) for k, v in ru.state_dict().items()})

    crit_img = refutils.CEMDiceLossImage(cediceweight=[1.0, 1.0], ceclassweight=torch.tensor([1.0, 1.0]),
                                         diceclassweight=[1.0, 1.0])
    crit_mean = refutils.CEMDiceLoss(cediceweight=[1.0, 1.0], ceclassweight=torch.tensor([1.0, 1.0]),
                                     diceclassweight=[1.0, 1.0])
    mse_none = refutils.MulticlassMSELoss(reduction="none")

    # ---------------------------------------------------------------- small cases with full tensors
    for tag, (b, h, w) in {"s32": (2, 32, 32), "s48x64": (3, 48, 64)}.items():
        (x1, x2), t1, t2, _ = O.synthetic_batch(b, h, w, seed=1234)
        case = {}
        # fuseunet, train mode, fwd + per-image loss + grads
        torch.manual_seed(2)
        rf = RefFuse(num_classes=2)
        pf = O.clone_params(dict(rf.state_dict()), requires_grad=True)
        rf.train()
        y_ref = rf(x1, x2)
        y_or = O.fuseunet_forward(pf, x1, x2, training=True)
        same(y_ref, y_or, tag + " fuse logits")
        check_state(rf, {k: v.detach() for k, v in pf.items()}, tag + " fuse post-fwd buffers")
        l_ref = crit_img(y_ref, t2)
        l_or = O.ce_dice_per_image(y_or, t2)
        same(l_ref, l_or, tag + " CEMDiceLossImage")
        lm_ref = crit_mean(y_ref, t2)
        same(lm_ref, O.ce_dice_mean(y_or, t2), tag + " CEMDiceLoss")
        names = [k for k, _ in rf.named_parameters()]
        g_ref = grads_of(lm_ref, [p for _, p in rf.named_parameters()])
        g_or = grads_of(O.ce_dice_mean(y_or, t2), [pf[k] for k in names])
        for k, a, c in zip(names, g_ref, g_or):
            same(a, c, tag + " grad " + k)
        same(refutils.Dice_fn(y_ref.detach().clone(), t2), O.dice_fn(y_or.detach(), t2), tag + " Dice_fn")
        case["fuse"] = dict(
            logits=y_ref.detach().clone(), loss_img=l_ref.detach().clone(), loss_mean=lm_ref.item(),
            sort_idx=l_ref.sort()[1].clone(), dice_fn=refutils.Dice_fn(y_ref.detach().clone(), t2).item(),
            grad_last_w=g_ref[names.index("last_conv1.weight")].clone(),
            grad_first_w=g_ref[names.index("modal1_downblock1.block.conv1.weight")].clone(),
            grad_m2_first_w=g_ref[names.index("modal2_downblock1.block.conv1.weight")].clone(),
            grad_bn_w=g_ref[names.index("up_block1.bilinear_up.2.weight")].clone(),
            grad_bn_b=g_ref[names.index("modal1_downblock3.block.bn2.bias")].clone(),
            grad_absmax={k: g.abs().max().item() for k, g in zip(names, g_ref)},
            grad_sum={k: g.double().sum().item() for k, g in zip(names, g_ref)},
            rm_first=rf.modal1_downblock1.block.bn1.running_mean.clone(),
            rv_first=rf.modal1_downblock1.block.bn1.running_var.clone(),
            rv_up=rf.up_block2.bilinear_up[2].running_var.clone())
        # eval-mode forward with the updated running stats
        rf.eval()
        with torch.no_grad():
            ye = rf(x1, x2)
            same(ye, O.fuseunet_forward({k: v.detach() for k, v in pf.items()}, x1, x2, training=False),
                 tag + " fuse eval logits")
        case["fuse"]["logits_eval"] = ye.clone()

        # UNet
        torch.manual_seed(2)
        ru = RefUNet(num_classes=2)
        pu = O.clone_params(dict(ru.state_dict()), requires_grad=True)
        ru.train()
        y_ref = ru(x1)
        y_or = O.unet_forward(pu, x1, training=True)
        same(y_ref, y_or, tag + " unet logits")
        d_ref = refutils.DiceLoss()(y_ref, t1)
        same(d_ref, O.dice_loss_mean(y_or, t1), tag + " DiceLoss")
        names = [k for k, _ in ru.named_parameters()]
        g_ref = grads_of(d_ref, [p for _, p in ru.named_parameters()])
        g_or = grads_of(O.dice_loss_mean(y_or, t1), [pu[k] for k in names])
        for k, a, c in zip(names, g_ref, g_or):
            same(a, c, tag + " unet grad " + k)
        case["unet"] = dict(
            logits=y_ref.detach().clone(), dice_loss=d_ref.item(),
            dice_fn=refutils.Dice_fn(y_ref.detach().clone(), t1).item(),
            grad_last_w=g_ref[names.index("last_conv1.weight")].clone(),
            grad_first_w=g_ref[names.index("down_block1.block.conv1.weight")].clone(),
            grad_absmax={k: g.abs().max().item() for k, g in zip(names, g_ref)},
            grad_sum={k: g.double().sum().item() for k, g in zip(names, g_ref)})
        out[tag] = case

    # ---------------------------------------------------------------- loss-only vectors (no network)
    g = torch.Generator().manual_seed(77)
    lg = torch.randn(5, 2, 24, 40, generator=g) * 2.0
    lg2 = torch.randn(5, 2, 24, 40, generator=g) * 2.0
    tg = (torch.rand(5, 24, 40, generator=g) < 0.3).long()
    tg[3] = 0                                   # empty-target image (Dice_fn special case)
    lg[3, 1] = lg[3, 0] - 1.0                   # ... predicted empty as well -> dice 1
    q, wmap = O.pseudo_label([lg, lg2], 1.0)
    mse_ref = (wmap * mse_none(lg2, q)).mean()
    same(mse_ref, O.weighted_mse_mean(lg2, q, wmap), "weighted mse")
    cw = torch.tensor([0.3, 1.7])
    crit_w = refutils.CEMDiceLossImage(cediceweight=[0.5, 2.0], ceclassweight=cw, diceclassweight=[1.0, 1.0])
    same(crit_w(lg, tg), O.ce_dice_per_image(lg, tg, (0.5, 2.0), (0.3, 1.7)), "weighted CEMDiceLossImage")
    out["loss"] = dict(
        logits=lg, logits2=lg2, targets=tg,
        cedice_img=crit_img(lg, tg).clone(), cedice_img_w=crit_w(lg, tg).clone(),
        cedice_mean=crit_mean(lg, tg).item(), dice_loss=refutils.DiceLoss()(lg, tg).item(),
        dice_fn=refutils.Dice_fn(lg.clone(), tg).item(),
        ce_none=refutils.CrossEntropyLoss2d(reduction="none")(lg, tg).clone(),
        q=q.clone(), wmap=wmap.clone(), wmse=mse_ref.item(),
        q_T2_chaos=O.pseudo_label([lg, lg2], 2.0, "chaos")[0].clone(),
        q_T2_kidney=O.pseudo_label([lg, lg2], 2.0, "kidney")[0].clone())
    # reference co-teaching classes (reduction='none' is the only mode that runs)
    ct = {}
    for cls in ("Coteachingloss_dropimage", "Coteachingloss_weightimage", "Coteachingloss_dropregionce",
                "Coteachingloss_dropimagedroppixel"):
        a, b_ = getattr(refutils, cls)(reduction="none")(lg[:4], lg2[:4], tg[:4], 0.5)
        ct[cls] = (float(a), float(b_))
    # _weightimage adds a [num_remember] and a [num_drop] vector (coteach_loss.py:144-147): it only runs when
    # the two counts are equal; the other three also run with unequal counts:
    for cls in ("Coteachingloss_dropimage", "Coteachingloss_dropregionce", "Coteachingloss_dropimagedroppixel"):
        a, b_ = getattr(refutils, cls)(reduction="none")(lg, lg2, tg, 0.4)
        ct[cls + "@5x0.4"] = (float(a), float(b_))
    out["loss"]["coteach_classes"] = ct

    # ---------------------------------------------------------------- AIDE step, small (64x64, B=4)
    (x1, x2), t1, t2, augs = O.synthetic_batch(4, 64, 64, seed=1234, n_aug=4)
    torch.manual_seed(2)
    n1, n2 = RefFuse(num_classes=2), RefFuse(num_classes=2)
    p1 = O.clone_params(dict(n1.state_dict()), requires_grad=True)
    p2 = O.clone_params(dict(n2.state_dict()), requires_grad=True)
    n1.train(); n2.train()
    # reference flow, line by line (trainchaos_proposed_30cases1labeled.py:263-321), reverse-aug = identity
    a1 = [n1(*a).detach() for a in augs]
    a2 = [n2(*a).detach() for a in augs]
    import torch.nn.functional as F
    for i in range(4):
        m1, m2 = F.softmax(a1[i], dim=1), F.softmax(a2[i], dim=1)
        if i == 0:
            pl1, pl2 = m1, m2
        else:
            pl1 += m1; pl2 += m2
    pl1, pl2 = pl1 / 4.0, pl2 / 4.0
    sh = lambda m, T: torch.pow(m, T) / torch.pow(m, T).sum(dim=1).unsqueeze(dim=1)
    pl1, pl2 = sh(pl1, 1.0), sh(pl2, 1.0)
    wm1 = (1.0 - 4.0 * pl1[:, 0] * pl1[:, 1]).unsqueeze(1)
    wm2 = (1.0 - 4.0 * pl2[:, 0] * pl2[:, 1]).unsqueeze(1)
    o1, o2 = n1(x1, x2), n2(x1, x2)
    l1p, l2p = crit_img(o1, t2), crit_img(o2, t1)
    _, i1 = l1p.sort(); _, i2 = l2p.sort()
    rate = 0.25
    l1s1 = crit_img(o1[i2[0:2]], t2[i2[0:2]]).mean(); l2s1 = crit_img(o2[i1[0:2]], t1[i1[0:2]]).mean()
    l1s2 = crit_img(o1[i2[2:]], t2[i2[2:]]).mean(); l2s2 = crit_img(o2[i1[2:]], t1[i1[2:]]).mean()
    l1c = (wm2[i2[2:]] * mse_none(o1[i2[2:]], pl2[i2[2:]])).mean()
    l2c = (wm1[i1[2:]] * mse_none(o2[i1[2:]], pl1[i1[2:]])).mean()
    loss1 = 1.0 * (l1s1 + (1 - rate) * l1s2) + 10.0 * rate * l1c
    loss2 = 1.0 * (l2s1 + (1 - rate) * l2s2) + 10.0 * rate * l2c
    res = O.aide_step(O.fuseunet_forward, p1, p2, (x1, x2), augs, t1, t2, rate)
    same(o1.detach(), res["out1"], "aide out1"); same(o2.detach(), res["out2"], "aide out2")
    same(loss1.detach(), res["loss1"].detach(), "aide loss1"); same(loss2.detach(), res["loss2"].detach(), "aide loss2")
    assert torch.equal(i1, res["indx1"]) and torch.equal(i2, res["indx2"])
    names = [k for k, _ in n1.named_parameters()]
    gr1 = grads_of(loss1, [p for _, p in n1.named_parameters()])
    for k, a, c in zip(names, gr1, [res["grads1"][k] for k in names]):
        same(a, c, "aide grad1 " + k)
    check_state(n1, {k: v.detach() for k, v in p1.items()}, "aide net1 buffers (5 BN updates)")
    out["aide64"] = dict(
        out1=o1.detach().clone(), out2=o2.detach().clone(), pre1=l1p.detach().clone(), pre2=l2p.detach().clone(),
        indx1=i1.clone(), indx2=i2.clone(), loss1=loss1.item(), loss2=loss2.item(),
        q1_sum=pl1.double().sum().item(), w1_sum=wm1.double().sum().item(),
        q2=pl2.clone(), w2=wm2.clone(),
        dice1=refutils.Dice_fn(o1.detach().clone(), t2).item(), dice2=refutils.Dice_fn(o2.detach().clone(), t1).item(),
        grad1_last_w=gr1[names.index("last_conv1.weight")].clone(),
        grad1_absmax={k: g.abs().max().item() for k, g in zip(names, gr1)},
        grad1_sum={k: g.double().sum().item() for k, g in zip(names, gr1)},
        nbt=int(n1.modal1_downblock1.block.bn1.num_batches_tracked),
        rm_last=n1.up_block4.block.bn2.running_mean.clone())

    # ---------------------------------------------------------------- SURVEY 8c known answers @256
    torch.manual_seed(2)
    f = RefFuse(num_classes=2); u = RefUNet(num_classes=2)
    g = torch.Generator().manual_seed(1234)
    x1 = torch.randn(4, 3, 256, 256, generator=g); x2 = torch.randn(4, 3, 256, 256, generator=g)
    t = (torch.rand(4, 256, 256, generator=g) < 0.08).long()
    with torch.no_grad():
        yf = f(x1, x2); yu = u(x1)
        li = crit_img(yf, t)
        pfz = {k: v.clone() for k, v in dict(RefFuse.state_dict(f)).items()}
    torch.manual_seed(2)
    pf0 = O.init_fuseunet(2); pu0 = O.init_unet(2)
    with torch.no_grad():
        same(yf, O.fuseunet_forward(pf0, x1, x2, True), "256 fuse logits")
        same(yu, O.unet_forward(pu0, x1, True), "256 unet logits")
    out["ka256"] = dict(
        t_sum=int(t.sum()), fuse_sum=yf.double().sum().item(), fuse_absmax=yf.abs().max().item(),
        fuse_fg=int((yf.argmax(1) == 1).sum()), fuse_loss_img=li.clone(), fuse_sort=li.sort()[1].clone(),
        unet_sum=yu.double().sum().item(), unet_absmax=yu.abs().max().item(),
        unet_dice_loss=refutils.DiceLoss()(yu, t).item(), unet_dice_fn=refutils.Dice_fn(yu.clone(), t).item(),
        fuse_argmax_packed=torch.from_numpy(__import__("numpy").packbits((yf.argmax(1) == 1).numpy().reshape(-1))),
        fuse_logits_sub=yf[:, :, ::8, ::8].clone(), unet_logits_sub=yu[:, :, ::8, ::8].clone(),
        fuse_margin_min=(yf[:, 1] - yf[:, 0]).abs().min().item())
    # ---- argmax audit band: the SAME forward in fp64 (same weights, same inputs).  d = z1 - z0 is the argmax margin;
    # |d32 - d64| is the reference's own fp32 rounding noise on it.  An engine pixel whose argmax differs from the
    # reference's is a numerical tie iff its fp64 margin lies inside that noise band (tests/test_gpu_network.py).
    p64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in pf0.items()}
    with torch.no_grad():
        y64 = O.fuseunet_forward(p64, x1.double(), x2.double(), True)
    d64 = (y64[:, 1] - y64[:, 0]).reshape(-1)
    d32 = (yf[:, 1] - yf[:, 0]).double().reshape(-1)
    err = (d32 - d64).abs()
    band = err.max().item()
    cand = (d64.abs() < 64.0 * band).nonzero().flatten()
    out["ka256"].update(
        fp64_margin_band=band, fp64_logit_err_rel=((yf.double() - y64).abs().max() / y64.abs().max()).item(),
        fp64_cand_idx=cand.to(torch.int32), fp64_cand_margin=d64[cand].float(), fp64_cand_err=err[cand].float(),
        fp64_logits_sub=y64[:, :, ::8, ::8].clone(),
        fp32_flips_vs_fp64=int(((d32 > 0) != (d64 > 0)).sum()))
    # gradient of the scalar CE+Dice loss at this size (config 2: fuseunet forward+backward, batch 4, 256x256)
    f.train()
    yg = f(x1, x2)              # second train-mode forward of the same module: same logits, BN buffers move again
    lg_ = crit_mean(yg, t)
    names = [k for k, _ in f.named_parameters()]
    gr = grads_of(lg_, [q_ for _, q_ in f.named_parameters()])
    out["ka256"].update(
        fuse_loss_mean=lg_.item(), fuse_grad_last_w=gr[names.index("last_conv1.weight")].clone(),
        fuse_grad_last_b=gr[names.index("last_conv1.bias")].clone(),
        fuse_grad_absmax={k: g_.abs().max().item() for k, g_ in zip(names, gr)},
        fuse_grad_sum={k: g_.double().sum().item() for k, g_ in zip(names, gr)})
    torch.save(out, os.path.join(HERE, "golden.pt"))
    sz = os.path.getsize(os.path.join(HERE, "golden.pt"))
    print("oracle == reference on every check; wrote golden.pt (%.1f KB)" % (sz / 1024))
    for k, v in out["ka256"].items():
        if not torch.is_tensor(v) or v.numel() < 8:
            print("  ka256", k, v)


def main_learned_bilinear():
    """learned_bilinear=True (ConvTranspose2d decoder, netblocks.py:11-14 / UNet.py:6-9): the oracle against the
    unmodified reference, frozen into golden_lb.pt.   python tests/golden/make_golden.py lb"""
    out = {}
    torch.set_num_threads(8)
    crit_mean = refutils.CEMDiceLoss(cediceweight=[1.0, 1.0], ceclassweight=torch.tensor([1.0, 1.0]),
                                     diceclassweight=[1.0, 1.0])
    torch.manual_seed(2)
    rf, ru = RefFuse(num_classes=2, learned_bilinear=True), RefUNet(num_classes=2, learned_bilinear=True)
    torch.manual_seed(2)
    check_state(rf, O.init_fuseunet(2, True), "fuseunet(learned_bilinear)")
    check_state(ru, O.init_unet(2, True), "UNet(learned_bilinear)")
    out["fuse_keys"], out["unet_keys"] = list(rf.state_dict().keys()), list(ru.state_dict().keys())
    for tag, (b, h, w) in {"s32": (2, 32, 32), "s48x64": (3, 48, 64)}.items():
        (x1, x2), t1, t2, _ = O.synthetic_batch(b, h, w, seed=1234)
        case = {}
        for kind, Ref, fwd, xs in (("fuse", RefFuse, O.fuseunet_forward, (x1, x2)), ("unet", RefUNet, O.unet_forward, (x1,))):
            torch.manual_seed(2)
            net = Ref(num_classes=2, learned_bilinear=True)
            p = O.clone_params(dict(net.state_dict()), requires_grad=True)
            net.train()
            y_ref = net(*xs)
            y_or = fwd(p, *xs, training=True)
            same(y_ref, y_or, f"{tag} {kind} lb logits")
            check_state(net, {k: v.detach() for k, v in p.items()}, f"{tag} {kind} lb post-fwd buffers")
            l_ref = crit_mean(y_ref, t2)
            same(l_ref, O.ce_dice_mean(y_or, t2), f"{tag} {kind} lb CEMDiceLoss")
            names = [k for k, _ in net.named_parameters()]
            g_ref = grads_of(l_ref, [q for _, q in net.named_parameters()])
            g_or = grads_of(O.ce_dice_mean(y_or, t2), [p[k] for k in names])
            for k, a, c in zip(names, g_ref, g_or):
                same(a, c, f"{tag} {kind} lb grad {k}")
            net.eval()
            with torch.no_grad():
                ye = net(*xs)
                same(ye, fwd({k: v.detach() for k, v in p.items()}, *xs, training=False), f"{tag} {kind} lb eval logits")
            case[kind] = dict(logits=y_ref.detach().clone(), loss_mean=l_ref.item(), logits_eval=ye.clone(),
                              grad_last_w=g_ref[names.index("last_conv1.weight")].clone(),
                              grad_up1_w_sub=g_ref[names.index("up_block1.bilinear_up.0.weight")][::16, ::16].clone(),
                              grad_up1_w_sum=g_ref[names.index("up_block1.bilinear_up.0.weight")].double().sum().item(),
                              grad_up4_w=g_ref[names.index("up_block4.bilinear_up.0.weight")].clone(),
                              grad_up4_b=g_ref[names.index("up_block4.bilinear_up.0.bias")].clone(),
                              grad_absmax={k: g.abs().max().item() for k, g in zip(names, g_ref)},
                              rv_up=net.up_block2.bilinear_up[1].running_var.clone())
        out[tag] = case
    torch.save(out, os.path.join(HERE, "golden_lb.pt"))
    print("oracle == reference (learned_bilinear) on every check; wrote golden_lb.pt (%.1f KB)"
          % (os.path.getsize(os.path.join(HERE, "golden_lb.pt")) / 1024))


def main_attention():
    """Attention variants (fuseunetsa, fuseunetsaseparate, UNetsa: fuseunet.py:93-325, UNet.py:168-208, Spatial_Attention
    netblocks.py:68-89): the oracle against the unmodified reference, frozen into golden_sa.pt.
    python tests/golden/make_golden.py sa"""
    from models_twomodalinputs import fuseunetsa as RefSA, fuseunetsaseparate as RefSAS
    from models_singlemodalinput import UNetsa as RefUSA
    out = {}
    torch.set_num_threads(8)
    crit_mean = refutils.CEMDiceLoss(cediceweight=[1.0, 1.0], ceclassweight=torch.tensor([1.0, 1.0]),
                                     diceclassweight=[1.0, 1.0])
    cases = (("fuseunetsa", RefSA, lambda: O.init_fuseunetsa(2), O.fuseunetsa_forward, 2),
             ("fuseunetsaseparate", RefSAS, lambda: O.init_fuseunetsa(2, True), O.fuseunetsaseparate_forward, 2),
             ("unetsa", RefUSA, lambda: O.init_unetsa(2), O.unetsa_forward, 1))
    for kind, Ref, init, fwd, n_in in cases:
        torch.manual_seed(2)
        ref = Ref(num_classes=2)
        torch.manual_seed(2)
        check_state(ref, init(), kind)
        out[kind + "_keys"] = list(ref.state_dict().keys())
        for tag, (b, h, w) in {"s32": (2, 32, 32), "s48x64": (3, 48, 64)}.items():
            (x1, x2), t1, t2, _ = O.synthetic_batch(b, h, w, seed=1234)
            xs = (x1, x2)[:n_in]
            torch.manual_seed(2)
            net = Ref(num_classes=2)
            p = O.clone_params(dict(net.state_dict()), requires_grad=True)
            net.train()
            y_ref = net(*xs)
            y_or = fwd(p, *xs, training=True)
            same(y_ref, y_or, f"{tag} {kind} logits")
            check_state(net, {k: v.detach() for k, v in p.items()}, f"{tag} {kind} post-fwd buffers")
            l_ref = crit_mean(y_ref, t2)
            same(l_ref, O.ce_dice_mean(y_or, t2), f"{tag} {kind} CEMDiceLoss")
            names = [k for k, _ in net.named_parameters()]
            g_ref = grads_of(l_ref, [q for _, q in net.named_parameters()])
            g_or = grads_of(O.ce_dice_mean(y_or, t2), [p[k] for k in names])
            for k, a, c in zip(names, g_ref, g_or):
                same(a, c, f"{tag} {kind} grad {k}")
            net.eval()
            with torch.no_grad():
                ye = net(*xs)
                same(ye, fwd({k: v.detach() for k, v in p.items()}, *xs, training=False), f"{tag} {kind} eval logits")
            sa = [k for k in names if "sa" in k.split(".")[0]]
            out.setdefault(tag, {})[kind] = dict(
                logits=y_ref.detach().clone(), loss_mean=l_ref.item(), logits_eval=ye.clone(),
                grad_last_w=g_ref[names.index("last_conv1.weight")].clone(),
                grad_sa={k: g_ref[names.index(k)].clone() for k in sa if g_ref[names.index(k)].numel() <= 4096},
                grad_absmax={k: g.abs().max().item() for k, g in zip(names, g_ref)},
                sa_bn_rm={k: v.clone() for k, v in net.state_dict().items() if k.endswith("bn.running_mean")},
                sa_bn_rv={k: v.clone() for k, v in net.state_dict().items() if k.endswith("bn.running_var")})
    torch.save(out, os.path.join(HERE, "golden_sa.pt"))
    print("oracle == reference (attention variants) on every check; wrote golden_sa.pt (%.1f KB)"
          % (os.path.getsize(os.path.join(HERE, "golden_sa.pt")) / 1024))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sa":
        main_attention()
    elif len(sys.argv) > 1 and sys.argv[1] == "lb":
        main_learned_bilinear()
    else:
        main()
