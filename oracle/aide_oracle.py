"""CPU oracle for the AIDE hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``aide_b200``) never imports it and fails loudly when the CUDA library is
missing.

What it is: a *functional* fp32 restatement (torch CPU ops, explicit parameter
dictionaries, no nn.Module graph) of the reference algorithm:

* ``fuseunet_forward``      <- models_twomodalinputs/fuseunet.py:43-91,
                               models_twomodalinputs/netblocks.py:9-33,128-147
* ``unet_forward``          <- models_singlemodalinput/UNet.py:4-28,110-165
* ``fuseunetsa_forward`` / ``fuseunetsaseparate_forward`` / ``unetsa_forward`` (attention variants)
                            <- models_twomodalinputs/fuseunet.py:93-208,210-325, netblocks.py:68-89,
                               models_singlemodalinput/UNet.py:85-108,168-208
* ``ce_dice_per_image`` ... <- utils/loss2d.py:5-13,35-61,87-154
* ``dice_fn``               <- utils/metrics2d.py:8-29
* ``pseudo_label``          <- train_files/trainchaos_proposed_30cases1labeled.py:274-292,97-101
                               (kidney flavour: trainkidney_proposed_mask1.py:113-117)
* ``coteach_losses``        <- train_files/trainchaos_proposed_30cases1labeled.py:303-321
* ``aide_step``             <- train_files/trainchaos_proposed_30cases1labeled.py:263-325

Parity pinning: the reference ships no tests / golden vectors for this path
("parity unpinned by the reference").  The restatement is
therefore pinned against the reference *itself*: ``tests/golden/make_golden.py``
imports the unmodified reference modules from /root/reference in the build
container, checks this file against them bit-for-bit (same torch ops, same
order) and freezes known-answer vectors under ``tests/golden/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------------------
# parameter construction (same tensors, names and RNG order as the reference modules)
# --------------------------------------------------------------------------------------
def _conv_init(cout: int, cin: int, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """nn.Conv2d default init (kaiming_uniform a=sqrt(5) then bias U(-1/sqrt(fan_in), ..))."""
    w = torch.empty(cout, cin, k, k)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(cin * k * k)
    b = torch.empty(cout)
    torch.nn.init.uniform_(b, -bound, bound)
    return w, b


def _add_conv(p: Params, name: str, cin: int, cout: int, k: int = 3) -> None:
    p[name + ".weight"], p[name + ".bias"] = _conv_init(cout, cin, k)


def _add_bn(p: Params, name: str, c: int) -> None:
    p[name + ".weight"] = torch.ones(c)
    p[name + ".bias"] = torch.zeros(c)
    p[name + ".running_mean"] = torch.zeros(c)
    p[name + ".running_var"] = torch.ones(c)
    p[name + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _add_basic_block(p: Params, name: str, cin: int, cout: int) -> None:
    _add_conv(p, name + ".conv1", cin, cout)
    _add_bn(p, name + ".bn1", cout)
    _add_conv(p, name + ".conv2", cout, cout)
    _add_bn(p, name + ".bn2", cout)


def _add_up_block(p: Params, name: str, cin: int, cprev: int, cout: int, learned_bilinear: bool = False) -> None:
    if learned_bilinear:
        # bilinear_up = Sequential(ConvTranspose2d(k=2,s=2), BatchNorm2d, ReLU) (netblocks.py:11-14) -> indices 0, 1.
        # nn.ConvTranspose2d default init: kaiming_uniform(a=sqrt 5) on [Cin,Cout,2,2]; torch computes fan_in from
        # dim 1, i.e. Cout * 4, for the bias bound as well
        w = torch.empty(cin, cprev, 2, 2)
        torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        b = torch.empty(cprev)
        bound = 1.0 / math.sqrt(cprev * 4)
        torch.nn.init.uniform_(b, -bound, bound)
        p[name + ".bilinear_up.0.weight"], p[name + ".bilinear_up.0.bias"] = w, b
        _add_bn(p, name + ".bilinear_up.1", cprev)
    else:
        # bilinear_up = Sequential(Upsample, Conv2d, BatchNorm2d, ReLU) -> indices 1, 2 hold tensors
        _add_conv(p, name + ".bilinear_up.1", cin, cprev)
        _add_bn(p, name + ".bilinear_up.2", cprev)
    _add_basic_block(p, name + ".block", cprev * 2, cout)


FUSE_ENC = [  # (modal1 in, modal2 in, out) per level -- fuseunet.py:12-33
    (3, 3, 32), (64, 32, 64), (128, 64, 128), (256, 128, 256), (512, 256, 512)]
UNET_ENC = [(3, 64), (64, 128), (128, 256), (256, 512), (512, 1024)]  # UNet.py:139-143
DECODER = [(1024, 512, 512), (512, 256, 256), (256, 128, 128), (128, 64, 64)]


def init_fuseunet(num_classes: int = 2, learned_bilinear: bool = False) -> Params:
    """Same state_dict (keys, shapes, values under the same torch seed) as fuseunet()."""
    p: Params = {}
    for lvl, (c1, _, co) in enumerate(FUSE_ENC, 1):
        _add_basic_block(p, f"modal1_downblock{lvl}.block", c1, co)
    for lvl, (_, c2, co) in enumerate(FUSE_ENC, 1):
        _add_basic_block(p, f"modal2_downblock{lvl}.block", c2, co)
    for i, (ci, cp, co) in enumerate(DECODER, 1):
        _add_up_block(p, f"up_block{i}", ci, cp, co, learned_bilinear)
    _add_conv(p, "last_conv1", 64, num_classes, k=1)
    return p


def init_unet(num_classes: int = 2, learned_bilinear: bool = False) -> Params:
    p: Params = {}
    for lvl, (ci, co) in enumerate(UNET_ENC, 1):
        _add_basic_block(p, f"down_block{lvl}.block", ci, co)
    for i, (ci, cp, co) in enumerate(DECODER, 1):
        _add_up_block(p, f"up_block{i}", ci, cp, co, learned_bilinear)
    _add_conv(p, "last_conv1", 64, num_classes, k=1)
    return p


def _add_spatial_attention(p: Params, name: str, c: int, reduction: int = 16) -> None:
    """Spatial_Attention.__init__ (netblocks.py:69-79): conv1 1x1 C->r, conv2 / conv3 3x3 r->r (dilated), conv4 1x1 r->1,
    BatchNorm2d(1) -- tensors created in this order."""
    r = c // reduction
    _add_conv(p, name + ".conv1", c, r, k=1)
    _add_conv(p, name + ".conv2", r, r, k=3)
    _add_conv(p, name + ".conv3", r, r, k=3)
    _add_conv(p, name + ".conv4", r, 1, k=1)
    _add_bn(p, name + ".bn", 1)


def init_fuseunetsa(num_classes: int = 2, separate: bool = False, reduction: int = 16,
                    learned_bilinear: bool = False) -> Params:
    """Same state_dict as fuseunetsa() (fuseunet.py:94-136) / fuseunetsaseparate() (:211-253): every encoder block is
    followed by its Spatial_Attention; the `separate` variant's modal-1 encoder takes its own previous level
    (32 -> 64 -> ...) instead of the fused concat."""
    p: Params = {}
    for m in (1, 2):
        for lvl, (c1, c2, co) in enumerate(FUSE_ENC, 1):
            cin = c2 if (m == 2 or separate) else c1
            _add_basic_block(p, f"modal{m}_downblock{lvl}.block", cin, co)
            _add_spatial_attention(p, f"modal{m}_sa{lvl}", co, reduction)
    for i, (ci, cp, co) in enumerate(DECODER, 1):
        _add_up_block(p, f"up_block{i}", ci, cp, co, learned_bilinear)
    _add_conv(p, "last_conv1", 64, num_classes, k=1)
    return p


def init_unetsa(num_classes: int = 2, learned_bilinear: bool = False) -> Params:
    """Same state_dict as UNetsa() (UNet.py:169-188)."""
    p: Params = {}
    for lvl, (ci, co) in enumerate(UNET_ENC, 1):
        _add_basic_block(p, f"down_block{lvl}.block", ci, co)
        _add_spatial_attention(p, f"sa{lvl}", co)
    for i, (ci, cp, co) in enumerate(DECODER, 1):
        _add_up_block(p, f"up_block{i}", ci, cp, co, learned_bilinear)
    _add_conv(p, "last_conv1", 64, num_classes, k=1)
    return p


def is_buffer(name: str) -> bool:
    return name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def clone_params(p: Params, requires_grad: bool = False) -> Params:
    out = {}
    for k, v in p.items():
        t = v.detach().clone()
        if requires_grad and not is_buffer(k):
            t.requires_grad_(True)
        out[k] = t
    return out


# --------------------------------------------------------------------------------------
# network forward
# --------------------------------------------------------------------------------------
def _conv_bn_relu(p: Params, conv: str, bn: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    # netblocks.py:30-33 : relu(bn(conv(x))); conv has bias; BN eps 1e-5, momentum 0.1
    x = F.conv2d(x, p[conv + ".weight"], p[conv + ".bias"], stride=1, padding=1)
    if training:
        p[bn + ".num_batches_tracked"] += 1
    x = F.batch_norm(x, p[bn + ".running_mean"], p[bn + ".running_var"],
                     p[bn + ".weight"], p[bn + ".bias"], training, BN_MOMENTUM, BN_EPS)
    return F.relu(x)


def _basic_block(p: Params, name: str, x: torch.Tensor, training: bool) -> torch.Tensor:
    x = _conv_bn_relu(p, name + ".conv1", name + ".bn1", x, training)
    return _conv_bn_relu(p, name + ".conv2", name + ".bn2", x, training)


def _up_block(p: Params, name: str, skip: torch.Tensor, x: torch.Tensor, training: bool) -> torch.Tensor:
    # netblocks.py:137-147 : up -> conv/bn/relu -> cat((x, skip)) -> basic_block
    if name + ".bilinear_up.0.weight" in p:        # learned_bilinear=True (netblocks.py:11-14): ConvTranspose2d -> BN -> ReLU
        bn = name + ".bilinear_up.1"
        x = F.conv_transpose2d(x, p[name + ".bilinear_up.0.weight"], p[name + ".bilinear_up.0.bias"], stride=2)
        if training:
            p[bn + ".num_batches_tracked"] += 1
        x = F.relu(F.batch_norm(x, p[bn + ".running_mean"], p[bn + ".running_var"], p[bn + ".weight"], p[bn + ".bias"],
                                training, BN_MOMENTUM, BN_EPS))
        x = torch.cat((x, skip), dim=1)
        return _basic_block(p, name + ".block", x, training)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    x = _conv_bn_relu(p, name + ".bilinear_up.1", name + ".bilinear_up.2", x, training)
    x = torch.cat((x, skip), dim=1)
    return _basic_block(p, name + ".block", x, training)


def _decoder(p: Params, skips: Sequence[torch.Tensor], x: torch.Tensor, training: bool) -> torch.Tensor:
    for i, skip in enumerate(reversed(skips), 1):
        x = _up_block(p, f"up_block{i}", skip, x, training)
    return F.conv2d(x, p["last_conv1.weight"], p["last_conv1.bias"])


def fuseunet_forward(p: Params, modal1: torch.Tensor, modal2: torch.Tensor,
                     training: bool = True) -> torch.Tensor:
    """fuseunet.py:43-91.  Mutates the BN buffers in ``p`` when training (like the module)."""
    y = _basic_block(p, "modal1_downblock1.block", modal1, training)
    x = _basic_block(p, "modal2_downblock1.block", modal2, training)
    fused = [torch.cat((y, x), dim=1)]
    for lvl in range(2, 6):
        y = _basic_block(p, f"modal1_downblock{lvl}.block", F.max_pool2d(fused[-1], 2, 2), training)
        x = _basic_block(p, f"modal2_downblock{lvl}.block", F.max_pool2d(x, 2, 2), training)
        fused.append(torch.cat((y, x), dim=1))
    return _decoder(p, fused[:4], fused[4], training)


def _spatial_attention(p: Params, name: str, x: torch.Tensor, training: bool, dilation: int = 4) -> torch.Tensor:
    """Spatial_Attention.forward (netblocks.py:82-89) -> gate [N,1,H,W]."""
    y = F.conv2d(x, p[name + ".conv1.weight"], p[name + ".conv1.bias"])
    y = F.conv2d(y, p[name + ".conv2.weight"], p[name + ".conv2.bias"], padding=dilation, dilation=dilation)
    y = F.conv2d(y, p[name + ".conv3.weight"], p[name + ".conv3.bias"], padding=dilation, dilation=dilation)
    y = F.conv2d(y, p[name + ".conv4.weight"], p[name + ".conv4.bias"])
    if training:
        p[name + ".bn.num_batches_tracked"] += 1
    y = F.batch_norm(y, p[name + ".bn.running_mean"], p[name + ".bn.running_var"], p[name + ".bn.weight"],
                     p[name + ".bn.bias"], training, BN_MOMENTUM, BN_EPS)
    return torch.sigmoid(y)


def fuseunetsa_forward(p: Params, modal1: torch.Tensor, modal2: torch.Tensor, training: bool = True,
                       separate: bool = False) -> torch.Tensor:
    """fuseunetsa.forward (fuseunet.py:138-208); separate=True: fuseunetsaseparate.forward (:255-325), whose modal-1
    encoder max-pools its OWN gated output instead of the fused concat."""
    fused = []
    y, x = modal1, modal2
    for lvl in range(1, 6):
        if lvl > 1:
            y = F.max_pool2d(y if separate else fused[-1], 2, 2)
            x = F.max_pool2d(x, 2, 2)
        y = _basic_block(p, f"modal1_downblock{lvl}.block", y, training)
        y = _spatial_attention(p, f"modal1_sa{lvl}", y, training) * y
        x = _basic_block(p, f"modal2_downblock{lvl}.block", x, training)
        x = _spatial_attention(p, f"modal2_sa{lvl}", x, training) * x
        fused.append(torch.cat((y, x), dim=1))
    return _decoder(p, fused[:4], fused[4], training)


def fuseunetsaseparate_forward(p: Params, modal1: torch.Tensor, modal2: torch.Tensor, training: bool = True) -> torch.Tensor:
    return fuseunetsa_forward(p, modal1, modal2, training, separate=True)


def unetsa_forward(p: Params, x: torch.Tensor, training: bool = True) -> torch.Tensor:
    """UNetsa.forward (UNet.py:190-208)."""
    feats: List[torch.Tensor] = []
    for lvl in range(1, 6):
        if lvl > 1:
            x = F.max_pool2d(x, 2, 2)
        x = _basic_block(p, f"down_block{lvl}.block", x, training)
        x = _spatial_attention(p, f"sa{lvl}", x, training) * x
        feats.append(x)
    return _decoder(p, feats[:4], feats[4], training)


def unet_forward(p: Params, x: torch.Tensor, training: bool = True) -> torch.Tensor:
    """UNet.py:152-165 (max-pool lives inside down blocks 2..5, UNet.py:117-121)."""
    feats: List[torch.Tensor] = []
    for lvl in range(1, 6):
        if lvl > 1:
            x = F.max_pool2d(x, 2, 2)
        x = _basic_block(p, f"down_block{lvl}.block", x, training)
        feats.append(x)
    return _decoder(p, feats[:4], feats[4], training)


# --------------------------------------------------------------------------------------
# losses / metrics
# --------------------------------------------------------------------------------------
def ce_per_pixel(logits: torch.Tensor, targets: torch.Tensor,
                 class_weight: Optional[Sequence[float]] = None, ignore_index: int = 255) -> torch.Tensor:
    """loss2d.py:5-13 with reduction='none' -> [N,H,W]."""
    w = None if class_weight is None else torch.as_tensor(class_weight, dtype=logits.dtype)
    return F.cross_entropy(logits, targets, weight=w, reduction="none", ignore_index=ignore_index)


def dice_per_image(logits: torch.Tensor, targets: torch.Tensor, smooth: float = 1.0) -> torch.Tensor:
    """MulticlassDiceLoss (loss2d.py:95-107, 3-D target branch) -> DiceLoss (loss2d.py:42-58)."""
    n = targets.size(0)
    prob = F.softmax(logits, dim=1)[:, 1]
    iflat = prob.reshape(n, -1).float()
    tflat = targets.reshape(n, -1).float()
    inter = (iflat * tflat).sum(1)
    return 1.0 - (2.0 * inter + smooth) / (iflat.sum(1) + tflat.sum(1) + smooth)


def ce_dice_per_image(logits: torch.Tensor, targets: torch.Tensor,
                      cedice_w: Sequence[float] = (1.0, 1.0),
                      ce_class_w: Optional[Sequence[float]] = (1.0, 1.0)) -> torch.Tensor:
    """CEMDiceLossImage.forward, loss2d.py:146-154 -> [N]."""
    ce = ce_per_pixel(logits, targets, ce_class_w).mean(dim=[1, 2])
    return ce * cedice_w[0] + dice_per_image(logits, targets) * cedice_w[1]


def ce_dice_mean(logits: torch.Tensor, targets: torch.Tensor,
                 cedice_w: Sequence[float] = (1.0, 1.0),
                 ce_class_w: Optional[Sequence[float]] = (1.0, 1.0)) -> torch.Tensor:
    """CEMDiceLoss.forward with reduction='mean', loss2d.py:128-135 -> scalar."""
    w = None if ce_class_w is None else torch.as_tensor(ce_class_w, dtype=logits.dtype)
    ce = F.cross_entropy(logits, targets, weight=w, reduction="mean", ignore_index=255)
    dice = dice_per_image(logits, targets).sum() / targets.size(0)
    return ce * cedice_w[0] + dice * cedice_w[1]


def dice_loss_mean(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """DiceLoss()(logits, targets) with 4-D input and reduction='mean', loss2d.py:42-54."""
    return dice_per_image(logits, targets).sum() / targets.size(0)


def weighted_mse_mean(logits: torch.Tensor, pseudo: torch.Tensor, weightmap: torch.Tensor) -> torch.Tensor:
    """weightmap * MulticlassMSELoss('none')(logits, pseudo) then .mean()
    (loss2d.py:109-117; trainchaos_proposed_30cases1labeled.py:311-313)."""
    return (weightmap * (F.softmax(logits, dim=1) - pseudo) ** 2).mean()


def dice_fn(logits: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5) -> torch.Tensor:
    """metrics2d.py:8-29 -- returns the batch SUM of per-image Dice."""
    prob = F.softmax(logits, dim=1)[:, 1]
    pred = (prob >= threshold).float()
    total = torch.zeros(())
    for pr, tg in zip(pred, targets):
        i = pr.reshape(-1)
        t = tg.reshape(-1).float()
        if t.sum() == 0:
            total = total + (1.0 if i.sum() == 0 else 0.0)
        else:
            total = total + (2.0 * (i * t).sum()) / (i.sum() + t.sum())
    return total


def sharpen(mask: torch.Tensor, temperature: float, flavour: str = "chaos") -> torch.Tensor:
    """chaos: pow(mask, T) (trainchaos_proposed...:97-101); kidney: pow(mask, 1/T) (trainkidney...:113-117)."""
    e = temperature if flavour == "chaos" else 1.0 / temperature
    m = torch.pow(mask, e)
    return m / m.sum(dim=1).unsqueeze(dim=1)


def pseudo_label(aug_logits: Sequence[torch.Tensor], temperature: float = 1.0,
                 flavour: str = "chaos") -> Tuple[torch.Tensor, torch.Tensor]:
    """softmax of every augmented output, running sum, /n, sharpen, weightmap = 1 - 4 q0 q1
    (trainchaos_proposed_30cases1labeled.py:274-292).  Returns (q [N,2,H,W], w [N,1,H,W])."""
    acc = None
    for lg in aug_logits:
        sm = F.softmax(lg, dim=1)
        acc = sm if acc is None else acc + sm
    q = sharpen(acc / float(len(aug_logits)), temperature, flavour)
    w = (1.0 - 4.0 * q[:, 0] * q[:, 1]).unsqueeze(dim=1)
    return q, w


def coteach_losses(out1: torch.Tensor, out2: torch.Tensor, targets1: torch.Tensor, targets2: torch.Tensor,
                   q1: torch.Tensor, w1: torch.Tensor, q2: torch.Tensor, w2: torch.Tensor,
                   rate: float, segcor_weight: Sequence[float] = (1.0, 10.0), n_clean: int = 2,
                   cedice_w: Sequence[float] = (1.0, 1.0), ce_class_w: Sequence[float] = (1.0, 1.0)):
    """trainchaos_proposed_30cases1labeled.py:303-321.  net-1 is trained on net-2's ordering and
    vice versa; note net-1's outputs are scored against targets2 (and net-2's against targets1).
    Returns dict(loss1, loss2, indx1, indx2, pre1, pre2)."""
    crit = lambda o, t: ce_dice_per_image(o, t, cedice_w, ce_class_w)
    pre1 = crit(out1, targets2)
    pre2 = crit(out2, targets1)
    _, indx1 = pre1.sort()
    _, indx2 = pre2.sort()
    k = n_clean
    l1_seg1 = crit(out1[indx2[0:k]], targets2[indx2[0:k]]).mean()
    l2_seg1 = crit(out2[indx1[0:k]], targets1[indx1[0:k]]).mean()
    l1_seg2 = crit(out1[indx2[k:]], targets2[indx2[k:]]).mean()
    l2_seg2 = crit(out2[indx1[k:]], targets1[indx1[k:]]).mean()
    l1_cor = weighted_mse_mean(out1[indx2[k:]], q2[indx2[k:]], w2[indx2[k:]])
    l2_cor = weighted_mse_mean(out2[indx1[k:]], q1[indx1[k:]], w1[indx1[k:]])
    loss1 = segcor_weight[0] * (l1_seg1 + (1.0 - rate) * l1_seg2) + segcor_weight[1] * rate * l1_cor
    loss2 = segcor_weight[0] * (l2_seg1 + (1.0 - rate) * l2_seg2) + segcor_weight[1] * rate * l2_cor
    return dict(loss1=loss1, loss2=loss2, indx1=indx1, indx2=indx2, pre1=pre1, pre2=pre2)


# --------------------------------------------------------------------------------------
# the restated AIDE step (teacher-forcing friendly: explicit state in, explicit state out)
# --------------------------------------------------------------------------------------
def adam_amsgrad_step(params: Params, grads: Dict[str, torch.Tensor], state: Dict[str, Dict[str, torch.Tensor]],
                      step: int, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8) -> None:
    """torch.optim.Adam(lr, amsgrad=True) single-tensor formula (weight_decay 0), in place.
    ``step`` is the 1-based step count after this update."""
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k, g in grads.items():
        st = state.setdefault(k, dict(exp_avg=torch.zeros_like(g), exp_avg_sq=torch.zeros_like(g),
                                      max_exp_avg_sq=torch.zeros_like(g)))
        st["exp_avg"].lerp_(g, 1.0 - b1)
        st["exp_avg_sq"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        torch.maximum(st["max_exp_avg_sq"], st["exp_avg_sq"], out=st["max_exp_avg_sq"])
        denom = (st["max_exp_avg_sq"].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].data.addcdiv_(st["exp_avg"], denom, value=-(lr / bc1))


def aide_step(fwd, p1: Params, p2: Params, inputs: Tuple[torch.Tensor, ...],
              aug_inputs: Sequence[Tuple[torch.Tensor, ...]], targets1: torch.Tensor, targets2: torch.Tensor,
              rate: float, temperature: float = 1.0, flavour: str = "chaos", n_clean: int = 2,
              segcor_weight: Sequence[float] = (1.0, 10.0), augset: Optional[Dict] = None):
    """One AIDE iteration up to (not including) the optimiser step.

    chaos flavour: augmented forwards in train mode (BN buffers updated 4 extra times), sharpen pow(T)
    (trainchaos_proposed_30cases1labeled.py:263-325); kidney flavour: augmented forwards in eval mode,
    sharpen pow(1/T) (trainkidney_proposed_mask1.py:267-333).  Reverse-augmentation is the identity
    (degree 0, no flip) as in the synthetic benchmark.  ``p1``/``p2`` must have requires_grad params.
    Returns dict with logits, losses, index sets, grads (per net), dice_fn sums.
    """
    aug_train = flavour == "chaos"
    with torch.no_grad():
        aug1 = [fwd(p1, *a, training=aug_train) for a in aug_inputs]
        aug2 = [fwd(p2, *a, training=aug_train) for a in aug_inputs]
    if augset is not None:
        aug1 = reverseaug_pil(augset, aug1, aug1[0].shape[1])
        aug2 = reverseaug_pil(augset, aug2, aug2[0].shape[1])
    q1, w1 = pseudo_label(aug1, temperature, flavour)
    q2, w2 = pseudo_label(aug2, temperature, flavour)
    out1 = fwd(p1, *inputs, training=True)
    out2 = fwd(p2, *inputs, training=True)
    res = coteach_losses(out1, out2, targets1, targets2, q1, w1, q2, w2, rate, segcor_weight, n_clean)
    names1 = [k for k in p1 if not is_buffer(k)]
    g1 = torch.autograd.grad(res["loss1"], [p1[k] for k in names1], retain_graph=True)
    g2 = torch.autograd.grad(res["loss2"], [p2[k] for k in names1])
    res.update(out1=out1.detach(), out2=out2.detach(), q1=q1, w1=w1, q2=q2, w2=w2,
               grads1=dict(zip(names1, g1)), grads2=dict(zip(names1, g2)),
               dice1=dice_fn(out1.detach(), targets2), dice2=dice_fn(out2.detach(), targets1))
    return res


# --------------------------------------------------------------------------------------
# synthetic workload
# --------------------------------------------------------------------------------------
def synthetic_batch(b: int, h: int, w: Optional[int] = None, seed: int = 1234, n_modal: int = 2, n_aug: int = 0):
    """images randn(B,3,H,W); targets Bernoulli(0.08) int64; two independent target draws."""
    w = h if w is None else w
    g = torch.Generator().manual_seed(seed)
    imgs = tuple(torch.randn(b, 3, h, w, generator=g) for _ in range(n_modal))
    t1 = (torch.rand(b, h, w, generator=g) < 0.08).long()
    t2 = (torch.rand(b, h, w, generator=g) < 0.08).long()
    augs = [tuple(torch.randn(b, 3, h, w, generator=g) for _ in range(n_modal)) for _ in range(n_aug)]
    return imgs, t1, t2, augs


# --------------------------------------------------------------------------------------
# reverse augmentation of the augmented forwards' outputs ("next" row f2 of SURVEY.md section 8)
# --------------------------------------------------------------------------------------
def reverseaug_pil(augset: Dict, augoutput: List[torch.Tensor], classno: int) -> List[torch.Tensor]:
    """train_files/trainchaos_proposed_30cases1labeled.py:81-95, verbatim semantics: every (sample, view, class) plane goes
    through PIL (mode 'F'): optional FLIP_LEFT_RIGHT, then rotate(-degree, BILINEAR) about the image centre, fill 0.
    Pillow is the reference's third-party dependency for this step (unpinned in requirements.txt; 12.2.0 here)."""
    import numpy as np
    from PIL import Image
    out = [t.clone() for t in augoutput]
    for batch_idx in range(len(augset["augno"])):
        for aug_idx in range(int(augset["augno"][batch_idx])):
            imgflip = augset["hflip{}".format(aug_idx + 1)][batch_idx]
            rotation = 0 - augset["degree{}".format(aug_idx + 1)][batch_idx]
            for classidx in range(classno):
                mask = out[aug_idx][batch_idx, classidx, :, :].cpu().numpy()
                mask = Image.fromarray(mask, mode="F")
                if imgflip:
                    mask = mask.transpose(Image.FLIP_LEFT_RIGHT)
                mask = mask.rotate(float(rotation), Image.BILINEAR)
                out[aug_idx][batch_idx, classidx, :, :] = torch.from_numpy(np.array(mask))
    return out


def rotate_matrix(angle_deg: float, w: int, h: int):
    """PIL.Image.Image.rotate's inverse affine matrix (output pixel -> source position), PIL/Image.py `rotate`:
    returns (mode, matrix) with mode 0 = affine, 1 = copy (angle % 360 == 0), 2 = ROTATE_180, 3 = ROTATE_90,
    4 = ROTATE_270 (PIL's exact fast paths)."""
    angle = angle_deg % 360.0
    if angle == 0:
        return 1, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    if angle == 180:
        return 2, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    if angle in (90, 270) and w == h:
        return (3 if angle == 90 else 4), [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return 0, m


def reverseaug_plane(src, hflip: bool, angle_deg: float):
    """numpy restatement of what PIL does to one float32 plane (Pillow src/libImaging/Geometry.c: affine_transform +
    bilinear_filter32F, double arithmetic, clamped neighbours, out-of-image source -> 0).  Checked bit for bit against
    reverseaug_pil in tests/test_oracle_golden.py."""
    import numpy as np
    h, w = src.shape
    s = src[:, ::-1] if hflip else src
    mode, m = rotate_matrix(angle_deg, w, h)
    if mode == 1:
        return s.copy()
    if mode == 2:
        return s[::-1, ::-1].copy()
    if mode == 3:                                  # Transpose.ROTATE_90: counter-clockwise
        return np.ascontiguousarray(np.rot90(s, 1))
    if mode == 4:
        return np.ascontiguousarray(np.rot90(s, 3))
    ys, xs = np.mgrid[0:h, 0:w]
    xin = m[0] * (xs + 0.5) + m[1] * (ys + 0.5) + m[2]
    yin = m[3] * (xs + 0.5) + m[4] * (ys + 0.5) + m[5]
    inside = (xin >= 0.0) & (xin < w) & (yin >= 0.0) & (yin < h)
    xin = xin - 0.5
    yin = yin - 0.5
    x = np.floor(xin).astype(np.int64)
    y = np.floor(yin).astype(np.int64)
    dx, dy = xin - x, yin - y
    x0, x1 = np.clip(x, 0, w - 1), np.clip(x + 1, 0, w - 1)
    yc = np.clip(y, 0, h - 1)
    # BILINEAR(v, a, b, d): v = a + (b - a) * d with a, b FLOAT32 -> the difference is rounded to float32 first
    f64 = np.float64
    v1 = s[yc, x0].astype(f64) + (s[yc, x1] - s[yc, x0]).astype(f64) * dx
    has2 = (y + 1 >= 0) & (y + 1 < h)
    y2 = np.clip(y + 1, 0, h - 1)
    v2 = np.where(has2, s[y2, x0].astype(f64) + (s[y2, x1] - s[y2, x0]).astype(f64) * dx, v1)
    out = (v1 + (v2 - v1) * dy).astype(np.float32)
    return np.where(inside, out, np.float32(0.0))


# --------------------------------------------------------------------------------------
# forward augmentation of the input images ("next" row f2 of SURVEY.md section 8, second half)
# --------------------------------------------------------------------------------------
def forward_aug_pil(img_u8, degree: float, hflip: bool, mean: torch.Tensor, std: torch.Tensor) -> torch.Tensor:
    """One view exactly as the reference's loader makes it (datasetchaos_proposed/transform.py; Compose order of
    train_files/trainchaos_proposed_30cases1labeled.py:191-197): PIL rotate(degree, BILINEAR) on the uint8 RGB image
    (:81-106), FLIP_LEFT_RIGHT (:16-34), ToTensor (:108-131), Normalize with the given per-channel mean / std
    (:134-170).  img_u8: numpy [H,W,3] uint8.  Returns fp32 [3,H,W]."""
    import numpy as np
    from PIL import Image
    im = Image.fromarray(img_u8).rotate(degree, Image.BILINEAR)
    if hflip:
        im = im.transpose(Image.FLIP_LEFT_RIGHT)
    t = torch.from_numpy(np.array(im).transpose(2, 0, 1)).float() / 255.0
    return t.sub(mean.reshape(3, 1, 1)).div(std.reshape(3, 1, 1))


def rotate_u8(src, angle_deg: float):
    """numpy restatement of PIL's Image.rotate(angle, BILINEAR) on an RGB uint8 image [H,W,3] (Pillow
    src/libImaging/Geometry.c: affine_transform + bilinear_filter32RGB -- neighbours clamped, interpolation in double,
    result truncated to uint8, outside -> 0).  Checked bit for bit against PIL in tests/test_oracle_golden.py."""
    import numpy as np
    h, w, _ = src.shape
    mode, m = rotate_matrix(angle_deg, w, h)
    if mode == 1:
        return src.copy()
    if mode == 2:
        return src[::-1, ::-1].copy()
    if mode == 3:
        return np.ascontiguousarray(np.rot90(src, 1))
    if mode == 4:
        return np.ascontiguousarray(np.rot90(src, 3))
    ys, xs = np.mgrid[0:h, 0:w]
    xin = m[0] * (xs + 0.5) + m[1] * (ys + 0.5) + m[2]
    yin = m[3] * (xs + 0.5) + m[4] * (ys + 0.5) + m[5]
    inside = (xin >= 0.0) & (xin < w) & (yin >= 0.0) & (yin < h)
    xin, yin = xin - 0.5, yin - 0.5
    x, y = np.floor(xin).astype(np.int64), np.floor(yin).astype(np.int64)
    dx, dy = xin - x, yin - y
    x0, x1 = np.clip(x, 0, w - 1), np.clip(x + 1, 0, w - 1)
    yc, y2 = np.clip(y, 0, h - 1), np.clip(y + 1, 0, h - 1)
    has2 = (y + 1 >= 0) & (y + 1 < h)
    s = src.astype(np.float64)
    out = np.zeros_like(src)
    for b in range(3):
        v1 = s[yc, x0, b] + (s[yc, x1, b] - s[yc, x0, b]) * dx
        v2 = np.where(has2, s[y2, x0, b] + (s[y2, x1, b] - s[y2, x0, b]) * dx, v1)
        out[..., b] = np.where(inside, (v1 + (v2 - v1) * dy).astype(np.uint8), 0)
    return out
