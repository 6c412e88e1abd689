"""Losses / metrics of the AIDE hot path on the fused CUDA reductions (csrc/loss.cu).

Drop-in surface of the reference's utils package:
  utils/loss2d.py      CrossEntropyLoss2d :5-13, DiceLoss :35-61, Dice_Loss :63-85, MulticlassDiceLoss :87-107,
                       MulticlassMSELoss :109-117, CEMDiceLoss :119-135, CEMDiceLossImage :137-154, CEDiceLoss :156-171
  utils/metrics2d.py   Dice_fn :8-29
plus the fused form of the inline co-teaching step
  train_files/trainchaos_proposed_30cases1labeled.py:274-292 (pseudo label) and :303-321 (selection + loss).

One kernel pass over (logits, targets) yields eight per-image fp64 sums (CE, class-weight mass, Dice
intersection / prob mass / target mass, thresholded counts, weighted-MSE); every loss above is a
closed form of those, and their gradient w.r.t. the logits is a second elementwise kernel driven by
per-image coefficients.  Two-class segmentation only (num_classes == 2), as in every reference script.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import call, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _floats(w, default=(1.0, 1.0)) -> Tuple[float, float]:
    if w is None:
        return default
    if torch.is_tensor(w):
        w = w.detach().cpu().tolist()
    w = [float(x) for x in w]
    if len(w) != 2:
        raise NotImplementedError("aide_b200 losses support two classes (background / object) only")
    return w[0], w[1]


def _check(logits: torch.Tensor, targets: torch.Tensor):
    if not logits.is_cuda:
        raise RuntimeError("aide_b200 losses run on CUDA only (there is no CPU fallback)")
    if logits.dim() != 4 or logits.shape[1] != 2:
        raise NotImplementedError(f"fused losses need logits [N,2,H,W]; got {tuple(logits.shape)}")
    if targets.dim() == 4:                       # one-hot / soft targets: reference takes the argmax (loss2d.py:11-12)
        targets = torch.argmax(targets.float(), dim=1)
    if targets.shape != (logits.shape[0],) + tuple(logits.shape[2:]):
        raise ValueError(f"targets {tuple(targets.shape)} do not match logits {tuple(logits.shape)}")
    if targets.dtype != torch.int64:
        targets = targets.long()
    return logits.contiguous().float(), targets.contiguous()


def image_sums(logits: torch.Tensor, targets: torch.Tensor, q: Optional[torch.Tensor] = None,
               wm: Optional[torch.Tensor] = None, class_w=(1.0, 1.0), ignore_index: int = 255,
               threshold: float = 0.5) -> torch.Tensor:
    """[N,8] fp64 per-image sums (see include/aide_b200.h: aide_loss_sums)."""
    N, _, H, W = logits.shape
    sums = torch.empty((N, 8), dtype=torch.float64, device=logits.device)
    scratch = torch.empty((N * lib.aide_loss_blocks(H, W), 8), dtype=torch.float64, device=logits.device)
    call("aide_loss_sums", logits.data_ptr(), targets.data_ptr(), q.data_ptr() if q is not None else None,
         wm.data_ptr() if wm is not None else None, N, H, W, class_w[0], class_w[1], ignore_index, threshold,
         sums.data_ptr(), scratch.data_ptr(), _stream())
    return sums


def image_finalize(sums: torch.Tensor, H: int, W: int, w_ce: float, w_dice: float, smooth: float,
                   want_loss=True, want_dice=False, want_dice_fn=False):
    N = sums.shape[0]
    dev = sums.device
    loss = torch.empty(N, dtype=torch.float32, device=dev) if want_loss else None
    dice = torch.empty(N, dtype=torch.float32, device=dev) if want_dice else None
    dfn = torch.empty((), dtype=torch.float32, device=dev) if want_dice_fn else None
    call("aide_loss_image_finalize", sums.data_ptr(), N, H, W, w_ce, w_dice, smooth,
         loss.data_ptr() if want_loss else None, dice.data_ptr() if want_dice else None,
         dfn.data_ptr() if want_dice_fn else None, _stream())
    return loss, dice, dfn


def loss_backward(logits, targets, sums, a_ce, a_dice, a_mse=None, q=None, wm=None, class_w=(1.0, 1.0),
                  ignore_index=255, smooth=1.0) -> torch.Tensor:
    N, _, H, W = logits.shape
    d = torch.empty_like(logits)
    p = lambda t: t.data_ptr() if t is not None else None
    call("aide_loss_bwd", logits.data_ptr(), targets.data_ptr(), p(q), p(wm), sums.data_ptr(), p(a_ce), p(a_dice),
         p(a_mse), N, H, W, class_w[0], class_w[1], ignore_index, smooth, d.data_ptr(), _stream())
    return d


class _PerImageLoss(torch.autograd.Function):
    """per-image  w_ce * mean_hw(CE * wc[t]) + w_dice * dice  -> [N]   (CEMDiceLossImage.forward)."""

    @staticmethod
    def forward(ctx, logits, targets, w_ce, w_dice, class_w, smooth, ignore_index):
        N, _, H, W = logits.shape
        sums = image_sums(logits, targets, class_w=class_w, ignore_index=ignore_index)
        loss, _, _ = image_finalize(sums, H, W, w_ce, w_dice, smooth)
        ctx.save_for_backward(logits, targets, sums)
        ctx.cfg = (w_ce, w_dice, class_w, smooth, ignore_index, H * W)
        return loss

    @staticmethod
    def backward(ctx, go):
        logits, targets, sums = ctx.saved_tensors
        w_ce, w_dice, class_w, smooth, ignore_index, hw = ctx.cfg
        go = go.contiguous().float()
        d = loss_backward(logits, targets, sums, go * (w_ce / hw), go * w_dice, class_w=class_w,
                          ignore_index=ignore_index, smooth=smooth)
        return d, None, None, None, None, None, None


class _BatchLoss(torch.autograd.Function):
    """scalar  w_ce * CE_reduced + w_dice * dice_reduced  with reduction in {'mean','sum'}."""

    @staticmethod
    def forward(ctx, logits, targets, w_ce, w_dice, class_w, smooth, ignore_index, reduction):
        N, _, H, W = logits.shape
        sums = image_sums(logits, targets, class_w=class_w, ignore_index=ignore_index)
        _, dice, _ = image_finalize(sums, H, W, 0.0, 1.0, smooth, want_loss=False, want_dice=True)
        ce_tot = sums[:, 0].sum()
        if reduction == "mean":
            wsum = sums[:, 1].sum()
            ce = (ce_tot / wsum).float()
            dl = dice.sum() / N
            ctx.coef = (w_ce, w_dice / N)
        else:
            wsum = None
            ce = ce_tot.float()
            dl = dice.sum()
            ctx.coef = (w_ce, w_dice)
        ctx.save_for_backward(logits, targets, sums, wsum if wsum is not None else sums.new_ones(()))
        ctx.cfg = (class_w, smooth, ignore_index, reduction)
        return ce * w_ce + dl * w_dice

    @staticmethod
    def backward(ctx, go):
        logits, targets, sums, wsum = ctx.saved_tensors
        class_w, smooth, ignore_index, reduction = ctx.cfg
        c_ce, c_dice = ctx.coef
        N = logits.shape[0]
        go = go.float()
        a_ce = (go * c_ce / wsum.float()).expand(N).contiguous() if reduction == "mean" \
            else (go * c_ce).expand(N).contiguous()
        a_dice = (go * c_dice).expand(N).contiguous()
        d = loss_backward(logits, targets, sums, a_ce, a_dice, class_w=class_w, ignore_index=ignore_index,
                          smooth=smooth)
        return d, None, None, None, None, None, None, None


class _PixelLossMap(torch.autograd.Function):
    """[N,H,W] map: (mode & 1) class-weighted cross-entropy + (mode & 2) bidirectional KL between two logit tensors."""

    @staticmethod
    def forward(ctx, logits1, logits2, targets, class_w, ignore_index, mode):
        N, _, H, W = logits1.shape
        out = torch.empty((N, H, W), dtype=torch.float32, device=logits1.device)
        call("aide_pixel_loss_fwd", logits1.data_ptr(), logits2.data_ptr() if logits2 is not None else None,
             targets.data_ptr(), N, H, W, class_w[0], class_w[1], ignore_index, mode, out.data_ptr(), _stream())
        ctx.save_for_backward(logits1, logits2 if logits2 is not None else logits1, targets)
        ctx.cfg = (class_w, ignore_index, mode, logits2 is not None)
        return out

    @staticmethod
    def backward(ctx, go):
        l1, l2, targets = ctx.saved_tensors
        class_w, ignore_index, mode, has2 = ctx.cfg
        N, _, H, W = l1.shape
        go = go.contiguous().float()
        d1 = torch.empty_like(l1)
        d2 = torch.empty_like(l2) if (has2 and ctx.needs_input_grad[1]) else None
        call("aide_pixel_loss_bwd", l1.data_ptr(), l2.data_ptr() if has2 else None, targets.data_ptr(), go.data_ptr(), N, H,
             W, class_w[0], class_w[1], ignore_index, mode, d1.data_ptr(), d2.data_ptr() if d2 is not None else None,
             _stream())
        return d1, d2, None, None, None, None


def pixel_loss_map(logits1, targets, logits2=None, class_w=(1.0, 1.0), ignore_index=255, ce=True, kl=False):
    """Per-pixel loss map [N,H,W] on the fused kernels (CrossEntropyLoss2d('none'), coteach_loss.py's drop map)."""
    l1, t = _check(logits1, targets)
    l2 = None
    if kl:
        l2, _ = _check(logits2, targets)
    return _PixelLossMap.apply(l1, l2, t, class_w, ignore_index, (1 if ce else 0) | (2 if kl else 0))


class _SoftmaxMSE(torch.autograd.Function):
    """(softmax(z, 1) - target)^2 elementwise, [N,2,H,W] (MulticlassMSELoss, loss2d.py:109-117)."""

    @staticmethod
    def forward(ctx, logits, target):
        N, _, H, W = logits.shape
        out = torch.empty_like(logits)
        call("aide_softmax_mse_fwd", logits.data_ptr(), target.data_ptr(), N, H, W, out.data_ptr(), _stream())
        ctx.save_for_backward(logits, target)
        return out

    @staticmethod
    def backward(ctx, go):
        logits, target = ctx.saved_tensors
        N, _, H, W = logits.shape
        go = go.contiguous().float()
        d = torch.empty_like(logits)
        call("aide_softmax_mse_bwd", logits.data_ptr(), target.data_ptr(), go.data_ptr(), N, H, W, d.data_ptr(), _stream())
        return d, None


class _MaxPoolNCHW(torch.autograd.Function):
    """max_pool2d(kernel = stride = k, ceil_mode=True) on NCHW fp32 (coteach_loss.py:170-178)."""

    @staticmethod
    def forward(ctx, x, kh, kw):
        N, Cc, H, W = x.shape
        OH, OW = -(-H // kh), -(-W // kw)
        y = torch.empty((N, Cc, OH, OW), dtype=torch.float32, device=x.device)
        arg = torch.empty((N, Cc, OH, OW), dtype=torch.int32, device=x.device)
        call("aide_maxpool_nchw_fwd", x.data_ptr(), N * Cc, H, W, kh, kw, y.data_ptr(), arg.data_ptr(), _stream())
        ctx.save_for_backward(arg)
        ctx.cfg = (N * Cc, H, W, kh, kw, x.shape)
        return y

    @staticmethod
    def backward(ctx, gy):
        (arg,) = ctx.saved_tensors
        planes, H, W, kh, kw, shape = ctx.cfg
        gx = torch.empty(shape, dtype=torch.float32, device=gy.device)
        call("aide_maxpool_nchw_bwd", gy.contiguous().float().data_ptr(), arg.data_ptr(), planes, H, W, kh, kw, gx.data_ptr(),
             _stream())
        return gx, None, None


def maxpool_nchw(x: torch.Tensor, kh: int, kw: int) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("aide_b200 runs on CUDA only (there is no CPU fallback)")
    return _MaxPoolNCHW.apply(x.contiguous().float(), int(kh), int(kw))


# -------------------------------------------------------------------------------------------------
# drop-in classes (constructor signatures as in utils/loss2d.py)
# -------------------------------------------------------------------------------------------------
class CrossEntropyLoss2d(nn.Module):
    def __init__(self, weight=None, reduction="mean", ignore_index=255):
        super().__init__()
        self.class_w = _floats(weight)
        self.reduction, self.ignore_index = reduction, ignore_index

    def forward(self, inputs, targets):
        logits, targets = _check(inputs, targets)
        if self.reduction == "none":
            return _PixelLossMap.apply(logits, None, targets, self.class_w, self.ignore_index, 1)     # [N,H,W] map
        return _BatchLoss.apply(logits, targets, 1.0, 0.0, self.class_w, 1.0, self.ignore_index, self.reduction)


class DiceLoss(nn.Module):
    def __init__(self, weight=None, smooth=1.0, reduction="mean"):
        super().__init__()
        self.weight, self.smooth, self.reduction = weight, smooth, reduction

    def forward(self, input, target):
        if input.dim() <= 3:
            # probabilities given directly (loss2d.py:47-48): tensor plumbing, not the fused path
            n = target.size(0)
            i, t = input.reshape(n, -1).float(), target.reshape(n, -1).float()
            loss = 1.0 - (2.0 * (i * t).sum(1) + self.smooth) / (i.sum(1) + t.sum(1) + self.smooth)
            return {"mean": loss.sum() / n, "sum": loss.sum(), "none": loss}[self.reduction]
        logits, target = _check(input, target)
        if self.reduction == "none":
            return _PerImageLoss.apply(logits, target, 0.0, 1.0, (1.0, 1.0), self.smooth, 255)
        return _BatchLoss.apply(logits, target, 0.0, 1.0, (1.0, 1.0), self.smooth, 255, self.reduction)


class Dice_Loss(DiceLoss):
    def __init__(self, smooth=1.0, reduction="mean"):
        super().__init__(None, smooth, reduction)


class MulticlassDiceLoss(nn.Module):
    def __init__(self, weight=None, smooth=1.0, reduction="mean"):
        super().__init__()
        self.weight, self.smooth, self.reduction = weight, smooth, reduction
        self.dice = DiceLoss(smooth=smooth, reduction=reduction)

    def forward(self, input, target):
        if target.dim() > 3:      # per-class one-hot targets (loss2d.py:98-104): tensor plumbing
            prob = F.softmax(input, dim=1)
            total = 0
            for i in range(target.shape[1]):
                d = self.dice(prob[:, i], target[:, i])
                if self.weight is not None:
                    d = d * self.weight[i]
                total = total + d
            return total
        return self.dice(input, target)   # class weights are NOT applied in this branch (loss2d.py:106)


class MulticlassMSELoss(nn.Module):
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, input, target):
        # one kernel for softmax -> difference -> square (and one for its gradient); the AIDE step's weighted form is
        # fused further in coteach_step() below
        if not input.is_cuda:
            raise RuntimeError("aide_b200 losses run on CUDA only (there is no CPU fallback)")
        if input.dim() != 4 or input.shape[1] != 2 or target.shape != input.shape:
            raise NotImplementedError("MulticlassMSELoss needs logits and targets [N,2,H,W] (two classes)")
        m = _SoftmaxMSE.apply(input.contiguous().float(), target.detach().contiguous().float())
        return {"none": m, "mean": m.mean(), "sum": m.sum()}[self.reduction]


class CEMDiceLoss(nn.Module):
    def __init__(self, cediceweight=None, ceclassweight=None, diceclassweight=None, reduction="mean"):
        super().__init__()
        self.w = _floats(cediceweight)
        self.class_w = _floats(ceclassweight)
        self.reduction = reduction

    def forward(self, inputs, targets):
        logits, targets = _check(inputs, targets)
        if self.reduction == "none":
            raise NotImplementedError("CEMDiceLoss(reduction='none') adds a [N,H,W] map to a [N] vector in the "
                                      "reference and cannot broadcast; use CEMDiceLossImage")
        return _BatchLoss.apply(logits, targets, self.w[0], self.w[1], self.class_w, 1.0, 255, self.reduction)


class CEDiceLoss(CEMDiceLoss):
    def __init__(self, cediceweight=None, classweight=None, reduction="mean"):
        super().__init__(cediceweight, classweight, None, reduction)


class CEMDiceLossImage(nn.Module):
    def __init__(self, cediceweight=None, ceclassweight=None, diceclassweight=None, reduction="mean"):
        super().__init__()
        self.w = _floats(cediceweight)
        self.class_w = _floats(ceclassweight)

    def forward(self, inputs, targets):
        logits, targets = _check(inputs, targets)
        return _PerImageLoss.apply(logits, targets, self.w[0], self.w[1], self.class_w, 1.0, 255)


def Dice_fn(inputs, targets, threshold=0.5):
    """metrics2d.py:8-29: batch SUM of per-image Dice of the thresholded prediction (no host sync)."""
    logits, targets = _check(inputs.detach(), targets)
    N, _, H, W = logits.shape
    sums = image_sums(logits, targets, threshold=threshold)
    _, _, dfn = image_finalize(sums, H, W, 0.0, 0.0, 1.0, want_loss=False, want_dice_fn=True)
    return dfn


def predict_mask(logits: torch.Tensor) -> torch.Tensor:
    """uint8 [N,H,W] hard mask = argmax(softmax(logits, 1), 1) in one kernel (the per-slice mask emission of the
    pseudo-label rewrite and evaluation passes, trainchaos_proposed_30cases1labeled.py:407-409)."""
    if not logits.is_cuda:
        raise RuntimeError("aide_b200 runs on CUDA only (there is no CPU fallback)")
    x = logits.detach().contiguous().float()
    N, K, H, W = x.shape
    mask = torch.empty((N, H, W), dtype=torch.uint8, device=x.device)
    call("aide_argmax_mask", x.data_ptr(), mask.data_ptr(), N, K, H, W, _stream())
    return mask


# -------------------------------------------------------------------------------------------------
# pseudo labels + fused co-teaching step
# -------------------------------------------------------------------------------------------------
def pseudo_label(aug_logits: Sequence[torch.Tensor], temperature: float = 1.0, flavour: str = "chaos"):
    """softmax -> mean over augmentations -> sharpen -> weight map, one kernel
    (trainchaos_proposed_30cases1labeled.py:274-292; sharpen :97-101, kidney flavour pow(1/T))."""
    a = [t.detach().contiguous().float() for t in aug_logits]
    N, K, H, W = a[0].shape
    if K != 2 or not a[0].is_cuda:
        raise NotImplementedError("pseudo_label needs CUDA logits [N,2,H,W]")
    q = torch.empty_like(a[0])
    wm = torch.empty((N, 1, H, W), dtype=torch.float32, device=a[0].device)
    ptrs = (C.c_void_p * len(a))(*[t.data_ptr() for t in a])
    expo = float(temperature) if flavour == "chaos" else 1.0 / float(temperature)
    call("aide_pseudo_label", ptrs, len(a), N, H, W, expo, q.data_ptr(), wm.data_ptr(), _stream())
    return q, wm


class _CoteachLoss(torch.autograd.Function):
    """loss of ONE net given the OTHER net's per-image pre-loss (selection) -- fused forward + backward."""

    @staticmethod
    def forward(ctx, logits, targets, q_other, wm_other, pre_other, sums, loss_img, cfg):
        n_clean, rate, seg_w, cor_w, w_ce, w_dice, class_w = cfg
        N, _, H, W = logits.shape
        dev = logits.device
        idx = torch.empty(N, dtype=torch.int64, device=dev)
        coef = torch.empty((3, N), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        call("aide_coteach_select", pre_other.data_ptr(), loss_img.data_ptr(), sums.data_ptr(), N, H, W, n_clean,
             rate, seg_w, cor_w, w_ce, w_dice, idx.data_ptr(), coef[0].data_ptr(), coef[1].data_ptr(),
             coef[2].data_ptr(), loss.data_ptr(), _stream())
        ctx.save_for_backward(logits, targets, q_other, wm_other, sums, coef)
        ctx.class_w = class_w
        ctx.mark_non_differentiable(idx)
        return loss, idx

    @staticmethod
    def backward(ctx, go, _gidx):
        logits, targets, q, wm, sums, coef = ctx.saved_tensors
        c = coef * go.float()
        d = loss_backward(logits, targets, sums, c[0].contiguous(), c[1].contiguous(), c[2].contiguous(), q, wm,
                          class_w=ctx.class_w)
        return d, None, None, None, None, None, None, None


def coteach_step(out1, out2, targets1, targets2, q1, w1, q2, w2, rate: float,
                 segcor_weight=(1.0, 10.0), n_clean: int = 2, cedice_w=(1.0, 1.0), ce_class_w=(1.0, 1.0)):
    """Fused equivalent of trainchaos_proposed_30cases1labeled.py:303-321.

    net-1's outputs are scored against targets2 and trained on net-2's small-loss ordering (and vice
    versa).  No index-gather copies: the selection becomes per-image coefficients of one backward kernel.
    Returns dict(loss1, loss2, indx1, indx2, pre1, pre2, dice1, dice2) -- all device tensors, no host sync.
    """
    o1, t2 = _check(out1, targets2)
    o2, t1 = _check(out2, targets1)
    N, _, H, W = o1.shape
    cw = _floats(ce_class_w)
    s1 = image_sums(o1.detach(), t2, q2, w2, class_w=cw)     # net-1 vs targets2, consistency vs net-2's pseudo label
    s2 = image_sums(o2.detach(), t1, q1, w1, class_w=cw)
    pre1, _, d1 = image_finalize(s1, H, W, cedice_w[0], cedice_w[1], 1.0, want_dice_fn=True)
    pre2, _, d2 = image_finalize(s2, H, W, cedice_w[0], cedice_w[1], 1.0, want_dice_fn=True)
    cfg = (n_clean, float(rate), float(segcor_weight[0]), float(segcor_weight[1]), float(cedice_w[0]),
           float(cedice_w[1]), cw)
    loss1, indx2 = _CoteachLoss.apply(o1, t2, q2, w2, pre2, s1, pre1, cfg)   # ordering of net-2's losses
    loss2, indx1 = _CoteachLoss.apply(o2, t1, q1, w1, pre1, s2, pre2, cfg)
    return dict(loss1=loss1, loss2=loss2, indx1=indx1, indx2=indx2, pre1=pre1, pre2=pre2, dice1=d1, dice2=d2)
