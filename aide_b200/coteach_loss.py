"""API-compatible co-teaching losses (reference: utils/coteach_loss.py:94-254).

No reference script calls these classes; the co-teaching that actually runs is the
inline step fused in aide_b200.losses.coteach_step.  They are provided for import compatibility:
the image-level selection reuses the fused per-image CE+Dice kernel, the per-pixel / per-region maps
(cross-entropy, bidirectional KL, max-pool of logits and targets) run on aide_pixel_loss_* / aide_maxpool_nchw_*; only the
index bookkeeping on the small per-image / per-region vectors (argsort, gather, mean) is tensor plumbing.  As in the reference only ``reduction='none'`` is meaningful (the default 'mean'
raises there: torch.mean(scalar, dim=[1,2]), coteach_loss.py:102).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .losses import _PerImageLoss, _check, maxpool_nchw, pixel_loss_map


def _per_image(logits, targets, weight):
    lg, tg = _check(logits, targets)
    return _PerImageLoss.apply(lg, tg, float(weight), 1.0, (1.0, 1.0), 1.0, 255)


class _Base(nn.Module):
    def __init__(self, weight=1.0, reduction="mean"):
        super().__init__()
        if reduction != "none":
            raise ValueError("only reduction='none' is usable (the reference raises IndexError for 'mean'/'sum')")
        self.weight = weight

    def _split(self, inputs1, inputs2, targets, forget_rate):
        l1 = _per_image(inputs1, targets, self.weight)
        l2 = _per_image(inputs2, targets, self.weight)
        i1, i2 = torch.argsort(l1.detach()), torch.argsort(l2.detach())
        n_rem = int((1 - forget_rate) * l1.shape[0])
        return l1, l2, i1, i2, n_rem


class Coteachingloss_dropimage(_Base):          # coteach_loss.py:94-119
    def forward(self, inputs1, inputs2, targets, forget_rate):
        l1, l2, i1, i2, n = self._split(inputs1, inputs2, targets, forget_rate)
        return l1[i2[:n]].mean(dim=0), l2[i1[:n]].mean(dim=0)


class Coteachingloss_weightimage(_Base):        # coteach_loss.py:121-161
    def forward(self, inputs1, inputs2, targets, forget_rate):
        l1, l2, i1, i2, n = self._split(inputs1, inputs2, targets, forget_rate)
        if l1.shape[0] - n > 0:
            if l1.shape[0] - n != n:
                raise RuntimeError("kept and dropped index sets must have the same size (the reference adds a "
                                   "[num_remember] and a [num_drop] vector, coteach_loss.py:144-147)")
            return (l1[i2[:n]] + 0.1 * l1[i2[n:]]).mean(dim=0), (l2[i1[:n]] + 0.1 * l2[i1[n:]]).mean(dim=0)
        return l1[i2[:n]].mean(dim=0), l2[i1[:n]].mean(dim=0)


class Coteachingloss_dropregionce(nn.Module):   # coteach_loss.py:163-196
    def __init__(self, scale=0.5, reduction="none"):
        super().__init__()
        self.scale = scale

    def forward(self, inputs1, inputs2, targets, forget_rate):
        tw, th = inputs1.shape[2], inputs1.shape[3]
        pw, ph = int(tw * self.scale), int(th * self.scale)
        k = (int(tw / pw), int(th / ph))
        pool = lambda x: maxpool_nchw(x, k[0], k[1])           # aide_maxpool_nchw (ceil_mode, stride = kernel)
        p1, p2 = pool(inputs1), pool(inputs2)
        tp = pool(targets.float().unsqueeze(1)).squeeze(1).long()
        ce = lambda x: pixel_loss_map(x, tp).view(x.shape[0], -1)      # per-region cross-entropy map, one kernel
        l1, l2 = ce(p1), ce(p2)
        n = int((1 - forget_rate) * l1.shape[1])
        i1, i2 = torch.argsort(l1.detach(), dim=1)[:, :n], torch.argsort(l2.detach(), dim=1)[:, :n]
        return torch.gather(l1, 1, i2).mean(), torch.gather(l2, 1, i1).mean()


class Coteachingloss_dropimagedroppixel(_Base):  # coteach_loss.py:198-254
    def forward(self, inputs1, inputs2, targets, forget_rate):
        l1, l2, i1, i2, n = self._split(inputs1, inputs2, targets, forget_rate)
        out1, out2 = l1[i2[:n]].mean(dim=0), l2[i1[:n]].mean(dim=0)
        rem = 1 - forget_rate
        d1, d2 = i1[n:], i2[n:]
        n_rem2 = None
        if len(d1) > 0:
            a, b, t = inputs1[d2], inputs2[d2], targets[d2]
            drop = (pixel_loss_map(a, t, logits2=b, kl=True).view(-1) * t.view(-1).float())     # KL(a,b)+KL(b,a)+CE(a)
            fore = drop[drop > 0]
            order = torch.argsort(fore.detach())
            n_rem2 = int(rem * len(order))
            out1 = out1 + 0.25 * fore[order[:n_rem2]].mean()
        if len(d2) > 0:
            a, b, t = inputs1[d1], inputs2[d1], targets[d1]
            drop = (pixel_loss_map(b, t, logits2=a, kl=True).view(-1) * t.view(-1).float())     # symmetric KL + CE(b)
            fore = drop[drop > 0]
            order = torch.argsort(fore.detach())
            out2 = out2 + 0.25 * fore[order[:n_rem2]].mean()     # the reference reuses num_remember2 (:249)
        return out1, out2
