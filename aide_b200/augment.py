"""Reverse augmentation of the augmented forwards' outputs on the GPU.

Drop-in for the reference's ``reverseaug(augset, augoutput, classno)``
(train_files/trainchaos_proposed_30cases1labeled.py:81-95), which copies every (sample, view, class) plane to the CPU,
flips / rotates it with PIL (``Image.rotate(-degree, Image.BILINEAR)``) and copies it back.  Here the host only builds
PIL's inverse affine matrix per (sample, view) -- same formula and rounding as ``PIL.Image.Image.rotate`` -- and one
kernel per view reproduces Pillow's bilinear arithmetic bit for bit (csrc/augment.cu).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch

from ._lib import call


def rotate_matrix(angle_deg: float, w: int, h: int) -> Tuple[int, List[float]]:
    """(mode, inverse affine matrix) of ``Image.rotate(angle_deg, BILINEAR)`` on a w x h image: mode 0 = affine,
    1 = copy (angle % 360 == 0), 2 = rotate 180, 3 / 4 = PIL's ROTATE_90 / ROTATE_270 fast paths (square images)."""
    angle = float(angle_deg) % 360.0
    ident = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    if angle == 0:
        return 1, ident
    if angle == 180:
        return 2, ident
    if angle in (90, 270) and w == h:
        return (3 if angle == 90 else 4), ident
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return 0, m


def reverse_aug_tensor(x: torch.Tensor, degrees: Sequence[float], hflips: Sequence[int]) -> torch.Tensor:
    """x: [B,K,H,W] fp32 CUDA logits of ONE augmented view; degrees/hflips: the forward augmentation of each sample
    (the reverse rotation is ``0 - degree``, as in the reference).  Returns a new tensor."""
    if not x.is_cuda:
        raise RuntimeError("aide_b200 runs on CUDA only (there is no CPU fallback)")
    B, K, H, W = x.shape
    if len(degrees) != B or len(hflips) != B:
        raise ValueError("one (degree, hflip) pair per sample expected")
    mats, modes = [], []
    for d in degrees:
        mode, m = rotate_matrix(0 - float(d), W, H)
        modes.append(mode)
        mats.append(m)
    dev = x.device
    mt = torch.tensor(mats, dtype=torch.float64).to(dev)
    md = torch.tensor(modes, dtype=torch.int32).to(dev)
    fl = torch.tensor([1 if f else 0 for f in hflips], dtype=torch.int32).to(dev)
    src = x.detach().contiguous().float()
    out = torch.empty_like(src)
    call("aide_reverse_aug", src.data_ptr(), out.data_ptr(), mt.data_ptr(), md.data_ptr(), fl.data_ptr(), B, K, H, W,
         torch.cuda.current_stream().cuda_stream)
    return out


def reverseaug(augset: Dict, augoutput: List[torch.Tensor], classno: int) -> List[torch.Tensor]:
    """Same call as the reference's ``reverseaug``: ``augset['augno'][b]`` views are un-augmented for sample b using
    ``augset['hflip{k}'][b]`` and ``augset['degree{k}'][b]`` (k = 1-based view index); views beyond a sample's
    ``augno`` are left untouched.  ``augoutput`` is updated in place and returned, like the reference."""
    n_b = len(augset["augno"])
    for aug_idx in range(len(augoutput)):
        t = augoutput[aug_idx]
        if t.shape[1] != classno:
            raise ValueError(f"augoutput[{aug_idx}] has {t.shape[1]} classes, expected {classno}")
        deg = [float(augset[f"degree{aug_idx + 1}"][b]) for b in range(n_b)]
        flip = [int(augset[f"hflip{aug_idx + 1}"][b]) for b in range(n_b)]
        done = reverse_aug_tensor(t[:n_b], deg, flip)
        keep = torch.tensor([int(augset["augno"][b]) > aug_idx for b in range(n_b)], device=t.device)
        t[:n_b] = torch.where(keep.view(-1, 1, 1, 1), done, t[:n_b])
    return augoutput
