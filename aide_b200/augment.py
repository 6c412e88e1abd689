"""Augmentation on the GPU: the forward augmentation of the input images and the reverse augmentation of the
augmented forwards' outputs.

Forward (``forward_aug`` / ``augmented_views``): what the reference's data loader does per view on the CPU through
PIL -- datasetchaos_proposed/transform.py:81-106 (rotate +-degree, BILINEAR, uint8), :16-34 (horizontal flip),
:108-131 (ToTensor), :134-170 (Normalize) -- as one kernel per view, bit-exact.

Reverse (``reverseaug`` / ``reverse_aug_tensor``):

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


def forward_aug(images_u8: torch.Tensor, degrees: Sequence[float], hflips: Sequence[int], mean: torch.Tensor,
                std: torch.Tensor) -> torch.Tensor:
    """images_u8: [B,H,W,3] uint8 CUDA (the resized, un-augmented RGB images in PIL memory order).  Returns the view
    ``Normalize(ToTensor(flip(rotate(img, degree))))`` as fp32 [B,3,H,W]: rotation by ``degrees[b]`` (PIL semantics:
    counter-clockwise, BILINEAR, fill 0), then FLIP_LEFT_RIGHT where ``hflips[b]``, then /255 and (x - mean) / std with
    per-sample per-channel ``mean`` / ``std`` [B,3] (the statistics of the un-augmented image, transform.py:140-147)."""
    if not images_u8.is_cuda:
        raise RuntimeError("aide_b200 runs on CUDA only (there is no CPU fallback)")
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3:
        raise ValueError(f"expected uint8 [B,H,W,3] images, got {tuple(images_u8.shape)} {images_u8.dtype}")
    B, H, W, _ = images_u8.shape
    if len(degrees) != B or len(hflips) != B:
        raise ValueError("one (degree, hflip) pair per sample expected")
    mats, modes = [], []
    for d in degrees:
        mode, m = rotate_matrix(float(d), W, H)
        modes.append(mode)
        mats.append(m)
    dev = images_u8.device
    mt = torch.tensor(mats, dtype=torch.float64).to(dev)
    md = torch.tensor(modes, dtype=torch.int32).to(dev)
    fl = torch.tensor([1 if f else 0 for f in hflips], dtype=torch.int32).to(dev)
    mean = mean.to(device=dev, dtype=torch.float32).reshape(B, 3).contiguous()
    std = std.to(device=dev, dtype=torch.float32).reshape(B, 3).contiguous()
    src = images_u8.contiguous()
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    call("aide_forward_aug", src.data_ptr(), out.data_ptr(), mt.data_ptr(), md.data_ptr(), fl.data_ptr(), mean.data_ptr(),
         std.data_ptr(), B, H, W, torch.cuda.current_stream().cuda_stream)
    return out


def augmented_views(img1_u8: torch.Tensor, img2_u8: torch.Tensor, mean1, std1, mean2, std2, n_views: int = 4,
                    rotation: float = 60.0, rng=None):
    """The ``augset`` of the reference's data loader for a whole batch, built on the GPU: for every view k one random
    rotation in [-rotation, rotation) (transform.py:86-104) and one coin flip (:18-33) PER SAMPLE, applied to both
    modalities.  Returns (augset dict with ``augno``, ``degree{k}``, ``hflip{k}`` lists -- what ``reverseaug`` and
    ``AideTrainer.step(augset=...)`` take -- and a list of n_views (modal1, modal2) fp32 [B,3,H,W] pairs)."""
    import random as _random
    rng = rng or _random
    B = img1_u8.shape[0]
    augset: Dict = {"augno": [n_views] * B}
    for k in range(1, n_views + 1):           # RandomRotate runs before RandomHorizontallyFlip (train script :191-197)
        augset[f"degree{k}"] = [rng.random() * 2 * rotation - rotation for _ in range(B)]
    for k in range(1, n_views + 1):
        augset[f"hflip{k}"] = [1 if rng.random() < 0.5 else 0 for _ in range(B)]
    views = []
    for k in range(1, n_views + 1):
        views.append((forward_aug(img1_u8, augset[f"degree{k}"], augset[f"hflip{k}"], mean1, std1),
                      forward_aug(img2_u8, augset[f"degree{k}"], augset[f"hflip{k}"], mean2, std2)))
    return augset, views
