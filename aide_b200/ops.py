"""Tensor-level wrappers of the individual C-ABI kernels (used by the kernel parity tests, bench.py's
roofline leg and anyone who wants one operator without the network executor).

Activations are carried as ``Act``: NHWC planes in one of the operand formats (F32 / TF32X2 / BF16).
Every function launches on the current CUDA stream and returns device tensors; nothing here falls
back to PyTorch math.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from ._lib import FMT_BF16, FMT_F16X2, FMT_F32, FMT_TF32X2, call, lib


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


class Act:
    """NHWC activation in operand format: planes[P, N, H, W, C] (P = 2 for TF32X2: hi, lo)."""

    def __init__(self, N: int, H: int, W: int, C_: int, fmt: int, device):
        self.N, self.H, self.W, self.C, self.fmt = N, H, W, C_, fmt
        dt = {FMT_BF16: torch.bfloat16, FMT_F16X2: torch.float16}.get(fmt, torch.float32)
        self.planes = torch.zeros((2 if fmt in (FMT_TF32X2, FMT_F16X2) else 1, N, H, W, C_), dtype=dt, device=device)

    @property
    def p0(self) -> int:
        return self.planes[0].data_ptr()

    @property
    def p1(self) -> Optional[int]:
        return self.planes[1].data_ptr() if self.fmt in (FMT_TF32X2, FMT_F16X2) else None

    def view(self, coff: int = 0):
        return self.p0, self.p1, self.C, coff

    def float(self) -> torch.Tensor:
        """fp32 [N,H,W,C] value (hi + lo for TF32X2; (hi + lo) / 2^8 for F16X2)."""
        v = self.planes.float().sum(0)
        return v / 256.0 if self.fmt == FMT_F16X2 else v


def from_nchw(x: torch.Tensor, fmt: int, ctot: Optional[int] = None, coff: int = 0, out: Optional[Act] = None) -> Act:
    N, C_, H, W = x.shape
    a = out if out is not None else Act(N, H, W, ctot or C_, fmt, x.device)
    call("aide_nchw_to_nhwc", fmt, x.contiguous().data_ptr(), a.p0, a.p1, a.C, coff, N, C_, H, W, _st())
    return a


def to_nchw(a: Act, C_: Optional[int] = None, coff: int = 0) -> torch.Tensor:
    C_ = C_ or a.C
    out = torch.empty((a.N, C_, a.H, a.W), dtype=torch.float32, device=a.planes.device)
    call("aide_nhwc_to_nchw", a.fmt, a.p0, a.p1, a.C, coff, out.data_ptr(), a.N, C_, a.H, a.W, _st())
    return out


def nhwc_to_nchw(t: torch.Tensor) -> torch.Tensor:
    return t.permute(0, 3, 1, 2).contiguous()


def weight_prep(w: torch.Tensor, fmt: int, dgrad: bool = False):
    """OIHW fp32 -> (p0, p1, keepalive) planes: fwd [Cout][9][Cin] or dgrad [Cin][9][Cout]."""
    cout, cin = w.shape[:2]
    dt = {FMT_BF16: torch.bfloat16, FMT_F16X2: torch.float16}.get(fmt, torch.float32)
    P = 2 if fmt in (FMT_TF32X2, FMT_F16X2) else 1
    buf = torch.empty((P, cout * 9 * cin), dtype=dt, device=w.device)
    p0, p1 = buf[0].data_ptr(), (buf[1].data_ptr() if P == 2 else None)
    if dgrad:
        call("aide_weight_prep", fmt, w.contiguous().data_ptr(), cout, cin, None, None, p0, p1, _st())
    else:
        call("aide_weight_prep", fmt, w.contiguous().data_ptr(), cout, cin, p0, p1, None, None, _st())
    return p0, p1, buf


def conv3x3(x: Act, w: torch.Tensor, bias: Optional[torch.Tensor], cin: Optional[int] = None, coff: int = 0,
            stats: bool = False, fmt: Optional[int] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """z [N,H,W,Cout] fp32 (+ per-tile BN partial statistics [rows,2,Cout])."""
    fmt = x.fmt if fmt is None else fmt
    cout = w.shape[0]
    cin = cin or w.shape[1]
    w0, w1, keep = weight_prep(w, fmt)
    z = torch.empty((x.N, x.H, x.W, cout), dtype=torch.float32, device=w.device)
    part = None
    if stats:
        rows = lib.aide_conv3x3_stat_rows(fmt, cin, cout, x.N, x.H, x.W)
        part = torch.zeros((rows, 2, cout), dtype=torch.float32, device=w.device)
    call("aide_conv3x3_fwd", fmt, x.p0, x.p1, x.C, coff, cin, w0, w1, bias.data_ptr() if bias is not None else None,
         z.data_ptr(), cout, 0, cout, x.N, x.H, x.W, part.data_ptr() if stats else None, _st())
    return z, part


def _inv_scale(dz: Act, inv_scale):
    """Device scalar 1/s for F16X2 gradient planes that hold dz*s; Act planes made by from_nchw carry the static
    activation scale 2^8, i.e. s = 2^8."""
    if dz.fmt != FMT_F16X2:
        return None
    if inv_scale is None:
        inv_scale = torch.full((1,), 1.0 / 256.0, dtype=torch.float32, device=dz.planes.device)
    return inv_scale


def conv3x3_dgrad(dz: Act, w: torch.Tensor, inv_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dX [N,H,W,Cin] fp32 from dZ (operand format) and OIHW weights."""
    cout, cin = w.shape[:2]
    w0, w1, keep = weight_prep(w, dz.fmt, dgrad=True)
    dx = torch.empty((dz.N, dz.H, dz.W, cin), dtype=torch.float32, device=w.device)
    inv = _inv_scale(dz, inv_scale)
    call("aide_conv3x3_dgrad", dz.fmt, dz.p0, dz.p1, cout, w0, w1, inv.data_ptr() if inv is not None else None,
         dx.data_ptr(), cin, 0, cin, dz.N, dz.H, dz.W, _st())
    return dx


def conv3x3_wgrad(x: Act, dz: Act, cin: Optional[int] = None, coff: int = 0,
                  inv_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    cin = cin or x.C
    cout = dz.C
    nbytes = lib.aide_conv3x3_wgrad_workspace_bytes(x.fmt, cin, cout, x.N, x.H, x.W)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.planes.device)
    dw = torch.empty((cout, cin, 3, 3), dtype=torch.float32, device=x.planes.device)
    inv = _inv_scale(dz, inv_scale)
    call("aide_conv3x3_wgrad", x.fmt, x.p0, x.p1, x.C, coff, cin, dz.p0, dz.p1,
         inv.data_ptr() if inv is not None else None, cout, x.N, x.H, x.W, ws.data_ptr(), nbytes, dw.data_ptr(), _st())
    return dw


def bn_finalize(part, count, gamma, beta, rmean, rvar, training=True, momentum=0.1, eps=1e-5):
    C_ = gamma.numel()
    ss = torch.empty((2, C_), dtype=torch.float32, device=gamma.device)
    mr = torch.empty((2, C_), dtype=torch.float32, device=gamma.device)
    call("aide_bn_finalize", part.data_ptr() if part is not None else None, part.shape[0] if part is not None else 0,
         C_, float(count), gamma.data_ptr(), beta.data_ptr(), rmean.data_ptr() if rmean is not None else None,
         rvar.data_ptr() if rvar is not None else None, momentum, eps, 1 if training else 0, ss.data_ptr(),
         mr.data_ptr(), _st())
    return ss, mr


def bn_finalize_grouped(parts, count, gamma, beta, rmean, rvar, tickets=None, momentum=0.1, eps=1e-5):
    """parts [G, rows, 2, C] (consumed: folded in place) -> scale_shift [G,2,C], mean_rstd [G,2,C]; the running
    statistics are updated once per group, in order.  tickets: zero-initialised int32 [aide_bn_ticket_slots(C)] selects
    the one-launch fold + finalize path."""
    G, rows, _, C_ = parts.shape
    ss = torch.empty((G, 2, C_), dtype=torch.float32, device=gamma.device)
    mr = torch.empty((G, 2, C_), dtype=torch.float32, device=gamma.device)
    call("aide_bn_finalize_grouped", parts.data_ptr(), rows, G, C_, float(count), gamma.data_ptr(), beta.data_ptr(),
         rmean.data_ptr() if rmean is not None else None, rvar.data_ptr() if rvar is not None else None, momentum, eps, 1,
         ss.data_ptr(), mr.data_ptr(), tickets.data_ptr() if tickets is not None else None, _st())
    return ss, mr


def bn_relu_apply(z: torch.Tensor, ss: torch.Tensor, fmt: int, pool: bool = False):
    N, H, W, C_ = z.shape
    y = Act(N, H, W, C_, fmt, z.device)
    p = Act(N, H // 2, W // 2, C_, fmt, z.device) if pool else None
    none = (None, None, 0, 0)
    call("aide_bn_relu_apply", fmt, z.data_ptr(), N, H, W, C_, ss.data_ptr(), *y.view(), *(p.view() if pool else none),
         *none, _st())
    return y, p


def upsample2x(x: Act) -> Act:
    y = Act(x.N, 2 * x.H, 2 * x.W, x.C, x.fmt, x.planes.device)
    call("aide_upsample2x_fwd", x.fmt, *x.view(), *y.view(), x.N, x.H, x.W, x.C, _st())
    return y


def upsample2x_bwd(dhi: torch.Tensor) -> torch.Tensor:
    N, H2, W2, C_ = dhi.shape
    dlo = torch.empty((N, H2 // 2, W2 // 2, C_), dtype=torch.float32, device=dhi.device)
    call("aide_upsample2x_bwd", dhi.data_ptr(), C_, 0, dlo.data_ptr(), N, H2 // 2, W2 // 2, C_, _st())
    return dlo


def conv1x1(x: Act, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    K = w.shape[0]
    out = torch.empty((x.N, K, x.H, x.W), dtype=torch.float32, device=w.device)
    call("aide_conv1x1_fwd", x.fmt, *x.view(), x.C, w.contiguous().data_ptr(), bias.data_ptr(), out.data_ptr(), K,
         x.N, x.H, x.W, _st())
    return out


def conv1x1_bwd(x: Act, w: torch.Tensor, dlogits: torch.Tensor):
    K = w.shape[0]
    rows = lib.aide_conv1x1_bwd_rows(x.N, x.H, x.W, x.C)
    part = torch.empty((rows, K * x.C + K), dtype=torch.float32, device=w.device)
    dx = torch.empty((x.N, x.H, x.W, x.C), dtype=torch.float32, device=w.device)
    dwdb = torch.empty(K * x.C + K, dtype=torch.float32, device=w.device)
    call("aide_conv1x1_bwd", x.fmt, *x.view(), x.C, w.contiguous().data_ptr(), dlogits.contiguous().data_ptr(), K,
         x.N, x.H, x.W, dx.data_ptr(), dwdb.data_ptr(), part.data_ptr(), _st())
    return dx, dwdb[:K * x.C].view(K, x.C), dwdb[K * x.C:]
