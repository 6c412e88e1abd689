"""Host-side executor for the AIDE networks on libaide_b200 (sm_100a).

A network (fuseunet / UNet, reference: models_twomodalinputs/fuseunet.py:43-91,
models_singlemodalinput/UNet.py:152-165) is described once as a *plan*: a list of conv3x3+BN+ReLU
units, bilinear upsamples and the 1x1 head operating on named NHWC buffers.  Channel concatenation
(torch.cat in the reference) never copies: producers write into channel slices ("views") of the
consumer's input buffer, 2x2 max-pool outputs are emitted by the producer's BN-apply kernel.

Forward and backward are each a straight sequence of C-ABI calls on the current CUDA stream; there
is no CPU compute and no fallback.  The backward plan is derived from the forward plan: the
gradient of a unit's output is the sum of the consumers' dgrad slices (same resolution), max-pool
routed slices (half resolution) and upsample-transposed gradients.

PyTorch is used for device memory (one arena tensor per forward / backward), streams and autograd
registration only.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import FMT_BF16, FMT_F16X2, FMT_F32, FMT_TF32X2, call, lib

# mode -> (operand format of forwards that keep a tape for backward, format of inference-only forwards)
#   parity       fp32-equivalent split-precision tcgen05, F16X2 everywhere: two fp16 planes of the power-of-two
#                pre-scaled tensor, 3 kind::f16 MMAs (lo*hi + hi*lo + hi*hi) = 22-bit products at the full f16 tensor
#                rate; gradients carry a per-tensor power-of-two scale chosen on the device (aide_bn_relu_bwd_apply)
#   parity_tf32  TF32X2 everywhere (3 kind::tf32 MMAs at half the rate, twice the operand bytes)
#   parity_mixed TF32X2 for the train forward/backward, F16X2 for forwards without a backward
#   fast         single-pass BF16 (NOT a parity mode)      exact  fp32 CUDA cores
MODES = {"exact": (FMT_F32, FMT_F32), "parity": (FMT_F16X2, FMT_F16X2), "parity_tf32": (FMT_TF32X2, FMT_TF32X2),
         "parity_f16": (FMT_F16X2, FMT_F16X2), "parity_mixed": (FMT_TF32X2, FMT_F16X2), "fast": (FMT_BF16, FMT_BF16)}
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def default_mode() -> str:
    """Engine precision mode (see MODES); opt-in via AIDE_B200_MODE."""
    m = os.environ.get("AIDE_B200_MODE", "parity")
    if m not in MODES:
        raise ValueError(f"AIDE_B200_MODE must be one of {sorted(MODES)}, got {m!r}")
    return m


def mode_format(mode: str, keep_tape: bool) -> int:
    return MODES[mode][0 if keep_tape else 1]


def _planes(fmt: int) -> int:
    return 2 if fmt in (FMT_TF32X2, FMT_F16X2) else 1


def _esize(fmt: int) -> int:
    return 2 if fmt in (FMT_BF16, FMT_F16X2) else 4


def _align(n: int, a: int = 1024) -> int:
    return (n + a - 1) // a * a


# ------------------------------------------------------------------------------------------------
# plan description
# ------------------------------------------------------------------------------------------------
@dataclass
class Unit:
    """conv3x3(+bias) -> BatchNorm2d -> ReLU (netblocks.py:30-33 / :17-19)."""
    name: str            # e.g. "modal1_downblock1.block.1" (for messages)
    conv: str            # parameter prefix of the conv  (…conv1 / …bilinear_up.1)
    bn: str              # parameter prefix of the BN    (…bn1   / …bilinear_up.2)
    cin: int
    cout: int
    level: int           # resolution level: h = H >> level
    src: Tuple[str, int]             # (buffer, channel offset) of the input view
    dst: Optional[Tuple[str, int]]   # full-resolution destination view of y
    pools: List[Tuple[str, int]] = field(default_factory=list)  # half-resolution destinations of maxpool(y)
    first: bool = False              # input is a network input (fp32 NHWC, 3 channels): CUDA-core path
    transposed: bool = False         # the parameter is a ConvTranspose2d(k=2, s=2) weight [Cin,Cout,2,2] (netblocks.py:12):
                                     # run as conv3x3 on the zero-inserted input with K[co,ci,1-a,1-b] = W[ci,co,a,b]


@dataclass
class Attention:
    """Spatial_Attention gate (netblocks.py:68-89 / UNet.py:85-108) applied to a block's output: t = gate(y) * y.
    `src` is the block's ungated output (its own buffer), `dst` / `pools` are where the gated tensor and its 2x2 max-pool
    go -- the slots the plain networks let the block's second unit write directly."""
    name: str            # parameter prefix, e.g. "modal1_sa1" / "sa1"
    c: int
    r: int               # c // reduction
    level: int
    src: Tuple[str, int]
    dst: Optional[Tuple[str, int]]
    pools: List[Tuple[str, int]] = field(default_factory=list)
    dilation: int = 4


@dataclass
class Upsample:
    src: str
    dst: str
    c: int
    level: int           # level of the SOURCE (low resolution)
    mode: str = "bilinear"   # "bilinear": nn.Upsample(2, bilinear, align_corners=True); "zero": zero insertion (ConvTranspose2d)


@dataclass
class Plan:
    kind: str
    n_inputs: int
    num_classes: int
    bufs: Dict[str, Tuple[int, int]]   # activation buffers: name -> (level, channels)
    ops: list                          # forward order: ("input", i, buf) | Unit | Attention | Upsample | ("head", buf, cin)
    units: List[Unit]
    atts: List[Attention] = field(default_factory=list)


def _block(ops, units, prefix, cin, cout, level, src, dst, pools, mid_buf, bufs, first=False, att=None, atts=None):
    """basic_block = two units; the first writes `mid_buf`, the second writes dst/pools.  att = (parameter prefix,
    reduction, dilation): the block is followed by a Spatial_Attention gate -- the second unit then writes its own
    buffer and the gate op produces dst/pools."""
    bufs[mid_buf] = (level, cout)
    u1 = Unit(prefix + ".1", prefix + ".conv1", prefix + ".bn1", cin, cout, level, src, (mid_buf, 0), [], first)
    if att is None:
        u2 = Unit(prefix + ".2", prefix + ".conv2", prefix + ".bn2", cout, cout, level, (mid_buf, 0), dst, pools)
        ops += [u1, u2]
        units += [u1, u2]
        return
    name, reduction, dilation = att
    ybuf = "say:" + name
    bufs[ybuf] = (level, cout)
    u2 = Unit(prefix + ".2", prefix + ".conv2", prefix + ".bn2", cout, cout, level, (mid_buf, 0), (ybuf, 0), [])
    a = Attention(name, cout, cout // reduction, level, (ybuf, 0), dst, pools, dilation)
    ops += [u1, u2, a]
    units += [u1, u2]
    atts.append(a)


def _decoder(ops, units, bufs, bottom: str, skips_c: Sequence[int], num_classes: int, learned_bilinear: bool = False):
    """Four UNet_basic_up_block (netblocks.py:137-147) + last_conv1.  Skip tensors already live in the
    upper half of the cat buffers (cat((upsampled, skip)) -> channels [C, 2C)).  learned_bilinear: the up path is
    ConvTranspose2d(k=2,s=2) -> BN -> ReLU (netblocks.py:11-14; Sequential indices 0, 1) instead of
    Upsample -> Conv2d(3x3) -> BN -> ReLU (indices 1, 2)."""
    x, cx = bottom, skips_c[0] * 2
    for i, c in enumerate(skips_c, 1):            # c = 512, 256, 128, 64 ; level = 4 - i
        lvl = 4 - i
        up, cat, mid, out = f"up{i}", f"cat{i}", f"u{i}m", f"u{i}o"
        bufs[up] = (lvl, cx)
        bufs[out] = (lvl, c)
        ops.append(Upsample(x, up, cx, lvl + 1, "zero" if learned_bilinear else "bilinear"))
        ci, bi = (0, 1) if learned_bilinear else (1, 2)
        u = Unit(f"up_block{i}.up", f"up_block{i}.bilinear_up.{ci}", f"up_block{i}.bilinear_up.{bi}", cx, c, lvl,
                 (up, 0), (cat, 0), [], transposed=learned_bilinear)
        ops.append(u)
        units.append(u)
        _block(ops, units, f"up_block{i}.block", 2 * c, c, lvl, (cat, 0), (out, 0), [], mid, bufs)
        x, cx = out, c
    ops.append(("head", x, cx))


def plan_fuseunet(num_classes: int = 2, learned_bilinear: bool = False, attention: bool = False, separate: bool = False,
                  reduction: int = 16, dilation: int = 4) -> Plan:
    """fuseunet.forward (fuseunet.py:43-91): modal-1 encoder consumes the fused concat, modal-2 is independent.
    attention: fuseunetsa (fuseunet.py:138-208) -- every encoder block is gated by its Spatial_Attention; separate:
    fuseunetsaseparate (:255-325) -- modal-1 continues from its OWN gated output instead of the fused concat."""
    bufs: Dict[str, Tuple[int, int]] = {"in0": (0, 3), "in1": (0, 3)}
    ops: list = [("input", 0, "in0"), ("input", 1, "in1")]
    units: List[Unit] = []
    atts: List[Attention] = []
    width = [32, 64, 128, 256, 512]
    # fused tensors y1..y4 are the upper halves of cat4..cat1; y5 is its own buffer; p{l} = maxpool(y_l)
    for lvl, c in enumerate(width):
        fused = ("y5", 0) if lvl == 4 else (f"cat{4 - lvl}", 2 * c)     # cat width is 4c: [up 2c | skip 2c]
        if lvl == 4:
            bufs["y5"] = (4, 2 * c)
        else:
            bufs[f"cat{4 - lvl}"] = (lvl, 4 * c)
            bufs[f"p{lvl + 1}"] = (lvl + 1, 2 * c)
        pool_a = [] if lvl == 4 else [(f"p{lvl + 1}", 0)]
        pool_b = [] if lvl == 4 else [(f"p{lvl + 1}", c)]
        src_a = ("in0", 0) if lvl == 0 else (f"p{lvl}", 0)
        cin_a = 3 if lvl == 0 else (width[lvl - 1] if separate else 2 * width[lvl - 1])
        src_b = ("in1", 0) if lvl == 0 else (f"p{lvl}", width[lvl - 1])   # modal-2 = second half of the pooled concat
        cin_b = 3 if lvl == 0 else width[lvl - 1]
        att_a = (f"modal1_sa{lvl + 1}", reduction, dilation) if attention else None
        att_b = (f"modal2_sa{lvl + 1}", reduction, dilation) if attention else None
        _block(ops, units, f"modal1_downblock{lvl + 1}.block", cin_a, c, lvl, src_a, (fused[0], fused[1]), pool_a,
               f"a{lvl + 1}m", bufs, first=lvl == 0, att=att_a, atts=atts)
        _block(ops, units, f"modal2_downblock{lvl + 1}.block", cin_b, c, lvl, src_b, (fused[0], fused[1] + c), pool_b,
               f"b{lvl + 1}m", bufs, first=lvl == 0, att=att_b, atts=atts)
    _decoder(ops, units, bufs, "y5", [512, 256, 128, 64], num_classes, learned_bilinear)
    kind = "fuseunet" if not attention else ("fuseunetsaseparate" if separate else "fuseunetsa")
    return Plan(kind, 2, num_classes, bufs, ops, units, atts)


def plan_unet(num_classes: int = 2, learned_bilinear: bool = False, attention: bool = False, reduction: int = 16,
              dilation: int = 4, base: int = 64) -> Plan:
    """UNet.forward (UNet.py:152-165); max-pool inside down blocks 2..5 (UNet.py:117-121).  attention: UNetsa
    (UNet.py:190-208), every down block gated by sa{level}."""
    bufs: Dict[str, Tuple[int, int]] = {"in0": (0, 3)}
    ops: list = [("input", 0, "in0")]
    units: List[Unit] = []
    atts: List[Attention] = []
    width = [base << i for i in range(5)]          # UNet: 64..1024; UNet128 / UNet32 / ... (UNet.py:210-385): other bases
    for lvl, c in enumerate(width):
        if lvl == 4:
            bufs["x5"] = (4, c)
            dst = ("x5", 0)
        else:
            bufs[f"cat{4 - lvl}"] = (lvl, 2 * c)
            bufs[f"p{lvl + 1}"] = (lvl + 1, c)
            dst = (f"cat{4 - lvl}", c)
        pools = [] if lvl == 4 else [(f"p{lvl + 1}", 0)]
        src = ("in0", 0) if lvl == 0 else (f"p{lvl}", 0)
        cin = 3 if lvl == 0 else width[lvl - 1]
        _block(ops, units, f"down_block{lvl + 1}.block", cin, c, lvl, src, dst, pools, f"d{lvl + 1}m", bufs,
               first=lvl == 0, att=(f"sa{lvl + 1}", reduction, dilation) if attention else None, atts=atts)
    _decoder(ops, units, bufs, "x5", width[3::-1], num_classes, learned_bilinear)
    return Plan("unetsa" if attention else "unet", 1, num_classes, bufs, ops, units, atts)


# ------------------------------------------------------------------------------------------------
# memory layout of one forward pass ("tape")
# ------------------------------------------------------------------------------------------------
class Layout:
    """Byte offsets of every buffer of a plan for a given (N, H, W, fmt) inside one arena."""

    def __init__(self, plan: Plan, N: int, H: int, W: int, fmt: int, groups: int = 1):
        self.groups = groups
        if H % 16 or W % 16:
            raise ValueError(f"input height/width must be multiples of 16 (got {H}x{W}); the nets pool 4 times")
        self.plan, self.N, self.H, self.W, self.fmt = plan, N, H, W, fmt
        self.off: Dict[str, int] = {}
        self.plane_stride: Dict[str, int] = {}
        cur = 0
        for name, (lvl, c) in plan.bufs.items():
            h, w = H >> lvl, W >> lvl
            f = FMT_F32 if name.startswith("in") else fmt
            nbytes = _align(N * h * w * c * _esize(f))
            self.off[name] = cur
            self.plane_stride[name] = nbytes
            cur += nbytes * _planes(f)
        self.stat_rows: Dict[str, int] = {}
        for u in plan.units:
            h, w = H >> u.level, W >> u.level
            ufmt = FMT_F32 if u.first else fmt
            rows = lib.aide_conv3x3_stat_rows(ufmt, u.cin, u.cout, N, h, w)
            self.stat_rows[u.name] = rows
            self.off["z:" + u.name] = cur
            cur += _align(N * h * w * u.cout * 4)
            self.off["st:" + u.name] = cur
            cur += _align(rows * 2 * u.cout * 4)
            self.off["ss:" + u.name] = cur            # scale_shift [G][2][C] then mean_rstd [G][2][C]
            cur += _align(groups * 4 * u.cout * 4)
        for a in plan.atts:                           # Spatial_Attention branch: fp32 intermediates (kept for backward)
            h, w = H >> a.level, W >> a.level
            for key, n in (("sa1:", N * h * w * a.r), ("sa2:", N * h * w * a.r), ("sa3:", N * h * w * a.r),
                           ("saa:", N * h * w), ("sag:", N * h * w), ("sast:", lib.aide_sa_stat_rows(N, h, w) * 2),
                           ("sass:", groups * 4)):
                self.off[key + a.name] = cur
                cur += _align(n * 4)
        self.total = cur

    def buf_fmt(self, name: str) -> int:
        return FMT_F32 if name.startswith("in") else self.fmt


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_SIDE_STREAMS: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}


def _side_stream(main: "torch.cuda.Stream") -> "torch.cuda.Stream":
    """One helper stream per (device, main stream): the two networks of a step run on two streams and get two of these."""
    key = (main.device.index or 0, main.cuda_stream)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(main.device)
    return s


# ------------------------------------------------------------------------------------------------
# weights in operand format (re-derived when a parameter's version changes)
# ------------------------------------------------------------------------------------------------
def conv_weight_oihw(u: Unit, params: Dict[str, torch.Tensor]) -> torch.Tensor:
    """The unit's 3x3 weight in nn.Conv2d layout [Cout,Cin,3,3].  For a ConvTranspose2d(k=2,s=2) parameter
    W [Cin,Cout,2,2] this is K[co,ci,1-a,1-b] = W[ci,co,a,b] with the other five taps zero (tensor re-indexing only)."""
    w = params[u.conv + ".weight"]
    if not u.transposed:
        return w
    k = w.new_zeros((u.cout, u.cin, 3, 3))
    k[:, :, 0:2, 0:2] = w.detach().permute(1, 0, 2, 3).flip(2, 3)
    return k


def transposed_weight_grad(gk: torch.Tensor) -> torch.Tensor:
    """dW [Cin,Cout,2,2] of a ConvTranspose2d weight from the gradient of its 3x3 stand-in K [Cout,Cin,3,3]."""
    return gk[:, :, 0:2, 0:2].flip(2, 3).permute(1, 0, 2, 3).contiguous()


class PreparedWeights:
    """Operand-format copies of every conv3x3 weight: forward [Cout][9][Cin] and dgrad [Cin][9][Cout]
    planes (aide_weight_prep).  One instance per parameter *version*; tapes hold a reference so a later
    optimiser step cannot invalidate the weights a pending backward needs."""

    def __init__(self, plan: Plan, params: Dict[str, torch.Tensor], fmt: int, need_dgrad: bool):
        self.key = tuple((params[u.conv + ".weight"].data_ptr(), params[u.conv + ".weight"]._version)
                         for u in plan.units)
        self.has_dgrad = need_dgrad
        dev = params[plan.units[0].conv + ".weight"].device
        total = 0
        self.off: Dict[str, Tuple[int, int, int]] = {}   # name -> (fwd offset, dgrad offset, plane stride)
        for u in plan.units:
            f = FMT_F32 if u.first else fmt
            pb = _align(u.cout * 9 * u.cin * _esize(f), 256)
            fwd = total
            total += pb * _planes(f)
            dg = -1
            if need_dgrad and not u.first:
                dg = total
                total += pb * _planes(f)
            self.off[u.name] = (fwd, dg, pb)
        self.arena = torch.empty(total, dtype=torch.uint8, device=dev)
        base = self.arena.data_ptr()
        st = _stream()
        batch = []                                  # tensor-core layers: ONE launch for the whole network
        for u in plan.units:
            f = FMT_F32 if u.first else fmt
            fwd, dg, pb = self.off[u.name]
            two = _planes(f) == 2
            w = conv_weight_oihw(u, params)
            args = (w.data_ptr(), u.cout, u.cin, base + fwd, base + fwd + pb if two else None,
                    base + dg if dg >= 0 else None, base + dg + pb if (dg >= 0 and two) else None)
            if f != FMT_F32 and u.cout % 32 == 0 and u.cin % 32 == 0 and not u.transposed:
                batch.append(args)
            else:
                call("aide_weight_prep", f, *args, st)
        if batch:
            n = len(batch)
            col = lambda i, ty: (ty * n)(*[b[i] for b in batch])
            call("aide_weight_prep_batch", fmt, n, col(0, C.c_void_p), col(1, C.c_int), col(2, C.c_int),
                 col(3, C.c_void_p), col(4, C.c_void_p), col(5, C.c_void_p), col(6, C.c_void_p), st)

    def fwd(self, u: Unit, fmt: int):
        fwd, _, pb = self.off[u.name]
        b = self.arena.data_ptr() + fwd
        return b, (b + pb if _planes(FMT_F32 if u.first else fmt) == 2 else None)

    def dgrad(self, u: Unit, fmt: int):
        _, dg, pb = self.off[u.name]
        b = self.arena.data_ptr() + dg
        return b, (b + pb if _planes(fmt) == 2 else None)


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
class Tape:
    __slots__ = ("layout", "arena", "weights", "training", "inputs_shape", "group")


def _view(layout: Layout, base: int, name: str, coff: int):
    """(p0, p1, ctot, coff) of a channel view of activation buffer `name`."""
    lvl, c = layout.plan.bufs[name]
    p0 = base + layout.off[name]
    p1 = p0 + layout.plane_stride[name] if _planes(layout.buf_fmt(name)) == 2 else None
    return p0, p1, c, coff


def ticket_offsets(plan: Plan) -> Tuple[Dict[str, int], int]:
    """Word offsets of every unit's statistics tickets inside one per-network int32 buffer (zero-initialised once; the
    kernels leave the counters at zero), and the buffer's length."""
    off, cur = {}, 0
    for u in plan.units:
        off[u.name] = cur
        cur += lib.aide_bn_ticket_slots(u.cout)
    return off, cur


def run_forward(plan: Plan, layout: Layout, params: Dict[str, torch.Tensor], weights: PreparedWeights,
                inputs: Sequence, training: bool, logits: torch.Tensor, arena: torch.Tensor, groups: int = 1,
                tickets: Optional[Tuple[torch.Tensor, Dict[str, int]]] = None) -> None:
    """One forward pass.  groups > 1: `inputs` is a list of `groups` input tuples (the AIDE step's augmented views,
    trainchaos_proposed_30cases1labeled.py:265-269) stacked along the batch: every convolution / upsample / the head
    run ONCE on the stacked batch, while BatchNorm statistics, the running-statistics updates and the normalisation
    stay per group, in order -- exactly what `groups` separate forward calls compute.  Every group keeps its own
    scale/shift and mean/rstd rows, so a tape of the stacked forward serves a backward through any one group."""
    N, H, W, fmt = layout.N, layout.H, layout.W, layout.fmt
    G = groups
    Ng = N // G
    base = arena.data_ptr()
    st = _stream()

    def gview(name, coff, lvl, g):
        """channel view of buffer `name`, advanced to the first image of group g"""
        p0, p1, ctot, co = _view(layout, base, name, coff)
        skip = g * Ng * (H >> lvl) * (W >> lvl) * ctot * _esize(layout.buf_fmt(name))
        return p0 + skip, (p1 + skip if p1 is not None else None), ctot, co

    fuse_eval = (not training) and os.environ.get("AIDE_B200_EVAL_FUSE", "1") != "0"
    if fuse_eval:
        # eval mode: scale / shift of EVERY BatchNorm (units and attention gates) in one launch; the units' BN + ReLU
        # (+ max-pool) then run inside the conv epilogue (aide_conv3x3_bn_relu_fwd)
        rows = [(u.bn, u.cout, base + layout.off["ss:" + u.name]) for u in plan.units]
        rows += [(a.name + ".bn", 1, base + layout.off["sass:" + a.name]) for a in plan.atts]
        n = len(rows)
        arr = lambda vals, ty: (ty * n)(*vals)
        call("aide_bn_eval_scale_shift_batch", n,
             arr([params[b + ".weight"].data_ptr() for b, _, _ in rows], C.c_void_p),
             arr([params[b + ".bias"].data_ptr() for b, _, _ in rows], C.c_void_p),
             arr([params[b + ".running_mean"].data_ptr() for b, _, _ in rows], C.c_void_p),
             arr([params[b + ".running_var"].data_ptr() for b, _, _ in rows], C.c_void_p),
             arr([c for _, c, _ in rows], C.c_int), arr([o for _, _, o in rows], C.c_void_p), BN_EPS, st)
    for op in plan.ops:
        if isinstance(op, tuple) and op[0] == "input":
            _, i, name = op
            for g in range(G):
                src = inputs[g][i] if G > 1 else inputs[i]
                p0, p1, ctot, _ = gview(name, 0, 0, g)
                call("aide_nchw_to_nhwc", FMT_F32, src.data_ptr(), p0, None, ctot, 0, Ng, 3, H, W, st)
        elif isinstance(op, Unit):
            u = op
            h, w = H >> u.level, W >> u.level
            ufmt = FMT_F32 if u.first else fmt
            x0, x1, xct, xco = _view(layout, base, u.src[0], u.src[1])
            w0, w1 = weights.fwd(u, fmt)
            z = base + layout.off["z:" + u.name]
            stp = base + layout.off["st:" + u.name]
            ss = base + layout.off["ss:" + u.name]
            d = _view(layout, base, u.dst[0], u.dst[1]) if u.dst else (None, None, 0, 0)
            pa = _view(layout, base, u.pools[0][0], u.pools[0][1]) if len(u.pools) > 0 else (None, None, 0, 0)
            pb = _view(layout, base, u.pools[1][0], u.pools[1][1]) if len(u.pools) > 1 else (None, None, 0, 0)
            if fuse_eval and not u.first and lib.aide_conv3x3_bn_relu_ok(ufmt, u.cin, u.cout, N, h, w):
                call("aide_conv3x3_bn_relu_fwd", ufmt, x0, x1, xct, xco, u.cin, w0, w1, params[u.conv + ".bias"].data_ptr(),
                     ss, u.cout, N, h, w, *d, *pa, *pb, st)
                continue
            call("aide_conv3x3_fwd", ufmt, x0, x1, xct, xco, u.cin, w0, w1, params[u.conv + ".bias"].data_ptr(),
                 z, u.cout, 0, u.cout, N, h, w, stp if training else None, st)
            if fuse_eval:                       # scale / shift already there (one group: every image shares them)
                call("aide_bn_relu_apply_grouped", fmt, z, N, N, h, w, u.cout, ss, *d, *pa, *pb, st)
                continue
            rows_g = layout.stat_rows[u.name] // G
            call("aide_bn_finalize_grouped", stp, rows_g, G, u.cout, float(Ng * h * w),
                 params[u.bn + ".weight"].data_ptr(), params[u.bn + ".bias"].data_ptr(),
                 params[u.bn + ".running_mean"].data_ptr(), params[u.bn + ".running_var"].data_ptr(),
                 BN_MOMENTUM, BN_EPS, 1 if training else 0, ss, ss + G * 2 * u.cout * 4,
                 tickets[0].data_ptr() + 4 * tickets[1][u.name] if tickets is not None else None, st)
            call("aide_bn_relu_apply_grouped", fmt, z, N, Ng, h, w, u.cout, ss, *d, *pa, *pb, st)
        elif isinstance(op, Attention):
            a = op
            h, w = H >> a.level, W >> a.level
            y = _view(layout, base, a.src[0], a.src[1])
            P = lambda k: params[f"{a.name}.{k}"].data_ptr()
            o = lambda k: base + layout.off[k + a.name]
            call("aide_sa_fwd", fmt, *y, a.c, a.r, a.dilation, P("conv1.weight"), P("conv1.bias"), P("conv2.weight"),
                 P("conv2.bias"), P("conv3.weight"), P("conv3.bias"), P("conv4.weight"), P("conv4.bias"),
                 o("sa1:"), o("sa2:"), o("sa3:"), o("saa:"), o("sast:"), N, h, w, st)
            rows_g = lib.aide_sa_stat_rows(N, h, w) // G
            ss = o("sass:")
            if not fuse_eval:
                call("aide_bn_finalize_grouped", o("sast:"), rows_g, G, 1, float(Ng * h * w), P("bn.weight"), P("bn.bias"),
                     P("bn.running_mean"), P("bn.running_var"), BN_MOMENTUM, BN_EPS, 1 if training else 0, ss,
                     ss + G * 2 * 4, None, st)
            d = _view(layout, base, a.dst[0], a.dst[1]) if a.dst else (None, None, 0, 0)
            pa = _view(layout, base, a.pools[0][0], a.pools[0][1]) if len(a.pools) > 0 else (None, None, 0, 0)
            pb = _view(layout, base, a.pools[1][0], a.pools[1][1]) if len(a.pools) > 1 else (None, None, 0, 0)
            call("aide_sa_gate_apply", fmt, *y, o("saa:"), ss, N, N if fuse_eval else Ng, h, w, a.c, o("sag:"), *d, *pa,
                 *pb, st)
        elif isinstance(op, Upsample):
            h, w = H >> op.level, W >> op.level
            s = _view(layout, base, op.src, 0)
            d = _view(layout, base, op.dst, 0)
            call("aide_upsample2x_fwd" if op.mode == "bilinear" else "aide_zero_insert2x_fwd", fmt, *s, *d, N, h, w, op.c, st)
        else:
            _, name, cin = op
            x0, x1, xct, xco = _view(layout, base, name, 0)
            call("aide_conv1x1_fwd", fmt, x0, x1, xct, xco, cin, params["last_conv1.weight"].data_ptr(),
                 params["last_conv1.bias"].data_ptr(), logits.data_ptr(), plan.num_classes, N, H, W, st)


# ------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------
class BackwardPlan:
    """Gradient routing derived from the forward plan (who consumes which channel slice)."""

    def __init__(self, plan: Plan):
        self.plan = plan
        # consumers of each activation buffer: list of (kind, obj, coff, c)
        cons: Dict[str, list] = {b: [] for b in plan.bufs}
        for op in plan.ops:
            if isinstance(op, Unit):
                cons[op.src[0]].append(("unit", op, op.src[1], op.cin))
            elif isinstance(op, Attention):
                cons[op.src[0]].append(("att", op, op.src[1], op.c))
            elif isinstance(op, Upsample):
                cons[op.src].append(("ups", op, 0, op.c))
            elif op[0] == "head":
                cons[op[1]].append(("head", op, 0, op[2]))
        self.sources: Dict[str, Tuple[list, list]] = {}
        for u in list(plan.units) + list(plan.atts):      # an attention gate routes gradients like a unit's output does
            cw = u.cout if isinstance(u, Unit) else u.c
            direct, pooled = [], []
            if u.dst:
                direct = self._covering(cons[u.dst[0]], u.dst[1], cw, u)
            for (pbuf, pco) in u.pools:
                pooled += self._covering(cons[pbuf], pco, cw, u)
            if not direct and not pooled:
                raise RuntimeError(f"unit {u.name} has no consumer")
            if len(direct) > 3 or len(pooled) > 3:
                raise RuntimeError(f"unit {u.name}: too many gradient sources")
            self.sources[u.name] = (direct, pooled)

    @staticmethod
    def _covering(consumers, coff, c, u):
        out = []
        for kind, obj, ccoff, cc in consumers:
            if ccoff <= coff and coff + c <= ccoff + cc:
                out.append((kind, obj, coff - ccoff, cc))       # (kind, consumer, offset inside its dX, its dX width)
            elif not (coff + c <= ccoff or ccoff + cc <= coff):
                raise RuntimeError(f"partial overlap between {u.name} output and a consumer view")
        return out


class GradLayout:
    """Offsets (in floats) of every parameter gradient inside ONE flat fp32 buffer -- the unit of the
    data-parallel all-reduce and of the fused Adam step.  BN beta/gamma gradients and the
    head's weight/bias gradients are adjacent because the kernels emit them as one vector."""

    def __init__(self, plan: Plan):
        self.off: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        cur = 0

        def add(name, shape):
            nonlocal cur
            n = 1
            for d in shape:
                n *= d
            self.off[name] = (cur, tuple(shape))
            cur += (n + 3) // 4 * 4

        for u in plan.units:
            # a ConvTranspose2d(k=2, s=2) parameter keeps ITS shape [Cin,Cout,2,2]: the slot is the parameter's gradient
            # (the 3x3 stand-in's gradient lives in a scratch buffer of the backward pass and is re-indexed into it)
            add(u.conv + ".weight", (u.cin, u.cout, 2, 2) if u.transposed else (u.cout, u.cin, 3, 3))
            add(u.conv + ".bias", (u.cout,))
            add(u.bn + ".bias", (u.cout,))       # dbeta  } adjacent: written as one [2][C] vector
            add(u.bn + ".weight", (u.cout,))     # dgamma }
        for a in plan.atts:       # aide_sa_bwd_chain writes {dW, db} pairs contiguously and {dbeta, dgamma} as one pair
            for conv, shape in (("conv1", (a.r, a.c, 1, 1)), ("conv2", (a.r, a.r, 3, 3)), ("conv3", (a.r, a.r, 3, 3)),
                                ("conv4", (1, a.r, 1, 1))):
                n = shape[0] * shape[1] * shape[2] * shape[3]
                self.off[f"{a.name}.{conv}.weight"] = (cur, shape)
                self.off[f"{a.name}.{conv}.bias"] = (cur + n, (shape[0],))
                cur = (cur + n + shape[0] + 3) // 4 * 4
            self.off[f"{a.name}.bn.bias"] = (cur, (1,))
            self.off[f"{a.name}.bn.weight"] = (cur + 1, (1,))
            cur += 4
        head = next(op for op in plan.ops if isinstance(op, tuple) and op[0] == "head")
        add("last_conv1.weight", (plan.num_classes, head[2], 1, 1))
        self.off["last_conv1.bias"] = (self.off["last_conv1.weight"][0] + plan.num_classes * head[2],
                                       (plan.num_classes,))
        cur = self.off["last_conv1.bias"][0] + (plan.num_classes + 3) // 4 * 4
        self.total = cur

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        o, shape = self.off[name]
        n = 1
        for d in shape:
            n *= d
        return flat[o:o + n].view(shape)


def run_backward(plan: Plan, bplan: BackwardPlan, glayout: GradLayout, layout: Layout,
                 params: Dict[str, torch.Tensor], weights: PreparedWeights, arena: torch.Tensor,
                 dlogits: torch.Tensor, grad_flat: torch.Tensor, group: int = 0, on_done=None, sync_names=None) -> None:
    """Writes every parameter gradient into `grad_flat` (fp32, laid out by `glayout`).  `group`: which statistics group
    of a stacked-batch forward the backward runs through (the train forward rides as the LAST group of the stacked
    pseudo-label forward, AideTrainer); its images, raw conv outputs and BatchNorm statistics are slices of the tape.

    AIDE_B200_WGRAD_STREAM=1 runs the weight-gradient kernels on a side stream: wgrad(U) only needs dZ(U) and the taped
    input of U, nothing downstream needs it before the optimiser, so it can overlap the dgrad / BatchNorm backward chain of
    the next units -- two alternating dZ buffers (and dZ scale slots) keep a buffer alive until the wgrad that reads it is
    done.  OFF by default: measured 44.75 vs 44.40 ms per step in the same session (two A/B pairs) -- with the two networks
    already on two streams the GPU is throughput-bound under its power cap, extra concurrency only adds scheduling.  `on_done(name)` is called with the side stream joined when `name` is in `sync_names` (None: always):
    the gradients of everything the backward pass has visited so far are then final on the current stream."""
    H, W, fmt = layout.H, layout.W, layout.fmt
    G = layout.groups
    N = layout.N // G                                  # images of the group
    base = arena.data_ptr()
    st = _stream()
    gbase = grad_flat.data_ptr()

    def gptr(name):
        return gbase + glayout.off[name][0] * 4

    def aview(name, coff):
        """channel view of activation buffer `name`, advanced to the first image of the group"""
        p0, p1, ctot, co = _view(layout, base, name, coff)
        lvl = plan.bufs[name][0]
        skip = group * N * (H >> lvl) * (W >> lvl) * ctot * _esize(layout.buf_fmt(name))
        return p0 + skip, (p1 + skip if p1 is not None else None), ctot, co

    # ---- backward arena: dX per unit / upsample / head, shared scratch for g, dz, partials, wgrad workspace
    off: Dict[str, int] = {}
    cur = 0

    def reserve(key, nbytes):
        nonlocal cur
        off[key] = cur
        cur += _align(nbytes)

    max_g = max_dz = max_part = max_ws = 0
    for u in plan.units:
        h, w = H >> u.level, W >> u.level
        ufmt = FMT_F32 if u.first else fmt
        if not u.first:
            reserve("dx:" + u.name, N * h * w * u.cin * 4)
        if u.transposed:
            reserve("dk:" + u.name, u.cout * u.cin * 9 * 4)
        max_g = max(max_g, N * h * w * u.cout * 4)
        max_dz = max(max_dz, _align(N * h * w * u.cout * _esize(ufmt)) * _planes(ufmt))
        max_part = max(max_part, lib.aide_bn_bwd_rows(N, h, w, u.cout) * 2 * u.cout * 4)
        max_ws = max(max_ws, lib.aide_conv3x3_wgrad_workspace_bytes(ufmt, u.cin, u.cout, N, h, w))
    max_sa_pix = max_sa_rows = max_sa_ws = 0
    for a in plan.atts:
        h, w = H >> a.level, W >> a.level
        reserve("dy:" + a.name, N * h * w * a.c * 4)           # gradient w.r.t. the block's ungated output
        max_sa_pix = max(max_sa_pix, N * h * w)
        max_sa_rows = max(max_sa_rows, lib.aide_sa_bwd_rows(N, h, w, 1), lib.aide_sa_bwd_rows(N, h, w, 0))
        max_sa_ws = max(max_sa_ws, lib.aide_sa_bwd_workspace_floats(a.c, a.r, N, h, w))
    if plan.atts:
        reserve("sa_dahat", max_sa_pix * 4)
        reserve("sa_part", max_sa_rows * 2 * 4)
        reserve("sa_ws", max_sa_ws * 4)
    for op in plan.ops:
        if isinstance(op, Upsample):
            reserve("dlo:" + op.dst, N * (H >> op.level) * (W >> op.level) * op.c * 4)
        elif isinstance(op, tuple) and op[0] == "head":
            reserve("dx:head", N * H * W * op[2] * 4)
            hrows = lib.aide_conv1x1_bwd_rows(N, H, W, op[2])
            reserve("headpart", hrows * (plan.num_classes * op[2] + plan.num_classes) * 4)
    reserve("g", max_g)
    reserve("dz0", max_dz)
    reserve("dz1", max_dz)
    reserve("part", max_part)
    reserve("part2", max_part)
    reserve("ws", max_ws)
    reserve("gscale", 128)         # [0] max |g| (bits), [1..2] = {s, 1/s}: dynamic power-of-two scale of an F16X2 dZ;
                                   # words [16, 32): tickets of the one-launch reductions.  Zero on entry, left zero.
    barena = torch.empty(cur, dtype=torch.uint8, device=arena.device)
    bb = barena.data_ptr()
    barena[off["gscale"]:off["gscale"] + 128].zero_()
    main = torch.cuda.current_stream()
    side = _side_stream(main) if os.environ.get("AIDE_B200_WGRAD_STREAM", "0") == "1" else None
    busy = [None, None]            # event after the last wgrad that read dZ buffer k
    last_wgrad = None
    n_unit = 0

    def src_ptr(kind, obj):
        if kind == "unit":
            return bb + off["dx:" + obj.name]
        if kind == "att":
            return bb + off["dy:" + obj.name]
        if kind == "ups":
            return bb + off["dlo:" + obj.dst]
        return bb + off["dx:head"]

    for op in reversed(plan.ops):
        if isinstance(op, tuple) and op[0] == "head":
            _, name, cin = op
            x0, x1, xct, xco = aview(name, 0)
            call("aide_conv1x1_bwd", fmt, x0, x1, xct, xco, cin, params["last_conv1.weight"].data_ptr(),
                 dlogits.data_ptr(), plan.num_classes, N, H, W, bb + off["dx:head"], gptr("last_conv1.weight"),
                 bb + off["headpart"], st)
            if on_done is not None:
                on_done("head")
        elif isinstance(op, Upsample):
            # the (single) consumer of op.dst is the up-conv unit; its dX is the high-resolution gradient
            consumer = next(u for u in plan.units if u.src[0] == op.dst)
            h, w = H >> op.level, W >> op.level
            call("aide_upsample2x_bwd" if op.mode == "bilinear" else "aide_zero_insert2x_bwd",
                 bb + off["dx:" + consumer.name], consumer.cin, 0, bb + off["dlo:" + op.dst], N, h, w, op.c, st)
        elif isinstance(op, Attention):
            a = op
            h, w = H >> a.level, W >> a.level
            direct, pooled = bplan.sources[a.name]
            dptr = (C.c_void_p * 3)(*[src_ptr(k, o) for k, o, _, _ in direct])
            dct = (C.c_int * 3)(*[cc for _, _, _, cc in direct])
            dco = (C.c_int * 3)(*[co for _, _, co, _ in direct])
            pptr = (C.c_void_p * 3)(*[src_ptr(k, o) for k, o, _, _ in pooled])
            pct = (C.c_int * 3)(*[cc for _, _, _, cc in pooled])
            pco = (C.c_int * 3)(*[co for _, _, co, _ in pooled])
            y = aview(a.src[0], a.src[1])
            npix = N * h * w
            o = lambda k, per: base + layout.off[k + a.name] + group * npix * per * 4     # this group's slice
            mr = base + layout.off["sass:" + a.name] + (G + group) * 2 * 4                # mean_rstd [G][2] after scale_shift
            P = lambda k: params[f"{a.name}.{k}"].data_ptr()
            dy, dahat, part = bb + off["dy:" + a.name], bb + off["sa_dahat"], bb + off["sa_part"]
            call("aide_sa_bwd_gate", fmt, *y, a.c, o("sag:", 1), o("saa:", 1), mr, N, h, w, dptr, dct, dco, len(direct),
                 pptr, pct, pco, len(pooled), dy, dahat, part, st)
            call("aide_sa_bwd_chain", fmt, *y, a.c, a.r, a.dilation, P("conv1.weight"), P("conv2.weight"),
                 P("conv3.weight"), P("conv4.weight"), P("bn.weight"), o("sa1:", a.r), o("sa2:", a.r), o("sa3:", a.r),
                 o("saa:", 1), mr, dahat, part, lib.aide_sa_bwd_rows(N, h, w, 1 if pooled else 0), N, h, w,
                 bb + off["sa_ws"], max_sa_ws, dy, gptr(f"{a.name}.conv1.weight"), gptr(f"{a.name}.conv1.bias"),
                 gptr(f"{a.name}.conv2.weight"), gptr(f"{a.name}.conv2.bias"), gptr(f"{a.name}.conv3.weight"),
                 gptr(f"{a.name}.conv3.bias"), gptr(f"{a.name}.conv4.weight"), gptr(f"{a.name}.conv4.bias"),
                 gptr(f"{a.name}.bn.bias"), st)
            if on_done is not None:
                on_done(a.name)
        elif isinstance(op, Unit):
            u = op
            h, w = H >> u.level, W >> u.level
            ufmt = FMT_F32 if u.first else fmt      # first-layer units: CUDA-core wgrad, no dgrad -> fp32 dz
            direct, pooled = bplan.sources[u.name]
            dptr = (C.c_void_p * 3)(*[src_ptr(k, o) for k, o, _, _ in direct])
            dct = (C.c_int * 3)(*[cc for _, _, _, cc in direct])
            dco = (C.c_int * 3)(*[co for _, _, co, _ in direct])
            pptr = (C.c_void_p * 3)(*[src_ptr(k, o) for k, o, _, _ in pooled])
            pct = (C.c_int * 3)(*[cc for _, _, _, cc in pooled])
            pco = (C.c_int * 3)(*[co for _, _, co, _ in pooled])
            z = base + layout.off["z:" + u.name] + group * N * h * w * u.cout * 4
            ss = base + layout.off["ss:" + u.name] + group * 2 * u.cout * 4          # scale_shift [G][2][C] ...
            mr = base + layout.off["ss:" + u.name] + (G + group) * 2 * u.cout * 4    # ... then mean_rstd [G][2][C]
            g, part, part2 = bb + off["g"], bb + off["part"], bb + off["part2"]
            dyn = ufmt == FMT_F16X2
            k = n_unit & 1                          # alternate dZ buffer / scale slot {s, 1/s} at gscale words [1+2k, 2+2k]
            n_unit += 1
            gmax = bb + off["gscale"] if dyn else None
            dz_scale = bb + off["gscale"] + 4 + 8 * k if dyn else None
            dz_inv = bb + off["gscale"] + 8 + 8 * k if dyn else None
            if side is not None and busy[k] is not None:
                main.wait_event(busy[k])            # the wgrad two units back is done with this dZ buffer
            call("aide_bn_relu_bwd_reduce", z, ss, mr, N, h, w, u.cout, dptr, dct, dco, len(direct),
                 pptr, pct, pco, len(pooled), g, part, gmax, st)
            rows = lib.aide_bn_bwd_rows(N, h, w, u.cout)
            dz0 = bb + off["dz%d" % k]
            dz1 = dz0 + _align(N * h * w * u.cout * _esize(ufmt)) if _planes(ufmt) == 2 else None
            # the conv-bias gradient (column sums of part2) rides in the wgrad's split-K reduction launch -- unless the
            # wgrad runs on the side stream, where part2 could be overwritten by the next unit before it is read
            fuse_db = side is None or u.transposed
            call("aide_bn_relu_bwd_apply", ufmt, g, z, mr, params[u.bn + ".weight"].data_ptr(), part, rows,
                 N, h, w, u.cout, dz0, dz1, gptr(u.bn + ".weight"), gptr(u.bn + ".bias"),
                 None if fuse_db else gptr(u.conv + ".bias"), part2, gmax, dz_scale, bb + off["gscale"] + 64, st)
            x0, x1, xct, xco = aview(u.src[0], u.src[1])
            wst = st
            dw_dst = bb + off["dk:" + u.name] if u.transposed else gptr(u.conv + ".weight")
            if side is not None and u.transposed and last_wgrad is not None:
                main.wait_event(last_wgrad)          # this wgrad runs on the main stream: the shared workspace must be free
            if side is not None and not u.transposed:
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                wst = side.cuda_stream
            if fuse_db:
                call("aide_conv3x3_wgrad_ex", ufmt, x0, x1, xct, xco, u.cin, dz0, dz1, dz_inv, u.cout, N, h, w,
                     bb + off["ws"], max_ws, dw_dst, part2, rows, u.cout, gptr(u.conv + ".bias"), wst)
            else:
                call("aide_conv3x3_wgrad", ufmt, x0, x1, xct, xco, u.cin, dz0, dz1, dz_inv, u.cout, N, h, w,
                     bb + off["ws"], max_ws, dw_dst, wst)
            if u.transposed:                        # dW[ci,co,a,b] = dK[co,ci,1-a,1-b]: tensor re-indexing into the slot
                o0 = off["dk:" + u.name]
                dk = barena[o0:o0 + u.cout * u.cin * 36].view(torch.float32).view(u.cout, u.cin, 3, 3)
                glayout.view(grad_flat, u.conv + ".weight").copy_(transposed_weight_grad(dk))
            elif side is not None:
                busy[k] = last_wgrad = torch.cuda.Event()
                last_wgrad.record(side)
            if not u.first:
                w0, w1 = weights.dgrad(u, fmt)
                call("aide_conv3x3_dgrad", fmt, dz0, dz1, u.cout, w0, w1, dz_inv,
                     bb + off["dx:" + u.name], u.cin, 0, u.cin, N, h, w, st)
            if on_done is not None:
                if last_wgrad is not None and (sync_names is None or u.name in sync_names):
                    main.wait_event(last_wgrad)
                on_done(u.name)
    if last_wgrad is not None:
        main.wait_event(last_wgrad)                 # the optimiser (and the arena's next user) come after every wgrad


def gradient_buckets(plan: Plan, glayout: GradLayout, n_buckets: int = 4):
    """Split the flat gradient buffer into ranges that become final one after the other while the backward pass walks
    the network from the head to the first encoder level: [(trigger, [(lo, hi), ...]), ...] in completion order --
    after the unit / gate `trigger` is done, every listed float range holds final gradients.  The flat layout follows
    the forward order of the units, so the ranges of the units are suffixes cut at level boundaries (the decoder's
    ~16 M and the deepest encoder level's ~8 M parameters are ready before the expensive high-resolution encoder
    levels run their backward); attention-gate and head parameters live behind the units and travel with the LAST bucket
    (the first level's gate finishes just before the end)."""
    units = plan.units
    first_off = {u.name: glayout.off[u.conv + ".weight"][0] for u in units}
    tail_start = min([glayout.off[f"{a.name}.conv1.weight"][0] for a in plan.atts] + [glayout.off["last_conv1.weight"][0]])
    # cut points: the first unit (in forward order) of the decoder and of the deepest encoder levels
    dec0 = next(u for u in units if u.name.startswith("up_block1"))
    cuts = [dec0]
    level_first = {}
    for u in units:
        if not u.name.startswith("up_block"):
            level_first.setdefault(u.level, u)
    for lvl in sorted(level_first, reverse=True)[:max(n_buckets - 2, 0)]:
        cuts.append(level_first[lvl])
    out, hi = [], tail_start
    for u in cuts:
        lo = first_off[u.name]
        if lo < hi:
            out.append((u.name, [(lo, hi)]))
            hi = lo
    out.append((units[0].name, [(0, hi), (tail_start, glayout.total)]))
    return out
