"""Drop-in nn.Module surface of the reference networks, executed by the B200 engine.

``fuseunet`` mirrors models_twomodalinputs/fuseunet.py:6-91 and ``UNet`` mirrors
models_singlemodalinput/UNet.py:135-165: same constructor signatures, same module tree (hence the
same ``state_dict()`` keys/shapes/order and -- because nn.Conv2d / nn.BatchNorm2d are instantiated
in the reference's order -- the same random initialisation under the same torch seed), same
``forward`` signatures.  The sub-modules only *hold* parameters and buffers; ``forward`` hands
their storage to the CUDA engine (aide_b200.engine) through one autograd.Function per network.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import engine as E


# -------------------------------------------------------------------------------------------------
# parameter containers (names follow netblocks.py:9-33,128-147 and UNet.py:4-28,110-133)
# -------------------------------------------------------------------------------------------------
class basic_block(nn.Module):
    def __init__(self, input_channel: int, output_channel: int):
        super().__init__()
        self.conv1 = nn.Conv2d(input_channel, output_channel, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(output_channel)
        self.conv2 = nn.Conv2d(output_channel, output_channel, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(output_channel)
        self.relu = nn.ReLU()


def UNet_up_conv_bn_relu(input_channel: int, output_channel: int, learned_bilinear: bool = False) -> nn.Sequential:
    if learned_bilinear:                              # netblocks.py:11-14 / UNet.py:6-9
        return nn.Sequential(nn.ConvTranspose2d(input_channel, output_channel, kernel_size=2, stride=2),
                             nn.BatchNorm2d(output_channel),
                             nn.ReLU())
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                         nn.Conv2d(input_channel, output_channel, kernel_size=3, padding=1),
                         nn.BatchNorm2d(output_channel),
                         nn.ReLU())


class Spatial_Attention(nn.Module):
    """Parameter container of netblocks.py:68-80 / UNet.py:85-97 (same tensors, same creation order)."""

    def __init__(self, input_channel: int, reduction: int = 16, dilation: int = 4):
        super().__init__()
        r = input_channel // reduction
        self.conv1 = nn.Conv2d(input_channel, r, kernel_size=1, stride=1, padding=0)
        self.conv2 = nn.Conv2d(r, r, kernel_size=3, dilation=dilation, stride=1, padding=dilation)
        self.conv3 = nn.Conv2d(r, r, kernel_size=3, dilation=dilation, stride=1, padding=dilation)
        self.conv4 = nn.Conv2d(r, 1, kernel_size=1, stride=1, padding=0)
        self.bn = nn.BatchNorm2d(1)
        self.sigmoid = nn.Sigmoid()


class UNet_basic_down_block(nn.Module):
    def __init__(self, input_channel: int, output_channel: int, down_size: Optional[bool] = None):
        super().__init__()
        self.block = basic_block(input_channel, output_channel)
        if down_size is not None:                     # single-modal flavour (UNet.py:110-121)
            self.max_pool = nn.MaxPool2d(2, 2)
            self.down_size = down_size


class UNet_basic_up_block(nn.Module):
    def __init__(self, input_channel: int, prev_channel: int, output_channel: int, learned_bilinear: bool = False):
        super().__init__()
        self.bilinear_up = UNet_up_conv_bn_relu(input_channel, prev_channel, learned_bilinear)
        self.block = basic_block(prev_channel * 2, output_channel)


# -------------------------------------------------------------------------------------------------
# autograd bridge
# -------------------------------------------------------------------------------------------------
class _NetFunction(torch.autograd.Function):
    """logits = net(inputs).  Saved state is the engine's activation arena (the 'tape')."""

    @staticmethod
    def forward(ctx, net, n_in, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        logits, tape = net._engine_forward(inputs, keep_tape=True)
        ctx.net, ctx.tape, ctx.n_in, ctx.n_params = net, tape, n_in, len(params)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        grads = ctx.net._engine_backward(ctx.tape, dlogits)
        return (None, None) + (None,) * ctx.n_in + tuple(grads)


class _EngineNet(nn.Module):
    """Shared host logic of fuseunet / UNet."""

    _plan: E.Plan

    def _engine_setup(self, plan: E.Plan, mode: Optional[str]):
        self._plan = plan
        self._bplan = E.BackwardPlan(plan)
        self._glayout = E.GradLayout(plan)
        self.engine_mode = mode or E.default_mode()
        if self.engine_mode not in E.MODES:
            raise ValueError(f"mode must be one of {sorted(E.MODES)}")
        self._layouts: Dict[tuple, E.Layout] = {}
        self._weights: Dict[int, E.PreparedWeights] = {}     # per operand format
        self._tensors: Optional[Dict[str, torch.Tensor]] = None
        self._scratch: Optional[torch.Tensor] = None
        self.last_grad_flat: Optional[torch.Tensor] = None
        self._grad_flat_static: Optional[torch.Tensor] = None
        self._auto_graph = os.environ.get("AIDE_B200_AUTOGRAPH", "1") != "0" and os.environ.get("AIDE_B200_GRAPH", "1") != "0"
        self._auto_graph_max = int(os.environ.get("AIDE_B200_AUTOGRAPH_MAX", "4"))
        self._auto_graphs: Dict[tuple, object] = {}       # key -> False (seen once, eager) | _GraphedForward
        self._ticket_off, self._ticket_total = E.ticket_offsets(plan)
        self._tickets: Optional[torch.Tensor] = None          # zero-initialised once per device, self-resetting

    # ---- bookkeeping ---------------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._tensors, self._weights, self._scratch, self._tickets = None, {}, None, None
        self._auto_graphs = {}
        return out

    def _named(self) -> Dict[str, torch.Tensor]:
        if self._tensors is None:
            t = dict(self.named_parameters())
            t.update(dict(self.named_buffers()))
            self._tensors = t
        return self._tensors

    def _param_order(self):
        return [p for _, p in self.named_parameters()]

    def _check_inputs(self, inputs):
        x0 = inputs[0]
        for x in inputs:
            if not x.is_cuda:
                raise RuntimeError("aide_b200 runs on CUDA only (there is no CPU fallback); move inputs to the GPU")
            if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
                raise ValueError(f"expected fp32 [N,3,H,W] inputs, got {tuple(x.shape)} {x.dtype}")
            if x.shape != x0.shape:
                raise ValueError("both modalities must have the same shape")
        p = next(self.parameters())
        if p.device != x0.device:
            raise RuntimeError(f"module is on {p.device}, inputs on {x0.device}")

    # ---- engine calls --------------------------------------------------------------------------
    def _engine_forward_grouped(self, views):
        """`views`: list of input tuples of identical shape (the augmented views of one AIDE step).  Equivalent to
        [self._engine_forward(v, keep_tape=False)[0] for v in views] -- same BatchNorm statistics per view, same
        order of running-statistics updates -- but every convolution runs once on the stacked batch (more tiles per
        launch: better SM occupancy on the deep, small feature maps; a quarter of the launches).  Returns the list of
        per-view logits (views of one [G*B,K,H,W] tensor)."""
        G = len(views)
        if G == 1:
            return [self._engine_forward(views[0], keep_tape=False)[0]]
        for v in views:
            self._check_inputs(v)
            if v[0].shape != views[0][0].shape:
                raise ValueError("grouped forward needs identically shaped views")
        logits, _ = self._engine_forward(views, keep_tape=False, groups=G)
        return list(logits.chunk(G, dim=0))

    def graphed_eval(self, *example_inputs):
        """A CUDA-graph replay of the eval-mode forward for inputs shaped like `example_inputs` ("next" row f3 of
        SURVEY.md section 8: the per-epoch case loops of trainchaos_proposed_30cases1labeled.py:373-496 and
        evalchaos_comparison_1cases.py:203-214 run thousands of single-slice forwards; eagerly each one is ~100
        launches of ~50 us host time).  Returns a callable `f(*inputs) -> logits`; the logits tensor is the graph's
        static output (overwritten by the next call).  The graph is re-captured when a weight changes."""
        return _GraphedEval(self, example_inputs)

    def _engine_forward(self, inputs, keep_tape: bool, groups: int = 1, arena: Optional[torch.Tensor] = None):
        plan = self._plan
        fmt = E.mode_format(self.engine_mode, keep_tape)
        first = inputs[0][0] if groups > 1 else inputs[0]
        N, _, H, W = first.shape
        N *= groups
        key = (N, H, W, fmt, groups)
        layout = self._layouts.get(key)
        if layout is None:
            layout = self._layouts[key] = E.Layout(plan, N, H, W, fmt, groups)
        named = self._named()
        training = self.training
        wkey = tuple((named[u.conv + ".weight"].data_ptr(), named[u.conv + ".weight"]._version) for u in plan.units)
        need_dgrad = keep_tape or (getattr(self, "_prep_dgrad_always", False)
                                   and fmt == E.mode_format(self.engine_mode, True))
        if self._weights is None:
            self._weights = {}
        wts = self._weights.get(fmt)
        if wts is None or wts.key != wkey or (need_dgrad and not wts.has_dgrad):
            wts = self._weights[fmt] = E.PreparedWeights(plan, named, fmt, need_dgrad)
        dev = first.device
        if arena is not None:
            if arena.numel() < layout.total:
                raise ValueError("arena too small for this forward")
        elif keep_tape:
            arena = torch.empty(layout.total, dtype=torch.uint8, device=dev)
        else:
            if self._scratch is None or self._scratch.numel() < layout.total or self._scratch.device != dev:
                self._scratch = torch.empty(layout.total, dtype=torch.uint8, device=dev)
            arena = self._scratch
        logits = torch.empty((N, plan.num_classes, H, W), dtype=torch.float32, device=dev)
        if groups > 1:
            xs = [[x.detach().contiguous() for x in v] for v in inputs]
        else:
            xs = [x.detach().contiguous() for x in inputs]
        if self._tickets is None or self._tickets.device != dev:
            self._tickets = torch.zeros(max(self._ticket_total, 1), dtype=torch.int32, device=dev)
        E.run_forward(plan, layout, named, wts, xs, training, logits, arena, groups, (self._tickets, self._ticket_off))
        if training:
            torch._foreach_add_([named[u.bn + ".num_batches_tracked"] for u in plan.units]
                                + [named[a.name + ".bn.num_batches_tracked"] for a in plan.atts], groups)
        tape = None
        if keep_tape:
            tape = E.Tape()
            tape.layout, tape.arena, tape.weights, tape.training = layout, arena, wts, training
            tape.group = groups - 1          # a stacked forward with a tape: the backward runs through the LAST group
        return logits, tape

    def _engine_backward(self, tape, dlogits, on_done=None, sync_names=None):
        if not tape.training:
            raise NotImplementedError("backward through an eval-mode (running-statistics) BatchNorm forward is not "
                                      "supported; the reference never does this")
        named = self._named()
        # zeros, not empty: the 4-element alignment gaps between tensors are part of the flat Adam / all-reduce buffer
        if self._grad_flat_static is not None:       # a persistent (peer-mapped) buffer supplied by AideTrainer
            flat = self._grad_flat_static
            flat.zero_()
        else:
            flat = torch.zeros(self._glayout.total, dtype=torch.float32, device=dlogits.device)
        self.last_grad_flat = flat
        E.run_backward(self._plan, self._bplan, self._glayout, tape.layout, named, tape.weights, tape.arena,
                       dlogits.contiguous(), flat, tape.group, on_done, sync_names)
        return [self._glayout.view(flat, n) for n, _ in self.named_parameters()]

    def _run(self, *inputs):
        self._check_inputs(inputs)
        params = self._param_order()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _NetFunction.apply(self, len(inputs), *inputs, *params)
        if self._auto_graph and not torch.cuda.is_current_stream_capturing():
            return self._nograd_forward_graphed(inputs)
        logits, _ = self._engine_forward(inputs, keep_tape=False)
        return logits

    def _nograd_forward_graphed(self, inputs):
        """No-grad forwards of an unmodified reference script -- the 8 pseudo-label forwards of every training step
        (trainchaos_proposed_30cases1labeled.py:263-272, train-mode BatchNorm) and the thousands of single-slice
        evaluation forwards per epoch (:373-496) -- are ~100-170 launches of ~30 us host time each when issued eagerly.
        The second call with the same input shape / BatchNorm mode captures the forward INCLUDING its weight preparation
        into a CUDA graph; later calls copy the inputs into the graph's static buffers, replay, and return a copy of the
        logits.  In-place parameter updates (torch.optim) need no re-capture: the replay re-derives the operand planes
        from the parameters' current values.  AIDE_B200_AUTOGRAPH=0 disables it."""
        named = self._named()
        w0 = named[self._plan.units[0].conv + ".weight"]
        key = (tuple(inputs[0].shape), len(inputs), self.training, id(named), w0.data_ptr(), inputs[0].device.index)
        ent = self._auto_graphs.get(key)
        if ent is None:                                   # first sight: eager (this is also the capture's warm-up)
            self._auto_graphs[key] = False
            while len(self._auto_graphs) > self._auto_graph_max:
                self._auto_graphs.pop(next(iter(self._auto_graphs)))
            logits, _ = self._engine_forward(inputs, keep_tape=False)
            return logits
        if ent is False:
            ent = self._auto_graphs[key] = _GraphedForward(self, inputs)
        else:
            self._auto_graphs[key] = self._auto_graphs.pop(key)     # most recently used last
        return ent(*inputs)


class _GraphedForward:
    """Captured no-grad forward of one network for one input shape and BatchNorm mode, weight preparation included
    (see _EngineNet._nograd_forward_graphed).  Must be constructed right after an eager forward of the same shape."""

    def __init__(self, net: "_EngineNet", inputs):
        self.static_in = [torch.empty_like(x) for x in inputs]
        N, _, H, W = inputs[0].shape
        fmt = E.mode_format(net.engine_mode, False)
        total = net._layouts[(N, H, W, fmt, 1)].total
        self._arena = torch.empty(total, dtype=torch.uint8, device=inputs[0].device)      # owned by the graph
        saved = net._weights
        net._weights = {}                                 # force the weight preparation INTO the graph ...
        self.graph = torch.cuda.CUDAGraph()
        n0 = E.lib.aide_launch_count()
        try:
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.out, _ = net._engine_forward(self.static_in, keep_tape=False, arena=self._arena)
            self._planes = net._weights                   # ... whose operand planes belong to the graph's memory pool
        finally:
            net._weights = saved                          # eager paths never alias graph-owned planes
        self.n_kernels = E.lib.aide_launch_count() - n0

    def __call__(self, *inputs):
        for d, s_ in zip(self.static_in, inputs):
            d.copy_(s_, non_blocking=True)
        self.graph.replay()
        return self.out.clone()


class _GraphedEval:
    """Captured eval-mode forward of one network for one input shape (see _EngineNet.graphed_eval)."""

    def __init__(self, net: "_EngineNet", example_inputs):
        if net.training:
            raise RuntimeError("graphed_eval() is for eval mode: call net.eval() first")
        net._check_inputs(example_inputs)
        self.net = net
        self.static_in = [torch.empty_like(x) for x in example_inputs]
        self.graph = None
        self.out = None
        self._wkey = None
        self._arena = None

    def _weights_key(self):
        named = self.net._named()
        return tuple((named[u.conv + ".weight"].data_ptr(), named[u.conv + ".weight"]._version)
                     for u in self.net._plan.units)

    def _capture(self):
        net = self.net
        with torch.no_grad():
            out, _ = net._engine_forward(self.static_in, keep_tape=False)      # warm-up: weight planes, layouts
            torch.cuda.synchronize()
            N, _, H, W = self.static_in[0].shape
            fmt = E.mode_format(net.engine_mode, False)
            total = net._layouts[(N, H, W, fmt, 1)].total
            self._arena = torch.empty(total, dtype=torch.uint8, device=self.static_in[0].device)   # owned by the graph
            self.graph = torch.cuda.CUDAGraph()
            n0 = E.lib.aide_launch_count()
            with torch.cuda.graph(self.graph):
                self.out, _ = net._engine_forward(self.static_in, keep_tape=False, arena=self._arena)
            self.n_kernels = E.lib.aide_launch_count() - n0          # kernels of this library per replay
        self._wkey = self._weights_key()
        self._weights = net._weights.get(fmt)          # keep the operand planes the captured kernels read alive

    def __call__(self, *inputs):
        if self.net.training:
            raise RuntimeError("the captured forward is an eval-mode forward")
        if self.graph is None or self._weights_key() != self._wkey:
            self._capture()
        for d, s_ in zip(self.static_in, inputs):
            d.copy_(s_, non_blocking=True)
        self.graph.replay()
        return self.out


# -------------------------------------------------------------------------------------------------
# the two networks
# -------------------------------------------------------------------------------------------------
class fuseunet(_EngineNet):
    """Two-encoder / one-decoder U-Net (models_twomodalinputs/fuseunet.py:6-91)."""

    def __init__(self, num_classes=2, reduction=16, dilation=4, learned_bilinear=False, mode: Optional[str] = None):
        super().__init__()
        enc = [(3, 3, 32), (64, 32, 64), (128, 64, 128), (256, 128, 256), (512, 256, 512)]
        for i, (c1, _, co) in enumerate(enc, 1):
            setattr(self, f"modal1_downblock{i}", UNet_basic_down_block(c1, co))
            if i < 5:
                setattr(self, f"modal1_maxpool{i}", nn.MaxPool2d(kernel_size=2, stride=2))
        for i, (_, c2, co) in enumerate(enc, 1):
            setattr(self, f"modal2_downblock{i}", UNet_basic_down_block(c2, co))
            if i < 5:
                setattr(self, f"modal2_maxpool{i}", nn.MaxPool2d(kernel_size=2, stride=2))
        self.up_block1 = UNet_basic_up_block(1024, 512, 512, learned_bilinear)
        self.up_block2 = UNet_basic_up_block(512, 256, 256, learned_bilinear)
        self.up_block3 = UNet_basic_up_block(256, 128, 128, learned_bilinear)
        self.up_block4 = UNet_basic_up_block(128, 64, 64, learned_bilinear)
        self.last_conv1 = nn.Conv2d(64, num_classes, 1, padding=0)
        self._engine_setup(E.plan_fuseunet(num_classes, learned_bilinear), mode)

    def forward(self, modal1_inputs, modal2_inputs):
        return self._run(modal1_inputs, modal2_inputs)


class UNet(_EngineNet):
    """Classic 5-level U-Net 3->64->...->1024 (models_singlemodalinput/UNet.py:135-165).  `_base` is the width of the
    first level: the reference's UNet128 / UNet32 / UNet16 / UNet8 / UNet4 (UNet.py:210-368) are the same network at
    other widths (subclasses below)."""
    _base = 64

    def __init__(self, num_classes=2, learned_bilinear=False, mode: Optional[str] = None):
        super().__init__()
        b = self._base
        self.down_block1 = UNet_basic_down_block(3, b, False)
        self.down_block2 = UNet_basic_down_block(b, 2 * b, True)
        self.down_block3 = UNet_basic_down_block(2 * b, 4 * b, True)
        self.down_block4 = UNet_basic_down_block(4 * b, 8 * b, True)
        self.down_block5 = UNet_basic_down_block(8 * b, 16 * b, True)
        self.up_block1 = UNet_basic_up_block(16 * b, 8 * b, 8 * b, learned_bilinear)
        self.up_block2 = UNet_basic_up_block(8 * b, 4 * b, 4 * b, learned_bilinear)
        self.up_block3 = UNet_basic_up_block(4 * b, 2 * b, 2 * b, learned_bilinear)
        self.up_block4 = UNet_basic_up_block(2 * b, b, b, learned_bilinear)
        self.last_conv1 = nn.Conv2d(b, num_classes, 1, padding=0)
        if b % 32 and b % 4 == 0:
            # channel counts below the tcgen05 tile granularity (32): every layer runs on the fp32 CUDA-core kernels
            mode = "exact"
        elif b % 4:
            raise NotImplementedError(f"UNet{b}: the NHWC kernels move 4 channels per access; widths below 4 are not built")
        self._engine_setup(E.plan_unet(num_classes, learned_bilinear, base=b), mode)

    def forward(self, x):
        return self._run(x)


class UNet128(UNet):
    _base = 128


class UNet32(UNet):
    _base = 32


class UNet16(UNet):
    _base = 16


class UNet8(UNet):
    _base = 8


class UNet4(UNet):
    _base = 4


class UNet2(UNet):
    _base = 2


class _fuseunetsa_base(_EngineNet):
    def __init__(self, num_classes, reduction, dilation, learned_bilinear, mode, separate):
        super().__init__()
        enc = [(3, 3, 32), (64, 32, 64), (128, 64, 128), (256, 128, 256), (512, 256, 512)]
        for m in (1, 2):
            for i, (c1, c2, co) in enumerate(enc, 1):
                cin = c2 if (m == 2 or separate) else c1
                setattr(self, f"modal{m}_downblock{i}", UNet_basic_down_block(cin, co))
                setattr(self, f"modal{m}_sa{i}", Spatial_Attention(co, reduction=reduction, dilation=dilation))
                if i < 5:
                    setattr(self, f"modal{m}_maxpool{i}", nn.MaxPool2d(kernel_size=2, stride=2))
        self.up_block1 = UNet_basic_up_block(1024, 512, 512, learned_bilinear)
        self.up_block2 = UNet_basic_up_block(512, 256, 256, learned_bilinear)
        self.up_block3 = UNet_basic_up_block(256, 128, 128, learned_bilinear)
        self.up_block4 = UNet_basic_up_block(128, 64, 64, learned_bilinear)
        self.last_conv1 = nn.Conv2d(64, num_classes, 1, padding=0)
        self._engine_setup(E.plan_fuseunet(num_classes, learned_bilinear, attention=True, separate=separate,
                                           reduction=reduction, dilation=dilation), mode)

    def forward(self, modal1_inputs, modal2_inputs):
        return self._run(modal1_inputs, modal2_inputs)


class fuseunetsa(_fuseunetsa_base):
    """fuseunet with a Spatial_Attention gate after every encoder block (models_twomodalinputs/fuseunet.py:93-208)."""

    def __init__(self, num_classes=2, reduction=16, dilation=4, learned_bilinear=False, mode: Optional[str] = None):
        super().__init__(num_classes, reduction, dilation, learned_bilinear, mode, separate=False)


class fuseunetsaseparate(_fuseunetsa_base):
    """Two fully separate gated encoders, fused only in the decoder's skips (fuseunet.py:210-325)."""

    def __init__(self, num_classes=2, reduction=16, dilation=4, learned_bilinear=False, mode: Optional[str] = None):
        super().__init__(num_classes, reduction, dilation, learned_bilinear, mode, separate=True)


class UNetsa(_EngineNet):
    """UNet with a Spatial_Attention gate after every down block (models_singlemodalinput/UNet.py:168-208)."""

    def __init__(self, num_classes=2, learned_bilinear=False, mode: Optional[str] = None):
        super().__init__()
        for i, (ci, co) in enumerate([(3, 64), (64, 128), (128, 256), (256, 512), (512, 1024)], 1):
            setattr(self, f"down_block{i}", UNet_basic_down_block(ci, co, i > 1))
            setattr(self, f"sa{i}", Spatial_Attention(co, reduction=16, dilation=4))
        self.up_block1 = UNet_basic_up_block(1024, 512, 512, learned_bilinear)
        self.up_block2 = UNet_basic_up_block(512, 256, 256, learned_bilinear)
        self.up_block3 = UNet_basic_up_block(256, 128, 128, learned_bilinear)
        self.up_block4 = UNet_basic_up_block(128, 64, 64, learned_bilinear)
        self.last_conv1 = nn.Conv2d(64, num_classes, 1, padding=0)
        self._engine_setup(E.plan_unet(num_classes, learned_bilinear, attention=True), mode)

    def forward(self, x):
        return self._run(x)
