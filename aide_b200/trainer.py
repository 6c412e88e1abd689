"""The AIDE training step on the B200 engine -- host orchestration only.

Mirrors the inner loop of train_files/trainchaos_proposed_30cases1labeled.py:263-325 (kidney/breast flavours:
trainkidney_proposed_mask1.py:267-333, trainbreast_dataset3_proposed_272cases25labeled.py:304):

    4 augmented forwards per net (detached)          :265-269   train-mode BN (chaos) / eval-mode BN (kidney)
    softmax, mean, sharpen, weight map               :274-292   -> aide_pseudo_label (one kernel per net)
    train forwards of both nets                      :301-302
    per-image CE+Dice, cross small-loss selection    :303-321   -> aide_loss_sums / aide_coteach_select
    loss1.backward(); opt1.step(); loss2.backward(); opt2.step()   :322-325
                                                     -> aide_loss_bwd, engine backward, aide_adam_amsgrad

The two networks are independent until the selection needs the other net's per-image losses, so they run on
two CUDA streams with two event joins (the reference runs them back to back on one stream).  Data-parallel
training: every rank holds both nets and a local batch; BatchNorm statistics and the small-loss
selection are local; the flat fp32 gradient of each net is all-reduced (mean) once per step -- net-1's
all-reduce overlaps net-2's backward because they sit on different streams.

Nothing here computes on the CPU; all device work goes through the C ABI.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import augment as AUG
from . import engine as E
from . import losses as L
from ._lib import call, lib
from .nets import UNet, UNetsa, fuseunet, fuseunetsa, fuseunetsaseparate
from .optim import FlatAdamAMSGrad


def grad_order_params(net) -> List[torch.nn.Parameter]:
    """Parameters in the order of the engine's flat gradient buffer (engine.GradLayout)."""
    named = dict(net.named_parameters())
    order = sorted(net._glayout.off.items(), key=lambda kv: kv[1][0])
    return [named[name] for name, _ in order]


def flat_adam_for(net, lr: float) -> FlatAdamAMSGrad:
    """Adam-amsgrad whose flat parameter buffer has exactly the layout of the net's flat gradient buffer."""
    order = sorted(net._glayout.off.items(), key=lambda kv: kv[1][0])
    named = dict(net.named_parameters())
    return FlatAdamAMSGrad([named[name] for name, _ in order], lr=lr, offsets=[o for _, (o, _) in order],
                           total=net._glayout.total)


NET_KINDS = {"fuseunet": fuseunet, "unet": UNet, "fuseunetsa": fuseunetsa, "fuseunetsaseparate": fuseunetsaseparate,
             "unetsa": UNetsa}


class AideTrainer:
    """Two co-trained networks + their Adam(amsgrad) state + the fused AIDE step.

    kind: 'fuseunet' (two modalities) or 'unet' (one).  flavour: 'chaos' (train-mode augmented forwards,
    sharpen = pow(T)) or 'kidney' (eval-mode augmented forwards, sharpen = pow(1/T)).
    """

    def __init__(self, kind: str = "fuseunet", mode: Optional[str] = None, device="cuda:0", seed: int = 2,
                 lr: float = 1e-4, n_clean: int = 2, segcor_weight=(1.0, 10.0), temperature: float = 1.0,
                 flavour: str = "chaos", two_streams: bool = True, process_group=None,
                 cuda_graph: Optional[bool] = None, global_select: bool = False, max_graphs: int = 4,
                 data_parallel: bool = True, comm: Optional[str] = None, net_kwargs: Optional[Dict] = None):
        self.device = torch.device(device)
        self.kind, self.flavour, self.temperature = kind, flavour, temperature
        self.n_clean, self.segcor_weight = n_clean, segcor_weight
        torch.manual_seed(seed)                      # net1 then net2: independent inits, as in :175-176
        if kind not in NET_KINDS:
            raise ValueError(f"kind must be one of {sorted(NET_KINDS)}")          # cf. 'Model not implemented', :78
        ctor = NET_KINDS[kind]
        kw = dict(net_kwargs or {})                  # e.g. learned_bilinear=True (netblocks.py:11-14), reduction / dilation
        self.net1 = ctor(num_classes=2, mode=mode, **kw).to(self.device).train()
        self.net2 = ctor(num_classes=2, mode=mode, **kw).to(self.device).train()
        self.opt1 = flat_adam_for(self.net1, lr)
        self.opt2 = flat_adam_for(self.net2, lr)
        for net, opt in ((self.net1, self.opt1), (self.net2, self.opt2)):
            if opt.flat.numel() != net._glayout.total:
                raise RuntimeError("flat parameter buffer and flat gradient layout disagree")
            net._tensors = None                       # parameters were re-pointed into the flat buffer
            net._prep_dgrad_always = True             # one weight preparation per step serves all 5 forwards + dgrad
        self.group = process_group
        self.world, self.rank = 1, 0
        if data_parallel and (process_group is not None or
                              (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
            self.rank = torch.distributed.get_rank(process_group)
        # Data-parallel selection semantics.  False (default): every rank selects n_clean "clean" images inside its
        # local batch and the gradients are averaged.  True: the reference's nn.DataParallel behaviour
        # (trainchaos_proposed_30cases1labeled.py:183-186,303-310): the per-image losses of all ranks are all-gathered,
        # the sort and the n_clean / rest split are GLOBAL, loss means are over the global batch, gradients are summed.
        self.global_select = bool(global_select) and self.world > 1
        # warm-up rate (:248) as a device scalar: it changes every epoch of the warm-up and must not re-capture the graph
        self.rate_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._rate_uploaded = None
        self.max_graphs = max_graphs
        # gradient all-reduce in buckets launched while the backward pass is still running (decoder first); 0 / 1 = one
        # all-reduce per net after its backward
        self.n_buckets = int(os.environ.get("AIDE_B200_BUCKETS", "4"))
        # diagnostic only (tools/sessions/s19.sh): skip the gradient all-reduce to time the slowest rank's compute alone --
        # the replicas diverge, never use it for training
        self._comm = os.environ.get("AIDE_B200_NO_COMM", "0") != "1"
        self._sync_only = os.environ.get("AIDE_B200_COMM_SYNC_ONLY", "0") == "1"    # diagnostic, like NO_COMM
        self._buckets = {id(n): {t: r for t, r in E.gradient_buckets(n._plan, n._glayout, self.n_buckets)}
                         for n in (self.net1, self.net2)}
        # Gradient all-reduce: "p2p" = aide_allreduce_p2p over NVLink peer memory (csrc/comm.cu; its CTAs share the SMs
        # with the tensor-core kernels), "nccl" = torch.distributed.all_reduce.  p2p needs the flat gradient buffers in
        # IPC-exported allocations: they are created here, once, and the backward pass writes into them.
        self.comm_kind = (comm or os.environ.get("AIDE_B200_COMM", "p2p")) if self.world > 1 else "none"
        self._peer = None
        if self.comm_kind not in ("p2p", "nccl", "none"):
            raise ValueError("comm must be 'p2p' or 'nccl'")
        if self.world > 1 and self.comm_kind == "p2p" and self._comm:
            from .comm import PeerBuffers
            try:            # PeerBuffers raises on EVERY rank if any rank could not allocate / map (it agrees on the verdict)
                self._peer = PeerBuffers(self.group, self.device, [self.net1._glayout.total, self.net2._glayout.total],
                                         blocks=int(os.environ.get("AIDE_B200_COMM_BLOCKS", "32")))
            except RuntimeError as e:
                # NCCL is a GPU collective too: a choice between two device paths, announced loudly, never a CPU path
                import warnings
                warnings.warn(f"peer-memory all-reduce unavailable ({e}); using NCCL")
                self._peer, self.comm_kind = None, "nccl"
            else:
                for i, net in enumerate((self.net1, self.net2)):
                    net._grad_flat_static = self._peer.tensors[i][:net._glayout.total]
                self._comm_streams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
        # run the 4 augmented forwards of a net as ONE stacked-batch forward with per-view BatchNorm statistics
        self.group_augs = os.environ.get("AIDE_B200_GROUP_AUGS", "1") != "0"
        # ... and the train forward as one more group of that stacked forward (train-mode views only, i.e. the chaos flavour)
        self.stack_train = os.environ.get("AIDE_B200_STACK_TRAIN", "1") != "0"
        self.two_streams = two_streams
        if two_streams:
            self.s1, self.s2 = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.steps = 0
        # One AIDE step is ~1 800 kernel launches; replaying it as ONE captured CUDA graph removes the per-launch
        # host cost (ctypes marshalling + tensor-map encoding, ~50 us each) that otherwise bounds the step.
        if cuda_graph is None:
            cuda_graph = os.environ.get("AIDE_B200_GRAPH", "1") != "0"
        self.cuda_graph = bool(cuda_graph)
        self._graphs: Dict[tuple, tuple] = {}
        self._aug_dev = None             # (matrices, modes, flips) of the current step's reverse augmentation, or None
        self._staging: Dict[tuple, list] = {}        # device staging buffers of step_from_host(prefetch=...)
        self._staged = None                          # (host batch, its device copy, event) started by the previous call
        self._copy_stream = None
        self._consumed = None
        self.graph_launches = 0          # kernels launched through graph replays (aide_launch_count() sees eager ones)

    # ------------------------------------------------------------------------------------------
    # checkpoint / resume.  The reference saves {'net': state_dict, 'loss', 'epoch'} per network
    # (trainchaos_proposed_30cases1labeled.py:504-526) and never the optimiser; state_dict() holds both so that a resumed
    # run continues bit for bit.
    def state_dict(self) -> Dict:
        return {"net1": self.net1.state_dict(), "net2": self.net2.state_dict(), "opt1": self.opt1.state_dict(),
                "opt2": self.opt2.state_dict(), "steps": self.steps}

    def load_state_dict(self, sd: Dict) -> None:
        torch.cuda.synchronize(self.device)
        self.net1.load_state_dict(sd["net1"])        # in place: the parameters stay views of the flat buffers
        self.net2.load_state_dict(sd["net2"])
        self.opt1.load_state_dict(sd["opt1"])
        self.opt2.load_state_dict(sd["opt2"])
        self.steps = int(sd.get("steps", 0))
        for net, opt in ((self.net1, self.opt1), (self.net2, self.opt2)):
            net._weights = None                       # operand-format weight planes follow the new values
            opt.bump_versions()

    def save_reference_checkpoints(self, path1: str, path2: str, epoch: int, loss1=None, loss2=None) -> None:
        """The two files the reference writes for its best epoch (:507-526); loadable by the unmodified scripts."""
        for net, path, loss in ((self.net1, path1, loss1), (self.net2, path2, loss2)):
            torch.save({"net": net.state_dict(), "loss": loss, "epoch": epoch}, path)

    # ------------------------------------------------------------------------------------------
    def _state_tensors(self) -> List[torch.Tensor]:
        out = []
        for opt in (self.opt1, self.opt2):
            out += [opt.flat, opt.m, opt.v, opt.vmax, opt.step_dev, opt.bc_dev]
        for net in (self.net1, self.net2):
            out += list(net.buffers())
        return out

    def _capture(self, key, x, t1, t2, augs, rate):
        """Warm up eagerly (lazy initialisation, NCCL communicators, allocator), restore the training state the
        warm-up step changed, then record the step into a CUDA graph on static input buffers."""
        new = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.device)
        static = dict(x=tuple(new(t) for t in x), t1=new(t1), t2=new(t2), augs=[tuple(new(t) for t in a) for a in augs])
        self._fill_static(static, x, t1, t2, augs)
        state = self._state_tensors()
        saved = [t.clone() for t in state]
        host = (self.opt1.t, self.opt2.t, self.steps)
        torch.cuda.synchronize(self.device)
        self._step_eager(static["x"], static["t1"], static["t2"], static["augs"], rate)
        torch.cuda.synchronize(self.device)
        with torch.no_grad():
            for t, s in zip(state, saved):
                t.copy_(s)
        self.opt1.t, self.opt2.t, self.steps = host
        for net in (self.net1, self.net2):
            net._weights = None                       # force the weight preparation INTO the graph
        graph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize(self.device)
        n0 = lib.aide_launch_count()
        with torch.cuda.graph(graph):
            out = self._step_eager(static["x"], static["t1"], static["t2"], static["augs"], rate)
        n_kernels = lib.aide_launch_count() - n0               # kernel nodes of this library recorded in the graph
        self.opt1.t, self.opt2.t, self.steps = host          # capture records, it does not execute
        self._graphs[key] = (graph, static, out, n_kernels)
        return self._graphs[key]

    @staticmethod
    def _fill_static(static, x, t1, t2, augs):
        for d, s in zip(static["x"], x):
            d.copy_(s, non_blocking=True)
        static["t1"].copy_(t1, non_blocking=True)
        static["t2"].copy_(t2, non_blocking=True)
        for da, sa in zip(static["augs"], augs):
            for d, s in zip(da, sa):
                d.copy_(s, non_blocking=True)

    def broadcast_parameters(self, src: int = 0) -> None:
        """Make every rank start from rank `src`'s weights (the reference's DataParallel replicates rank 0)."""
        if self.world > 1:
            for opt in (self.opt1, self.opt2):
                torch.distributed.broadcast(opt.flat, src, group=self.group)
            for net in (self.net1, self.net2):
                for b in net.buffers():
                    torch.distributed.broadcast(b, src, group=self.group)

    def _inputs(self, x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x,)

    # ---- reverse augmentation parameters (host builds PIL's matrices, the kernel does the resampling) -----------
    @staticmethod
    def _aug_host(augset, n_views: int, B: int, H: int, W: int):
        mats = torch.empty((n_views, B, 6), dtype=torch.float64)
        modes = torch.empty((n_views, B), dtype=torch.int32)
        flips = torch.empty((n_views, B), dtype=torch.int32)
        for v in range(n_views):
            for b in range(B):
                mode, m = AUG.rotate_matrix(0 - float(augset[f"degree{v + 1}"][b]), W, H)
                if int(augset["augno"][b]) <= v:          # this sample has fewer views: leave the plane untouched
                    mode, m = 1, [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]
                mats[v, b] = torch.tensor(m, dtype=torch.float64)
                modes[v, b] = mode
                flips[v, b] = int(augset[f"hflip{v + 1}"][b]) if int(augset["augno"][b]) > v else 0
        return mats, modes, flips

    def _reverse(self, t: torch.Tensor, v: int) -> torch.Tensor:
        mats, modes, flips = self._aug_dev
        B, K, H, W = t.shape
        out = torch.empty_like(t)
        call("aide_reverse_aug", t.data_ptr(), out.data_ptr(), mats[v].data_ptr(), modes[v].data_ptr(),
             flips[v].data_ptr(), B, K, H, W, torch.cuda.current_stream().cuda_stream)
        return out

    def _half(self, net, opt, xs, augs, other: Dict, me: Dict, stage: int, t_other, rate: float):
        """One network's share of the step, split in three stages around the two joins."""
        if stage == 0:
            # augmented forwards (no tape) -> pseudo label of THIS net; then the train forward (tape kept)
            was_training = net.training
            if self.flavour != "chaos":
                net.eval()
            stacked_train = (self.stack_train and self.group_augs and len(augs) > 1 and net.training)
            with torch.no_grad():
                if stacked_train:
                    # the train forward rides as the LAST statistics group of the stacked pseudo-label forward (same
                    # BatchNorm mode, same order of running-statistics updates as :265-269 followed by :301-302): one
                    # launch per layer for 5 B images; the backward later runs through that group's slice of the tape
                    views = [self._inputs(v) for v in augs] + [xs]
                    for v in views:
                        net._check_inputs(v)
                    logits_all, me["tape"] = net._engine_forward(views, keep_tape=True, groups=len(views))
                    chunks = list(logits_all.chunk(len(views), dim=0))
                    a, me["logits"] = chunks[:-1], chunks[-1]
                elif self.group_augs and len(augs) > 1:
                    a = net._engine_forward_grouped([self._inputs(v) for v in augs])
                else:
                    a = [net._engine_forward(self._inputs(v), keep_tape=False)[0] for v in augs]
                if self._aug_dev is not None:      # undo the views' flip / rotation (reference :271-272 -> :81-95)
                    a = [self._reverse(t, v) for v, t in enumerate(a)]
            net.train(was_training)
            if a:
                me["q"], me["w"] = L.pseudo_label(a, self.temperature, self.flavour)
            else:
                me["q"] = me["w"] = None
            if not stacked_train:
                me["logits"], me["tape"] = net._engine_forward(xs, keep_tape=True)
        elif stage == 1:
            # own per-image CE+Dice against the OTHER net's targets, consistency against the OTHER net's pseudo label
            lg = me["logits"]
            H, W = lg.shape[2:]
            me["sums"] = L.image_sums(lg, t_other, other["q"], other["w"])
            me["pre"], _, me["dice"] = L.image_finalize(me["sums"], H, W, 1.0, 1.0, 1.0, want_dice_fn=True)
        else:
            lg = me["logits"]
            N, _, H, W = lg.shape
            dev = lg.device
            idx = torch.empty(N, dtype=torch.int64, device=dev)
            coef = torch.empty((3, N), dtype=torch.float32, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            pre_all, n_total, first = other["pre"], N, 0
            if self.global_select:                 # DataParallel semantics: rank the gathered global batch
                n_total, first = N * self.world, N * self.rank
                pre_all = torch.empty(n_total, dtype=torch.float32, device=dev)
                torch.distributed.all_gather_into_tensor(pre_all, other["pre"].contiguous(), group=self.group)
                idx = torch.empty(n_total, dtype=torch.int64, device=dev)
            call("aide_coteach_select_ex", pre_all.data_ptr(), n_total, first, me["pre"].data_ptr(),
                 me["sums"].data_ptr(), N, H, W, min(self.n_clean, n_total), float(rate), self.rate_dev.data_ptr(),
                 float(self.segcor_weight[0]), float(self.segcor_weight[1]), 1.0, 1.0, idx.data_ptr(),
                 coef[0].data_ptr(), coef[1].data_ptr(), coef[2].data_ptr(), loss.data_ptr(),
                 torch.cuda.current_stream().cuda_stream)
            has_q = other["q"] is not None
            d = L.loss_backward(lg, t_other, me["sums"], coef[0], coef[1], coef[2] if has_q else None,
                                other["q"], other["w"])
            works = []
            bucketed = self.world > 1 and self.n_buckets > 1 and self._comm
            peer = self._peer
            if peer is not None:
                ci = 0 if net is self.net1 else 1
                cs = self._comm_streams[ci]

                def reduce_range(lo, hi):          # on the net's communication stream, behind everything enqueued so far
                    cs.wait_stream(torch.cuda.current_stream())
                    if self._sync_only:            # diagnostic: the cross-GPU barriers without the data
                        hi = lo + 4
                    peer.all_reduce(ci, lo, hi, cs.cuda_stream)
            else:
                def reduce_range(lo, hi):
                    works.append(torch.distributed.all_reduce(net.last_grad_flat[lo:hi], group=self.group, async_op=True))
            if bucketed:
                ranges = self._buckets[id(net)]

                def on_done(name):                 # these float ranges of the flat gradient are final: reduce them now
                    for lo, hi in ranges.get(name, ()):
                        reduce_range(lo, hi)
            net._engine_backward(me["tape"], d, on_done if bucketed else None,
                                 set(self._buckets[id(net)]) if bucketed else None)
            flat = net.last_grad_flat
            if self.world > 1 and self._comm:
                if not bucketed:
                    if peer is not None:
                        reduce_range(0, flat.numel())
                    else:
                        torch.distributed.all_reduce(flat, group=self.group)
                for w_ in works:
                    w_.wait()
                if peer is not None:
                    torch.cuda.current_stream().wait_stream(cs)
                if self.global_select:             # the scalar was this rank's share of the global loss
                    torch.distributed.all_reduce(loss, group=self.group)
            # local selection: mean of the per-rank gradients; global selection: the coefficients already carry the
            # global denominators, so the rank gradients add up to the reference's gradient
            opt.step(flat, grad_scale=1.0 if self.global_select else 1.0 / self.world)
            me["loss"], me["idx"] = loss, idx
            me["tape"] = None

    def step(self, x, t1: torch.Tensor, t2: torch.Tensor, augs: Sequence, rate: float,
             augset: Optional[Dict] = None) -> Dict[str, torch.Tensor]:
        """x / augs[i]: a tensor [B,3,H,W] (unet) or a pair of them (fuseunet), on the device OR in (pinned) host
        memory.  t1, t2: [B,H,W] int64.  Returns device scalars loss1, loss2, dice1, dice2 (Dice_fn batch sums), the
        index vectors and both logits; no host sync.  With cuda_graph=True the returned tensors are the graph's
        static outputs: they are overwritten by the next step().  augset: the reference's dict of the views' forward
        augmentation (``augno``, ``degree{k}``, ``hflip{k}`` per sample); the views' logits are flipped / rotated back
        on the GPU before the pseudo label is built (None = identity, as in the synthetic benchmark)."""
        xs = self._inputs(x)
        augs = [self._inputs(a) for a in augs]
        host_aug = None
        if augset is not None and len(augs) > 0:
            B, _, H, W = xs[0].shape
            host_aug = self._aug_host(augset, len(augs), B, H, W)
        if not self.cuda_graph:
            dev = self.device
            mv = lambda t: t if t.is_cuda else t.to(dev, non_blocking=True)
            self._aug_dev = tuple(t.to(dev) for t in host_aug) if host_aug is not None else None
            try:
                return self._step_eager(tuple(mv(t) for t in xs), mv(t1), mv(t2), [tuple(mv(t) for t in a) for a in augs],
                                        rate)
            finally:
                self._aug_dev = None
        self._upload_scalars(rate)
        key = (tuple(xs[0].shape), len(xs), len(augs), self.flavour, self.world, host_aug is not None,
               self.global_select)
        entry = self._graphs.get(key)
        if entry is not None:
            self._graphs[key] = self._graphs.pop(key)          # most recently used last
        if entry is None:
            while len(self._graphs) >= self.max_graphs:        # every graph owns a multi-GB private pool: evict the LRU
                self._graphs.pop(next(iter(self._graphs)))
                torch.cuda.empty_cache()
            if host_aug is not None:                           # static device buffers the captured kernels read
                self._aug_dev = tuple(t.to(self.device) for t in host_aug)
            try:
                entry = self._capture(key, xs, t1, t2, augs, rate)
                entry[1]["aug"] = self._aug_dev
            finally:
                self._aug_dev = None
        graph, static, out, n_kernels = entry
        self._fill_static(static, xs, t1, t2, augs)            # D2D, or H2D straight into the graph's inputs
        self._consumed = torch.cuda.Event()                     # the step's inputs have left their source buffers
        self._consumed.record()
        if host_aug is not None:
            for d, s_ in zip(static["aug"], host_aug):
                d.copy_(s_, non_blocking=True)
        graph.replay()
        self.graph_launches += n_kernels
        self.steps += 1
        self.opt1.t += 1
        self.opt2.t += 1
        for net in (self.net1, self.net2):                      # operand-format weight planes live in the graph's pool
            net._weights = None
        # the replayed Adam kernels rewrote the parameters: bump their versions so that everything keyed on
        # (data_ptr, _version) -- PreparedWeights, net.graphed_eval() -- re-derives its weight planes
        self.opt1.bump_versions()
        self.opt2.bump_versions()
        return out

    def _upload_scalars(self, rate: float) -> None:
        """Mirror the host scalars a captured step reads from device memory (rate, both learning rates)."""
        if rate != self._rate_uploaded:
            self.rate_dev.fill_(float(rate))
            self._rate_uploaded = rate
        self.opt1.sync_lr()
        self.opt2.sync_lr()

    def set_lr(self, lr: float) -> None:
        self.opt1.set_lr(lr)
        self.opt2.set_lr(lr)

    def _step_eager(self, x, t1: torch.Tensor, t2: torch.Tensor, augs: Sequence, rate: float) -> Dict[str, torch.Tensor]:
        xs = self._inputs(x)
        if not torch.cuda.is_current_stream_capturing():
            self._upload_scalars(rate)
        m1: Dict = {}
        m2: Dict = {}
        cur = torch.cuda.current_stream(self.device)
        if not self.two_streams:
            for stage in range(3):
                self._half(self.net1, self.opt1, xs, augs, m2, m1, stage, t2, rate)
                self._half(self.net2, self.opt2, xs, augs, m1, m2, stage, t1, rate)
        else:
            s1, s2 = self.s1, self.s2
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            for stage in range(3):
                with torch.cuda.stream(s1):
                    self._half(self.net1, self.opt1, xs, augs, m2, m1, stage, t2, rate)
                with torch.cuda.stream(s2):
                    self._half(self.net2, self.opt2, xs, augs, m1, m2, stage, t1, rate)
                if stage < 2:                      # join: each net needs the other's q/w (stage 1) and pre (stage 2)
                    e1, e2 = s1.record_event(), s2.record_event()
                    s1.wait_event(e2)
                    s2.wait_event(e1)
            cur.wait_stream(s1)
            cur.wait_stream(s2)
        self.steps += 1
        return dict(loss1=m1["loss"], loss2=m2["loss"], dice1=m1["dice"], dice2=m2["dice"],
                    indx1=m2["idx"], indx2=m1["idx"], pre1=m1["pre"], pre2=m2["pre"],
                    out1=m1["logits"], out2=m2["logits"])

    # ------------------------------------------------------------------------------------------
    def _stage(self, host_batch: Dict) -> None:
        """Start the H2D copy of the NEXT step's inputs on the copy stream, into staging buffers on the device: it runs
        while the current step computes (the loader-side prefetch of a training loop)."""
        flat = list(self._inputs(host_batch["x"])) + [host_batch["t1"], host_batch["t2"]] + \
            [t for a in host_batch["augs"] for t in self._inputs(a)]
        key = tuple((tuple(t.shape), t.dtype) for t in flat)
        st = self._staging.get(key)
        if st is None:
            st = self._staging[key] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in flat]
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)          # the previous contents of the staging buffers were copied onwards
        else:
            cs.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(cs):
            for d, s_ in zip(st, flat):
                d.copy_(s_, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(cs)
        nx, na = len(self._inputs(host_batch["x"])), len(host_batch["augs"])
        per = len(self._inputs(host_batch["augs"][0])) if na else 0
        dev = dict(x=tuple(st[:nx]), t1=st[nx], t2=st[nx + 1],
                   augs=[tuple(st[nx + 2 + v * per: nx + 2 + (v + 1) * per]) for v in range(na)])
        self._staged = (host_batch, dev, ready)

    def step_from_host(self, host_batch: Dict, rate: float, device_batch: Optional[Dict] = None,
                       prefetch: Optional[Dict] = None):
        """End-to-end form: `host_batch` holds pinned CPU tensors (x: tuple, t1, t2, augs: list of tuples).
        Copies them to the device (async), runs the step and reads the two losses and Dice sums back to the host.
        `prefetch`: the batch of the NEXT call -- its H2D copy is started on a copy stream right after this step's launch
        and overlaps the step's compute; the next call (given the same dict) then finds its inputs on the device.
        Returns (dict of python floats, h2d bytes, d2h bytes)."""
        h2d = 0
        for t in list(self._inputs(host_batch["x"])) + [host_batch["t1"], host_batch["t2"]] + \
                [t for a in host_batch["augs"] for t in self._inputs(a)]:
            h2d += t.numel() * t.element_size()
        src = host_batch
        if self._staged is not None and self._staged[0] is host_batch:          # prefetched by the previous call
            _, src, ready = self._staged
            torch.cuda.current_stream(self.device).wait_event(ready)
        self._staged = None
        out = self.step(src["x"], src["t1"], src["t2"], src["augs"], rate)
        if not self.cuda_graph:                    # eager steps read their inputs until they finish
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        if prefetch is not None:
            self._stage(prefetch)
        res = torch.stack([out["loss1"], out["loss2"], out["dice1"], out["dice2"]]).to("cpu")   # synchronising D2H
        vals = res.tolist()
        return dict(loss1=vals[0], loss2=vals[1], dice1=vals[2], dice2=vals[3]), h2d, res.numel() * 4
