"""Optimiser pieces ("next" row f1 of SURVEY.md section 8).

FlatAdamAMSGrad = torch.optim.Adam(lr, amsgrad=True) (reference:
train_files/trainchaos_proposed_30cases1labeled.py:231-232,323,325) applied by ONE kernel to flat fp32
parameter / gradient / state buffers (aide_adam_amsgrad).  It is a torch.optim.Optimizer, so the reference's
schedulers (StepLR :236-237, PolyLR :239-240 -> utils/poly_lr_scheduler.py:31-51) wrap it unchanged: they write
``param_groups[0]['lr']``, which is mirrored into a one-element DEVICE buffer the kernel reads -- a learning-rate
change therefore also reaches a captured CUDA graph of the training step.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch
from torch.optim.lr_scheduler import _LRScheduler

from ._lib import call


class FlatAdamAMSGrad(torch.optim.Optimizer):
    """Keeps a flat copy-free view of the parameters: on construction the parameters are re-pointed into
    one contiguous fp32 buffer (so ``module.parameters()`` keep working) and the three Adam states are
    flat buffers of the same size.  ``step(grad_flat)`` consumes a flat gradient laid out in the same
    parameter order (see flatten_grads) -- or the per-parameter ``.grad`` fields when called without."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 offsets: Optional[Sequence[int]] = None, total: Optional[int] = None):
        """offsets / total (in floats): an explicit layout of the flat buffer (engine.GradLayout packs some small
        tensors without padding); default: consecutive tensors, each padded to a multiple of 4 floats."""
        plist: List[torch.nn.Parameter] = [p for p in params]
        super().__init__(plist, dict(lr=lr, betas=betas, eps=eps))
        if len(self.param_groups) != 1:
            raise ValueError("FlatAdamAMSGrad keeps ONE flat buffer: pass a flat list of parameters, not groups")
        self.params = plist
        self.betas, self.eps, self.t = betas, eps, 0
        self.offsets, cur = [], 0
        for p in self.params:
            self.offsets.append(cur)
            cur += (p.numel() + 3) // 4 * 4
        if offsets is not None:
            self.offsets, cur = [int(o) for o in offsets], int(total)
            if len(self.offsets) != len(self.params):
                raise ValueError("one offset per parameter")
        dev = self.params[0].device
        self.flat = torch.zeros(cur, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, self.offsets):
            self.flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + p.numel()].view_as(p.data)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.vmax = torch.zeros_like(self.flat)
        self.gbuf = torch.zeros_like(self.flat)
        # device-resident step counter, bias-correction scratch and learning rate: nothing host-side changes between
        # steps, so a captured CUDA graph of the training step replays correctly
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.bc_dev = torch.zeros(2, dtype=torch.float32, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_uploaded = float(lr)

    # the learning rate lives in param_groups[0]['lr'] (what torch's schedulers read and write)
    @property
    def lr(self) -> float:
        return float(self.param_groups[0]["lr"])

    @lr.setter
    def lr(self, value: float) -> None:
        self.param_groups[0]["lr"] = float(value)

    def set_lr(self, value: float) -> None:
        self.lr = value
        self.sync_lr()

    def sync_lr(self) -> None:
        """Mirror param_groups[0]['lr'] into the device scalar (no-op while it is unchanged)."""
        lr = self.lr
        if lr != self._lr_uploaded:
            self.lr_dev.fill_(lr)
            self._lr_uploaded = lr

    def flatten_grads(self) -> torch.Tensor:
        for p, o in zip(self.params, self.offsets):
            if p.grad is not None:
                self.gbuf[o:o + p.numel()].copy_(p.grad.reshape(-1))
            else:
                self.gbuf[o:o + p.numel()].zero_()
        return self.gbuf

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_flat: torch.Tensor = None, grad_scale: float = 1.0):
        g = self.flatten_grads() if grad_flat is None else grad_flat
        if g.numel() != self.flat.numel():
            raise ValueError("flat gradient does not match the flat parameter buffer")
        self.t += 1                      # host mirror (informational; the kernel uses the device counter)
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lr()               # under capture the caller syncs before the replay (AideTrainer.step)
        call("aide_adam_amsgrad_dev", self.flat.data_ptr(), g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
             self.vmax.data_ptr(), self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps,
             self.step_dev.data_ptr(), self.bc_dev.data_ptr(), grad_scale, self.lr_dev.data_ptr(),
             torch.cuda.current_stream().cuda_stream)
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that the
        # engine re-derives its operand-format weight planes (PreparedWeights is keyed on _version)
        self.bump_versions()

    def bump_versions(self) -> None:
        torch.autograd.graph.increment_version(self.params)

    # ---- checkpoint / resume: the entries of torch.optim.Adam(amsgrad=True).state_dict() (the reference's optimiser,
    # trainchaos_proposed_30cases1labeled.py:231-232), parameters numbered in this optimiser's (flat-buffer) order
    def state_dict(self):
        step = float(int(self.step_dev.item()))
        state = {}
        if step > 0:
            for i, (p, o) in enumerate(zip(self.params, self.offsets)):
                sl = slice(o, o + p.numel())
                state[i] = dict(step=torch.tensor(step), exp_avg=self.m[sl].view_as(p).clone(),
                                exp_avg_sq=self.v[sl].view_as(p).clone(), max_exp_avg_sq=self.vmax[sl].view_as(p).clone())
        group = dict(lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=0, amsgrad=True,
                     params=list(range(len(self.params))))
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, sd) -> None:
        group = sd["param_groups"][0]
        if len(group["params"]) != len(self.params):
            raise ValueError("optimiser state has a different number of parameters")
        if not group.get("amsgrad", True) or group.get("weight_decay", 0) != 0:
            raise ValueError("FlatAdamAMSGrad is Adam(amsgrad=True, weight_decay=0)")
        self.betas, self.eps = tuple(group["betas"]), float(group["eps"])
        self.param_groups[0]["betas"], self.param_groups[0]["eps"] = self.betas, self.eps
        self.set_lr(float(group["lr"]))
        for buf in (self.m, self.v, self.vmax):
            buf.zero_()
        step = 0
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is None:
                continue
            sl = slice(o, o + p.numel())
            self.m[sl].copy_(st["exp_avg"].reshape(-1))
            self.v[sl].copy_(st["exp_avg_sq"].reshape(-1))
            self.vmax[sl].copy_(st["max_exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
        self.step_dev.fill_(step)
        self.t = step


class PolyLR(_LRScheduler):
    def __init__(self, optimizer, max_epoch, power=0.9, last_epoch=-1):
        self.max_epoch, self.power = max_epoch, power
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        return [b * ((1.0 - float(self.last_epoch % self.max_epoch) / float(self.max_epoch)) ** self.power)
                for b in self.base_lrs]
