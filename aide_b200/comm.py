"""Peer-memory gradient all-reduce of the data-parallel step (csrc/comm.cu): host-side plumbing.

Replaces the gradient reduction that ``nn.DataParallel`` performs in the reference
(train_files/trainchaos_proposed_30cases1labeled.py:188-189) for the one-process-per-GPU layout.
``torch.distributed`` is only used to exchange the 64-byte CUDA IPC handles once; the collective itself is the
``aide_allreduce_p2p`` kernel reading the peers' buffers over NVLink.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
import torch.distributed as dist

from ._lib import call, lib


class _RawCuda:
    """A cudaMalloc'ed float32 array owned by libaide_b200, visible to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerBuffers:
    """`sizes[i]` float32 buffers that exist on every rank of `group`, each rank holding pointers to all of them, plus
    the flag pad of the all-reduce kernel.  One channel per buffer: calls on a buffer must be stream-ordered (they are:
    one communication stream per network), calls on different buffers may overlap."""

    def __init__(self, group, device: torch.device, sizes: Sequence[int], blocks: int = 32):
        self.group, self.device = group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if not 2 <= self.world <= 8:
            raise ValueError("peer-memory all-reduce: 2..8 ranks on one NVLink domain")
        self.blocks = blocks
        self.channels = len(sizes)
        self.sizes = [int(n + (-n) % 4) for n in sizes]
        pad_words = lib.aide_comm_pad_words(self.world, self.channels)
        mine, handles = [], []
        with torch.cuda.device(device):
            for n in self.sizes + [pad_words]:
                ptr, h = C.c_void_p(), C.create_string_buffer(64)
                call("aide_comm_alloc", n * 4, C.byref(ptr), h)
                mine.append(ptr.value)
                handles.append(h.raw)
            torch.cuda.synchronize(device)
            everyone: List = [None] * self.world
            dist.all_gather_object(everyone, handles, group=group)
            # ptrs[k][r]: allocation k (buffers, then the pad) of rank r, mapped into this process
            self._opened = []
            self.ptrs = []
            for k in range(len(mine)):
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(mine[k])
                    else:
                        p = C.c_void_p()
                        call("aide_comm_open", everyone[r][k], C.byref(p))
                        self._opened.append(p.value)
                        row.append(p.value)
                self.ptrs.append(row)
        self._mine = mine
        self._raw = [_RawCuda(mine[i], self.sizes[i]) for i in range(self.channels)]
        self.tensors = [torch.as_tensor(r, device=device) for r in self._raw]
        for t, p in zip(self.tensors, mine):
            if t.data_ptr() != p:
                raise RuntimeError("torch copied the peer buffer instead of viewing it")
        self._pad_arr = (C.c_void_p * self.world)(*self.ptrs[-1])
        self._buf_arr = [(C.c_void_p * self.world)(*self.ptrs[i]) for i in range(self.channels)]
        dist.barrier(group=group)                    # every rank has mapped everything before the first kernel runs

    def all_reduce(self, i: int, lo: int, hi: int, stream: int) -> None:
        """Sum floats [lo, hi) of buffer i over the ranks, in place, on CUDA stream `stream` (raw handle)."""
        lo4 = lo - lo % 4
        hi4 = min(hi + (-hi) % 4, self.sizes[i])
        call("aide_allreduce_p2p", self._buf_arr[i], self._pad_arr, self.rank, self.world, i, self.channels, lo4, hi4 - lo4,
             self.blocks, stream)

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        for p in self._opened:
            call("aide_comm_close", p)
        self._opened = []
        self.tensors = []
        for p in self._mine:
            call("aide_comm_free", p)
        self._mine = []
