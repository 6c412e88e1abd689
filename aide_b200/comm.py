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
        # Every rank walks through the same collectives whether or not its own CUDA calls succeed; the verdict is agreed
        # on at the end, so a rank whose IPC mapping fails cannot leave the others waiting in a different collective.
        mine, handles, err = [], [], None
        self._opened, self.ptrs = [], []
        with torch.cuda.device(device):
            try:
                for n in self.sizes + [pad_words]:
                    ptr, h = C.c_void_p(), C.create_string_buffer(64)
                    call("aide_comm_alloc", n * 4, C.byref(ptr), h)
                    mine.append(ptr.value)
                    handles.append(h.raw)
                torch.cuda.synchronize(device)
            except Exception as e:  # noqa: BLE001
                err, handles = e, None
            everyone: List = [None] * self.world
            dist.all_gather_object(everyone, handles, group=group)
            if err is None and any(h is None for h in everyone):
                err = RuntimeError("a peer could not allocate its buffers")
            if err is None:
                try:
                    # ptrs[k][r]: allocation k (buffers, then the pad) of rank r, mapped into this process
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
                except Exception as e:  # noqa: BLE001 -- e.g. CUDA IPC not permitted between these processes
                    err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)       # also: everybody has mapped everything
        self._mine = mine
        if int(ok.item()) == 0:
            self._release()
            raise RuntimeError(f"peer-memory buffers unavailable on rank {self.rank} or one of its peers: {err!r}")
        self._raw = [_RawCuda(mine[i], self.sizes[i]) for i in range(self.channels)]
        self.tensors = [torch.as_tensor(r, device=device) for r in self._raw]
        for t, p in zip(self.tensors, mine):
            if t.data_ptr() != p:
                raise RuntimeError("torch copied the peer buffer instead of viewing it")
        self._pad_arr = (C.c_void_p * self.world)(*self.ptrs[-1])
        self._buf_arr = [(C.c_void_p * self.world)(*self.ptrs[i]) for i in range(self.channels)]

    def all_reduce(self, i: int, lo: int, hi: int, stream: int) -> None:
        """Sum floats [lo, hi) of buffer i over the ranks, in place, on CUDA stream `stream` (raw handle)."""
        lo4 = lo - lo % 4
        hi4 = min(hi + (-hi) % 4, self.sizes[i])
        call("aide_allreduce_p2p", self._buf_arr[i], self._pad_arr, self.rank, self.world, i, self.channels, lo4, hi4 - lo4,
             self.blocks, stream)

    def _release(self) -> None:
        for p in self._opened:
            call("aide_comm_close", p)
        self._opened = []
        self.tensors = []
        for p in self._mine:
            call("aide_comm_free", p)
        self._mine = []

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._release()
