"""Peer-mapped device arenas over NVLink: every rank allocates one arena and maps all the others' into its own GPU context
(CUDA IPC handles exchanged through torch.distributed, opened by libaria_b200 with the local device current).  Kernels
launched on the local GPU can then load/store the peers' arenas directly through NVSwitch — the basis of the fused
dispatch / combine kernels of the expert-parallel path."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


class PeerArena:
    """`nbytes` of device memory on every rank; `ptr(rank, offset)` is the address of byte `offset` of rank's arena as seen
    from THIS GPU.  `local` is the torch view (uint8) of this rank's own arena."""

    def __init__(self, nbytes: int, device, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device)
        self.local = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        lib = L.load()
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        with torch.cuda.device(self.device):
            L.check(lib.aria_ipc_export(C.c_void_p(self.local.data_ptr()), handle, C.byref(off)), "ipc_export")
        metas = [None] * self.world
        dist.all_gather_object(metas, (bytes(handle), int(off.value)), group=group)
        self.bases = []
        self._opened = []
        for r, (h, o) in enumerate(metas):
            if r == self.rank:
                self.bases.append(self.local.data_ptr())
                continue
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            base = C.c_void_p()
            with torch.cuda.device(self.device):
                L.check(lib.aria_ipc_open(buf, C.byref(base)), "ipc_open")
            self._opened.append(base.value)
            self.bases.append(base.value + o)
        dist.barrier(group)

    def ptr(self, rank: int, offset: int = 0) -> int:
        return self.bases[rank] + offset

    def local_view(self, offset: int, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        return self.local[offset:offset + nbytes].view(dtype).view(*shape)
