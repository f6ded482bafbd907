"""
Peer-shareable device buffers for the one-process-per-GPU layouts: a buffer allocated here can be read by
the CUDA kernels of every other rank directly over NVLink (CUDA IPC mapping, tt_peer_alloc / tt_peer_open in include/tt.h).
Used for row-sharded embedding tables (the forward gather reads rows where they live) and for the towers' dX blocks (the
owner of a row pulls its gradient rows from the rank that produced them).  Set-up only; nothing here runs per step.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

from pkg import _native as N

_TYPESTR = {"float32": "<f4", "int32": "<i4"}


class _Raw:
    """Minimal __cuda_array_interface__ carrier so torch can view memory this library allocated."""

    def __init__(self, ptr: int, shape: Sequence[int], typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerBuffer:
    """``local``: torch view of this rank's allocation; ``ptrs[r]``: address of rank r's allocation as seen from THIS
    process (``ptrs[rank]`` is the local one).  Creation is collective over ``group``."""

    def __init__(self, shape: Sequence[int], dtype: str = "float32", group=None):
        import torch.distributed as dist

        torch = N.require_cuda()
        lib = N.load()
        numel = 1
        for d in shape:
            numel *= int(d)
        nbytes = max(256, (numel * 4 + 255) // 256 * 256)
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        N.check(lib.tt_peer_alloc(nbytes, ctypes.byref(ptr), handle), "tt_peer_alloc")
        self._ptr = int(ptr.value)
        self.local = torch.as_tensor(_Raw(self._ptr, shape, _TYPESTR[dtype]), device=torch.device("cuda", torch.cuda.current_device()))
        assert self.local.data_ptr() == self._ptr
        world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        handles: List[bytes] = [b""] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs: List[int] = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self._ptr)
                continue
            p = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            N.check(lib.tt_peer_open(buf, ctypes.byref(p)), f"tt_peer_open(rank {r})")
            self.ptrs.append(int(p.value))
        # device-side pointer table (what tt_feature.table points at for a row-sharded table)
        self.ptr_table = torch.tensor(self.ptrs, dtype=torch.int64, device="cuda")
        self._group = group
        self._closed = False

    def close(self) -> None:
        """Collective over the group: every rank unmaps the peers' allocations, waits until all ranks have done so, then frees its own
        (an exporter must not free memory a peer still maps).  The views handed out (``local``, ``ptr_table``) are dead afterwards.
        Called when a data-parallel step workspace is evicted; process exit releases whatever is still open."""
        import torch.distributed as dist

        if self._closed:
            return
        torch = N.require_cuda()
        lib = N.load()
        torch.cuda.synchronize()
        for r, p in enumerate(self.ptrs):
            if r != self.rank:
                N.check(lib.tt_peer_close(ctypes.c_void_p(p)), f"tt_peer_close(rank {r})")
        dist.barrier(group=self._group)
        N.check(lib.tt_peer_free(ctypes.c_void_p(self._ptr)), "tt_peer_free")
        self._closed = True
        self.local = None
        self.ptr_table = None
