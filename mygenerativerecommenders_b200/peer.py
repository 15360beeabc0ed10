"""Peer-memory plumbing for the multi-GPU paths: symmetric buffers (torch symmetric memory gives
every rank the device pointers of all ranks' copies) and a stream barrier built on them.

The exchanges themselves are this package's kernels (``grb_p2p_put_rows``, ``grb_p2p_put_table_rows``):
plain stores / ``red.global`` into the peers' buffers over NVLink / NVSwitch, ordered by
``grb_p2p_barrier`` — no NCCL call and no host synchronisation on the data path."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def pointer_array(ptrs) -> "C.Array":
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def symmetric_empty(shape, dtype, device, group):
    """(tensor, handle, ctypes array of every rank's device pointer)."""
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(tuple(shape), dtype=dtype, device=device)
    h = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
    return t, h, pointer_array(h.buffer_ptrs)


class PeerBarrier:
    """``barrier(slot)``: every rank's stream waits until all ranks' streams got there; stores and
    reds a rank issued to peer memory before the barrier are visible to every rank after it."""

    SLOTS = 8

    def __init__(self, group, device) -> None:
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.signals, self._h, self._ptrs = symmetric_empty((self.SLOTS,), torch.int64, device, group)
        self.signals.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group=group)          # nobody signals before every array is zero
        self._epoch = [0] * self.SLOTS

    def barrier(self, slot: int, device) -> None:
        self._epoch[slot] += 1
        _lib.check(_lib.lib().grb_p2p_barrier(self._ptrs, self.world, self.rank, slot, self._epoch[slot],
                                              _lib.stream_ptr(device)))
