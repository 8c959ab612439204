# -*- coding: UTF-8 -*-
"""
Peer-memory arenas for the policy-sharded train step (one process per GPU on one NVSwitch node).

Each rank allocates an arena through the C ABI (sfgpi_peer_alloc: cudaMalloc + CUDA IPC handle), the 64-byte handles travel
once over the existing torch.distributed group, and every rank maps every peer's arena (sfgpi_peer_open).  After that the
per-step exchanges (GPI keys MAX reduce-scatter, [w | delta h] all-gather) are kernels of libsfgpi.so that pull
from the peers over NVLink -- see csrc/peer.cu.  The collectives in dist.py stay as the transport for everything else and as
the path taken when peer mapping is unavailable (SFGPI_PEER=0, uneven shards, > 16 ranks, IPC refused by the platform).
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib


def peer_mode_wanted():
    return os.environ.get('SFGPI_PEER', '1') != '0'


class PeerArena:
    """`nbytes` of zeroed device memory on every rank of `group`, mapped into every rank: ptrs[r] = rank r's arena here."""

    def __init__(self, nbytes, group=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f'peer exchange supports at most {_lib.MAX_PEERS} ranks')
        self.nbytes = int(nbytes)
        self.local = None
        self.ptrs = [None] * self.world
        self._opened = []
        handle = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
        p = C.c_void_p()
        err = None
        try:
            _lib.call('sfgpi_peer_alloc', self.nbytes, C.byref(p), handle)
            self.local = p.value
        except RuntimeError as e:
            err = str(e)
        mine = (err, bytes(handle))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        bad = [f'rank {r}: {e}' for r, (e, _) in enumerate(everyone) if e]
        if not bad:
            for r, (_, hb) in enumerate(everyone):
                if r == self.rank:
                    self.ptrs[r] = self.local
                    continue
                q = C.c_void_p()
                try:
                    _lib.call('sfgpi_peer_open', (C.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(hb), C.byref(q))
                    self.ptrs[r] = q.value
                    self._opened.append(q.value)
                except RuntimeError as e:
                    err = str(e)
                    break
            flags = [None] * self.world
            dist.all_gather_object(flags, err, group=group)
            bad = [f'rank {r}: {e}' for r, e in enumerate(flags) if e]
        if bad:
            self.close()
            raise RuntimeError('peer arena could not be mapped on every rank: ' + '; '.join(bad))

    def close(self):
        for q in self._opened:
            try:
                _lib.call('sfgpi_peer_close', C.c_void_p(q))
            except RuntimeError:
                pass
        self._opened = []
        if self.local is not None:
            try:
                _lib.call('sfgpi_peer_free', C.c_void_p(self.local))
            except RuntimeError:
                pass
            self.local = None

    def ctx(self):
        """sfgpi_peer_ctx with the flag blocks at offset 0 of every arena."""
        c = _lib.PeerCtx()
        c.world, c.rank = self.world, self.rank
        for r in range(self.world):
            c.flags[r] = self.ptrs[r]
        return c

    def view(self, offset, shape, dtype):
        """torch view of the LOCAL arena (tests / debugging): zero-copy through __cuda_array_interface__."""
        n = 1
        for s in shape:
            n *= s
        typestr = {torch.int64: '<i8', torch.float32: '<f4'}[dtype]

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(self.local + offset, False), version=2)
        return torch.as_tensor(raw, device=torch.device('cuda', torch.cuda.current_device()))


FLAG_BYTES = 1024          # SFGPI_PEER_CHANNELS * SFGPI_MAX_PEERS * 8 = 512, padded
