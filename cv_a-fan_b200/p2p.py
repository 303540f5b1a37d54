"""Peer-mapped mailboxes for the fused dual-BN statistics exchange (one process per GPU, one node).

Host plumbing only: allocate this rank's mailbox with the C ABI (cudaMalloc), all-gather the 64-byte cudaIpc
handles over torch.distributed, open every peer's mailbox (NVLink/NVSwitch peer access is enabled lazily by
the open), and hand the kernels a host array of `world` device pointers plus a local {seq, ticket, error} state.
The exchange itself happens INSIDE the BatchNorm kernels (csrc/afan_bn.cu: cluster_fold_p2p).
"""
import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import AfanError, check


class PeerMailbox:
    def __init__(self, process_group=None, device: Optional[torch.device] = None, cmax: int = 2048,
                 _loopback: Optional[List["PeerMailbox"]] = None, _rank: Optional[int] = None,
                 _world: Optional[int] = None):
        L = _lib.lib()
        self.pg = process_group
        self.rank = dist.get_rank(process_group) if _rank is None else _rank
        self.world = dist.get_world_size(process_group) if _world is None else _world
        self.cmax = int(cmax)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        if not 2 <= self.world <= 8:
            raise AfanError(f"PeerMailbox supports 2..8 ranks on one node, got {self.world}")
        self.nbytes = L.afan_bn_mailbox_bytes(self.world, self.cmax)
        with torch.cuda.device(self.device):
            own = ctypes.c_void_p()
            check(L.afan_p2p_alloc(ctypes.byref(own), self.nbytes), "afan_p2p_alloc")
        self._own = own.value
        self._opened = []
        self.state = torch.zeros(4, dtype=torch.int64, device=self.device)      # {seq, ticket, error, pad}
        self.peer_ptrs = (ctypes.c_void_p * 8)()
        if _loopback is None:
            self._exchange_handles()

    # ---- multi-process: cudaIpc handles over torch.distributed --------------------------------------
    def _exchange_handles(self):
        L = _lib.lib()
        handle = (ctypes.c_ubyte * 64)()
        check(L.afan_p2p_get_handle(self._own, handle), "afan_p2p_get_handle")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=self.pg)
        for r in range(self.world):
            if r == self.rank:
                self.peer_ptrs[r] = self._own
                continue
            buf = (ctypes.c_ubyte * 64)(*gathered[r].cpu().tolist())
            p = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                check(L.afan_p2p_open_handle(buf, ctypes.byref(p)), "afan_p2p_open_handle")
            self._opened.append(p.value)
            self.peer_ptrs[r] = p.value
        # no collective here: the caller agrees on success across ranks (trainer: all-reduce of an ok flag), which also
        # guarantees every peer finished opening before anyone can free its mailbox

    # ---- single-process loopback (tests): all "ranks" live on one GPU, peers are plain local pointers ----
    @classmethod
    def loopback(cls, world: int, device, cmax: int = 2048) -> List["PeerMailbox"]:
        boxes = [cls(None, device, cmax, _loopback=[], _rank=r, _world=world) for r in range(world)]
        for b in boxes:
            for r, other in enumerate(boxes):
                b.peer_ptrs[r] = other._own
        return boxes

    def error(self) -> bool:
        """True if some exchange timed out (a peer never arrived).  Synchronises the device."""
        return bool(int(self.state[2].item()))

    def check(self):
        if self.error():
            raise AfanError("fused BN statistics exchange timed out: a peer GPU did not reach the same BatchNorm call")

    def close(self):
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        for p in self._opened:
            L.afan_p2p_close_handle(p)
        self._opened = []
        if self._own:
            L.afan_p2p_free(self._own)
            self._own = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
