"""SyncBN exchanges over NVLink peer memory (``cgnn_peer_exchange``) instead of NCCL collectives.

Under data parallelism every BatchNorm layer needs the other ranks' statistics in the forward pass and their backward sums
in the backward pass: a few hundred bytes each, but six dependent collectives per training step and model sit between
two layer kernels, at ~30 us apiece through NCCL on eight GPUs.  ``PeerExchange`` keeps one small symmetric buffer per
rank (``torch.distributed._symmetric_memory`` - torch is the plumbing that maps the peers' memory), and one kernel per
exchange writes this rank's record into every peer's buffer, waits for theirs and merges them in rank order.

Falls back to the NCCL path (and says so once) when symmetric memory cannot be set up: not an NVLink box, gloo group,
torch without the module.
"""

from __future__ import annotations

import warnings
from typing import Optional

import torch

from . import _engine

__all__ = ["PeerExchange", "peer_exchange_for"]

_SLOTS = 64          # physical slots: 2 per logical exchange (alternating), 2 exchanges per layer, up to 16 layers
_REC_CAP = 520       # doubles per record: 1 + 2 * 256 channels, rounded up


class PeerExchange:
    def __init__(self, group, device: torch.device):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.device = device
        flags_doubles = (_SLOTS * self.world + 1) // 2
        self.buf = symm.empty(_SLOTS * self.world * _REC_CAP + flags_doubles, dtype=torch.float64, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.ptrs_dev = int(self.hdl.buffer_ptrs_dev)
        self.error = torch.zeros(1, dtype=torch.int32, device=device)
        self.seq: dict = {}
        torch.cuda.synchronize(device)
        dist.barrier(self.group)          # every rank's buffer is zeroed and mapped before the first record lands in it

    def _slot(self, logical: int):
        n = self.seq.get(logical, 0) + 1
        self.seq[logical] = n
        if 2 * logical + 1 >= _SLOTS:
            raise RuntimeError("PeerExchange: too many layers for the slot table")
        return 2 * logical + (n & 1), n

    def exchange(self, logical: int, mode: int, mine: torch.Tensor, channels: int, out: torch.Tensor) -> torch.Tensor:
        eng = _engine.engine_for(mine)
        slot, seq = self._slot(logical)
        n = mine.numel()
        if n > _REC_CAP:
            raise RuntimeError("PeerExchange: record too large")
        eng._call("cgnn_peer_exchange", self.ptrs_dev, self.rank, self.world, slot, _SLOTS, _REC_CAP, seq & 0xFFFFFFFF or 1, mode,
                  mine.data_ptr(), n, channels, out.data_ptr(), self.error.data_ptr(), eng.stream())
        return out

    def merge_stats(self, layer: int, stats: torch.Tensor, channels: int) -> torch.Tensor:
        """{count, mean[C], M2[C]} of this rank -> the merged record of all ranks (forward)."""
        return self.exchange(2 * layer, 0, stats.contiguous(), channels, torch.empty_like(stats))

    def sum(self, layer: int, sums64: torch.Tensor) -> torch.Tensor:
        """[2, C] float64 backward sums of this rank -> their sum over all ranks (backward)."""
        src = sums64.contiguous()
        return self.exchange(2 * layer + 1, 1, src, 0, torch.empty_like(src))

    def check(self) -> None:
        """Raises if an exchange timed out (one host read; call at an epoch boundary)."""
        if int(self.error.item()):
            raise RuntimeError("PeerExchange: a peer's record did not arrive (timeout) - results since then are invalid")


_CACHE: dict = {}


def peer_exchange_for(group, device: torch.device) -> Optional[PeerExchange]:
    """The exchange of (group, device), set up on first use; None when it cannot be (the caller then uses NCCL)."""
    import torch.distributed as dist
    key = (id(group) if group is not None else 0, device.index)
    if key in _CACHE:
        return _CACHE[key]
    px = None
    try:
        g = group if group is not None else dist.group.WORLD
        if device.type == "cuda" and dist.get_backend(g) == "nccl":
            px = PeerExchange(group, device)
    except Exception as exc:        # no NVLink peer access, old torch, ...: stay on the collectives
        warnings.warn(f"connectome_gnn: SyncBN over peer memory is not available ({exc!r}); using NCCL collectives")
        px = None
    _CACHE[key] = px
    return px
