"""One replica per GPU: batch / stream sharding and the timing plumbing around it.

The hot path shards by independent units (clips or streams): rank ``r`` of
``world`` owns a contiguous slice of every batch and the streaming state of the streams placed on
it.  There is NO collective on the data path; ``torch.distributed`` is used only to line the ranks
up (barrier) and to take the maximum of a per-rank timing.  The reference has no counterpart (its
only collectives are training leftovers in ``utils/distributed.py:84-199``, out of scope).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split of ``total`` units: the first ``total % world`` ranks get one more."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class ReplicaGroup:
    rank: int
    world: int
    local_rank: int
    initialised_here: bool = False

    def barrier(self, device: Optional[torch.device] = None) -> None:
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        if device is not None and device.type == "cuda":
            torch.cuda.synchronize(device)

    def max_over_ranks(self, value: float, device: Optional[torch.device] = None) -> float:
        if self.world == 1:
            return float(value)
        import torch.distributed as dist
        t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def close(self) -> None:
        if self.initialised_here:
            import torch.distributed as dist
            dist.destroy_process_group()
            self.initialised_here = False


def init_replica_group(backend: Optional[str] = None, device: Optional[torch.device] = None) -> ReplicaGroup:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE (torchrun) and joins the process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    grp = ReplicaGroup(rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
            dist.init_process_group(backend, **kw)
            grp.initialised_here = True
    return grp


class StreamPlacement:
    """Sticky placement of stream ids: a stream lives on one rank for its lifetime (its state never
    moves), and inside the rank it owns one slot of the resident state pool."""

    def __init__(self, world: int, slots_per_rank: int):
        self.world, self.slots_per_rank = world, slots_per_rank
        self._slot: Dict[int, Tuple[int, int]] = {}
        self._free: List[List[int]] = [list(range(slots_per_rank - 1, -1, -1)) for _ in range(world)]

    def place(self, stream_id: int) -> Tuple[int, int]:
        """(rank, slot) of a stream, assigning the least-loaded rank on first sight."""
        if stream_id in self._slot:
            return self._slot[stream_id]
        rank = max(range(self.world), key=lambda r: (len(self._free[r]), -r))
        if not self._free[rank]:
            raise RuntimeError("no free stream slot on any rank")
        self._slot[stream_id] = (rank, self._free[rank].pop())
        return self._slot[stream_id]

    def release(self, stream_id: int) -> None:
        rank, slot = self._slot.pop(stream_id)
        self._free[rank].append(slot)

    def local_slots(self, rank: int, stream_ids) -> List[int]:
        """Pool slots (on ``rank``) of the given streams, in order; raises if one lives elsewhere."""
        out = []
        for s in stream_ids:
            r, slot = self.place(s)
            if r != rank:
                raise ValueError(f"stream {s} lives on rank {r}, not {rank}")
            out.append(slot)
        return out
