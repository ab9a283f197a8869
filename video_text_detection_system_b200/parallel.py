"""Frame-sharded data parallelism over the GPUs of one box.

Frames are independent (pipeliine.py:104-139 touches only frames[i]), so rank r of W takes frames r, r+W, ...
with its own weight replica; the only exchange is one gather of the fixed-size detection records
(vtd_record, 128 B) and per-frame counts to rank 0.  NCCL on the GPU box, gloo in the CPU tests; the
functions below only need an initialised torch.distributed process group.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

RECORD_BYTES = 128


def shard_indices(n_frames: int, rank: int, world: int) -> List[int]:
    """Rank-strided shard: frames rank, rank+world, ..."""
    return list(range(rank, n_frames, world))


def frames_per_rank(n_frames: int, world: int) -> int:
    return (n_frames + world - 1) // world


class _CudaView:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes_as_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    """Zero-copy uint8 view of library-owned device memory (the packed records the kernels wrote)."""
    return torch.as_tensor(_CudaView(ptr, nbytes), device=device)


def gather_records(records: torch.Tensor, counts: torch.Tensor, dst: int = 0
                   ) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """records: uint8 [F, Kmax*128] (F frames of this rank, padded to the same F on every rank),
    counts: int32 [F].  Returns on `dst` (records [W,F,Kmax*128], counts [W,F]), elsewhere None.
    One collective per tensor: the payload is a few KB per frame, so this is latency- not bandwidth-bound."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    if world == 1:
        return records.unsqueeze(0), counts.unsqueeze(0)
    if dist.get_backend() == "nccl":
        out_r = torch.empty((world,) + tuple(records.shape), dtype=records.dtype, device=records.device)
        out_c = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
        dist.all_gather_into_tensor(out_r, records.contiguous())
        dist.all_gather_into_tensor(out_c, counts.contiguous())
        return (out_r, out_c) if rank == dst else None
    lr = [torch.empty_like(records) for _ in range(world)] if rank == dst else None
    lc = [torch.empty_like(counts) for _ in range(world)] if rank == dst else None
    dist.gather(records.contiguous(), lr, dst=dst)
    dist.gather(counts.contiguous(), lc, dst=dst)
    if rank != dst:
        return None
    return torch.stack(lr), torch.stack(lc)


def merge_gathered(records: np.ndarray, counts: np.ndarray, n_frames: int, kmax: int, record_dtype) -> List[np.ndarray]:
    """Undo the rank-strided sharding on rank 0: returns, per global frame index, its record array."""
    world, per = counts.shape
    recs = records.reshape(world, per, kmax * RECORD_BYTES)
    out: List[np.ndarray] = []
    for g in range(n_frames):
        r, i = g % world, g // world
        row = np.frombuffer(recs[r, i].tobytes(), dtype=record_dtype, count=kmax)
        out.append(row[:int(counts[r, i])].copy())
    return out
