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


def packed_bytes(frames: int, kmax: int) -> int:
    """Size of one rank's result block: `frames` x `kmax` records followed by `frames` int32 counts (the layout
    vtd_get_records documents: the counts directly follow the records in device memory)."""
    return frames * kmax * RECORD_BYTES + frames * 4


def gather_packed(block: torch.Tensor, dst: int = 0, out: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """ONE collective per step: every rank contributes its result block (uint8 [packed_bytes]); returns on `dst` the
    [W, packed_bytes] tensor (written into `out` if given), elsewhere None.  NCCL: all_gather_into_tensor (a single
    kernel; the payload is ~130 KB per rank, latency-bound); gloo: gather."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        if out is not None:
            out[0].copy_(block)
            return out
        return block.unsqueeze(0)
    if dist.get_backend() == "nccl":
        if out is None:
            out = torch.empty((world, block.numel()), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(out, block)
        return out if rank == dst else None
    parts = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block.contiguous(), parts, dst=dst)
    if rank != dst:
        return None
    res = torch.stack(parts)
    if out is not None:
        out.copy_(res)
        return out
    return res


def split_packed(blocks, frames: int, kmax: int) -> Tuple[np.ndarray, np.ndarray]:
    """[W, packed_bytes] uint8 (tensor or array, host) -> (records [W, frames, kmax*128] uint8, counts [W, frames] int32)."""
    a = blocks.numpy() if isinstance(blocks, torch.Tensor) else np.asarray(blocks)
    w = a.shape[0]
    nrec = frames * kmax * RECORD_BYTES
    recs = a[:, :nrec].reshape(w, frames, kmax * RECORD_BYTES)
    cnts = np.ascontiguousarray(a[:, nrec:nrec + frames * 4]).view(np.int32).reshape(w, frames)
    return recs, cnts


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
