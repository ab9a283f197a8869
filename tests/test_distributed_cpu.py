"""CPU, world_size 2, gloo: the frame-sharding + record-gather path of video_text_detection_system_b200/parallel.py
(the NCCL path of bench.py --gpus N runs the same functions on device tensors)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_text_detection_system_b200 import _lib, parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_records(frame_idx, kmax):
    """Deterministic records for global frame `frame_idx`: count = frame_idx % (kmax+1)."""
    rec = np.zeros(kmax, _lib.RECORD_DTYPE)
    cnt = frame_idx % (kmax + 1)
    for k in range(cnt):
        rec[k]["frame"] = frame_idx
        rec[k]["bbox"] = [k, frame_idx, k + 20, frame_idx + 12]
        rec[k]["det_conf"] = 0.5 + 0.01 * k
        rec[k]["len"] = 2
        rec[k]["ids"][:2] = [1 + k % 90, 2 + frame_idx % 90]
    return rec, cnt


def _worker(rank, world, port, n_frames, kmax, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_indices(n_frames, rank, world)
        per = parallel.frames_per_rank(n_frames, world)
        recs = np.zeros((per, kmax), _lib.RECORD_DTYPE)
        cnts = np.zeros(per, np.int32)
        for i, g in enumerate(mine):
            recs[i], cnts[i] = _fake_records(g, kmax)
        r = torch.from_numpy(recs.view(np.uint8).reshape(per, kmax * 128).copy())
        c = torch.from_numpy(cnts)
        got = parallel.gather_records(r, c, dst=0)
        if rank == 0:
            assert got is not None
            merged = parallel.merge_gathered(got[0].numpy(), got[1].numpy(), n_frames, kmax, _lib.RECORD_DTYPE)
            ok = True
            for g in range(n_frames):
                want, cnt = _fake_records(g, kmax)
                ok = ok and len(merged[g]) == cnt and merged[g].tobytes() == want[:cnt].tobytes()
            open(out_path, "w").write("ok" if ok else "mismatch")
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8])
def test_shard_and_gather_world2(tmp_path, n_frames):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), n_frames, 5, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


@pytest.mark.parametrize("n_frames", [3, 9])
def test_shard_and_gather_world4_with_idle_and_ragged_ranks(tmp_path, n_frames):
    """World size 4: 3 frames leave rank 3 without work (it still takes part in the gather with a zero count),
    9 frames give rank 0 one frame more than the others."""
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(4, _free_port(), n_frames, 5, out), nprocs=4, join=True)
    assert open(out).read() == "ok"


def _worker_packed(rank, world, port, n_frames, kmax, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_indices(n_frames, rank, world)
        per = parallel.frames_per_rank(n_frames, world)
        recs = np.zeros((per, kmax), _lib.RECORD_DTYPE)
        cnts = np.zeros(per, np.int32)
        for i, g in enumerate(mine):
            recs[i], cnts[i] = _fake_records(g, kmax)
        # the library's layout: the counts directly follow the records (vtd_get_records)
        block = torch.from_numpy(np.concatenate([recs.view(np.uint8).reshape(-1), cnts.view(np.uint8)]))
        assert block.numel() == parallel.packed_bytes(per, kmax)
        got = parallel.gather_packed(block, dst=0)
        if rank == 0:
            r, c = parallel.split_packed(got, per, kmax)
            merged = parallel.merge_gathered(r, c, n_frames, kmax, _lib.RECORD_DTYPE)
            ok = int(c.sum()) == sum(g % (kmax + 1) for g in range(n_frames))
            for g in range(n_frames):
                want, cnt = _fake_records(g, kmax)
                ok = ok and len(merged[g]) == cnt and merged[g].tobytes() == want[:cnt].tobytes()
            open(out_path, "w").write("ok" if ok else "mismatch")
        else:
            assert got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 8), (2, 7), (4, 9)])
def test_single_collective_gather_of_packed_blocks(tmp_path, world, n_frames):
    """bench.py's N>1 path: one collective per step carrying records and counts in one block."""
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker_packed, args=(world, _free_port(), n_frames, 5, out), nprocs=world, join=True)
    assert open(out).read() == "ok"


def test_shard_indices_cover_everything():
    for n in (0, 1, 5, 16, 3000):
        for w in (1, 2, 4, 8):
            got = sorted(i for r in range(w) for i in parallel.shard_indices(n, r, w))
            assert got == list(range(n))
            assert all(len(parallel.shard_indices(n, r, w)) <= parallel.frames_per_rank(n, w) for r in range(w))
