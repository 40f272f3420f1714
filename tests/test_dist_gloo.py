"""world_size-2 gloo tests (CPU) of the multi-GPU bookkeeping in node2vec_b200/dist.py:
start-vertex sharding, shard layout, vocabulary reduction, model averaging."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from node2vec_b200 import dist as n2v_dist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert n2v_dist.world() == (rank, world_size)
        # walks shard by start vertex
        start = torch.arange(0, 101, dtype=torch.int32)
        mine = n2v_dist.shard_start_vertices(start, rank, world_size)
        # shard layout: rank r holds 10 * (r + 1) walks, sees ids up to 50 + r
        off, total, rows = n2v_dist.shard_layout(10 * (rank + 1), 50 + rank + 1)
        assert total == sum(10 * (r + 1) for r in range(world_size))
        assert off == sum(10 * (r + 1) for r in range(rank)) and rows == 50 + world_size
        # vocabulary: counts add up, first positions take the minimum
        counts = torch.full((6,), rank + 1, dtype=torch.int64)
        first = torch.tensor([5, 9, 1, 7, 2 ** 62, 3], dtype=torch.int64) + 100 * rank
        n2v_dist.reduce_vocab(counts, first)
        assert counts.tolist() == [sum(r + 1 for r in range(world_size))] * 6
        assert first.tolist() == [5, 9, 1, 7, 2 ** 62, 3]
        # model averaging
        a = torch.full((4, 8), float(rank), dtype=torch.float32)
        b = torch.arange(8, dtype=torch.float32) * (rank + 1)
        n2v_dist.average_tables((a, b))
        mean_rank = (world_size - 1) / 2.0
        assert torch.allclose(a, torch.full((4, 8), mean_rank))
        assert torch.allclose(b, torch.arange(8, dtype=torch.float32) * (mean_rank + 1))
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), mine.numpy())
    finally:
        dist.destroy_process_group()


def _fd_worker(rank, world_size, port, out_dir):
    """The SCM_RIGHTS hand-over used for the peer-shareable buffer handles: every rank ends up with
    its OWN copies of every other rank's file descriptors."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from node2vec_b200.graph import _exchange_fds
        fds = []
        for k in range(3):
            path = os.path.join(out_dir, f"r{rank}_f{k}.txt")
            with open(path, "w") as f:
                f.write(f"rank {rank} file {k}")
            fds.append(os.open(path, os.O_RDONLY))
        got = _exchange_fds(fds, rank, world_size, None)
        assert sorted(got) == [p for p in range(world_size) if p != rank]
        for p, peer_fds in got.items():
            assert len(peer_fds) == 3 and all(fd not in fds for fd in peer_fds)
            for k, fd in enumerate(peer_fds):
                assert os.pread(fd, 100, 0).decode() == f"rank {p} file {k}"   # shared open-file description: no seek
                os.close(fd)
        for fd in fds:
            os.close(fd)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_fd_exchange_three_ranks(tmp_path):
    mp.spawn(_fd_worker, args=(3, _free_port(), str(tmp_path)), nprocs=3, join=True)


def test_two_rank_bookkeeping(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    shards = [np.load(tmp_path / f"shard{r}.npy") for r in range(2)]
    assert np.array_equal(np.concatenate(shards), np.arange(101))      # disjoint, ordered, complete
    assert abs(len(shards[0]) - len(shards[1])) <= 1


def test_shard_bounds_properties():
    for n in (0, 1, 7, 100, 1001):
        for w in (1, 2, 3, 8):
            b = n2v_dist.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_defaults():
    assert n2v_dist.world() == (0, 1)
    assert n2v_dist.shard_layout(12, 40) == (0, 12, 40)
    t = torch.ones(3)
    n2v_dist.average_tables((t,))
    assert t.tolist() == [1.0, 1.0, 1.0]


def _rmat_worker(rank, world_size, port, out_dir):
    """synth.rmat_partition_device on CPU tensors over gloo: every rank ends up with the arcs whose
    source it owns, and the union is the graph a single process generates (strong scaling walks ONE
    fixed graph whatever the number of ranks)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from node2vec_b200 import synth
        src, dst = synth.rmat_partition_device(10, 8, rank, world_size, torch.device("cpu"))
        S = (1024 + world_size - 1) // world_size
        assert bool(((src >= rank * S) & (src < (rank + 1) * S)).all())
        np.save(os.path.join(out_dir, f"arcs{rank}.npy"), np.stack([src.numpy(), dst.numpy()]))
    finally:
        dist.destroy_process_group()


def test_partitioned_rmat_is_the_same_graph_for_every_world_size(tmp_path):
    from node2vec_b200 import synth
    mp.spawn(_rmat_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"arcs{r}.npy") for r in range(2)]
    got = np.concatenate(parts, axis=1)
    one_s, one_d = synth.rmat_partition_device(10, 8, 0, 1, torch.device("cpu"))
    want = np.stack([one_s.numpy(), one_d.numpy()])
    assert got.shape == want.shape and np.array_equal(got, want)          # both sorted by (src, dst): rank ranges are ascending
    key = got[0].astype(np.int64) << 32 | got[1]
    assert len(np.unique(key)) == len(key) and (got[0] != got[1]).all()   # simple, loop-free
    rev = got[1].astype(np.int64) << 32 | got[0]
    assert np.array_equal(np.sort(rev), np.sort(key))                     # symmetric
