"""CPU, world_size 2, gloo: the sharding helpers and the one exchange of the path (all-gather of
embeddings + track ids) give every rank the same, complete, correctly re-labelled set."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyannote_video_b200.dist import shard_range, shard_shots, gather_embeddings


def test_shard_helpers():
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    cover = sorted(i for r in range(3) for i in range(*shard_range(7, r, 3)))
    assert cover == list(range(7))
    ranks = shard_shots([100, 50, 50, 200, 10, 90], 2)
    assert ranks == sorted(ranks) and set(ranks) == {0, 1}
    loads = [sum(n for n, r in zip([100, 50, 50, 200, 10, 90], ranks) if r == k) for k in range(2)]
    assert abs(loads[0] - loads[1]) <= 200


def _worker(rank, world, port, q, empty_rank=-1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    n = 0 if rank == empty_rank else 5 + 3 * rank
    emb = torch.randn(n, 128, generator=g)
    tracks = torch.arange(n) % (2 + rank)
    E, T = gather_embeddings(emb, tracks)
    q.put((rank, E.numpy(), T.numpy(), emb.numpy(), tracks.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_embeddings_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, E0, T0, e0, t0), (r1, E1, T1, e1, t1) = res
    assert np.array_equal(E0, E1) and np.array_equal(T0, T1)
    assert E0.shape == (5 + 8, 128)
    assert np.array_equal(E0[:5], e0) and np.array_equal(E0[5:], e1)
    stride = max(t0.max(), t1.max()) + 1
    assert np.array_equal(T0[:5], t0) and np.array_equal(T0[5:], t1 + stride)
    assert len(set(T0[:5]) & set(T0[5:])) == 0


def test_gather_embeddings_world3_with_an_empty_shard():
    """a rank without any face (a shard of frames with no detections) takes part in the exchange with zero rows"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 3, port, q, 1)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(3)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    E = res[0][1]
    assert E.shape == (5 + 0 + 11, 128)
    for r in res[1:]:
        assert np.array_equal(r[1], E) and np.array_equal(r[2], res[0][2])
    assert np.array_equal(E[:5], res[0][3]) and np.array_equal(E[5:], res[2][3])
    T = res[0][2]
    assert len(set(T[:5].tolist()) & set(T[5:].tolist())) == 0
