"""N > 1 host logic on CPU: balanced pair sharding, the header gather and the fragment-distribution
broadcast, world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vg_b200 import capi, shard


def test_shard_pairs_is_a_partition():
    for n in (0, 1, 7, 10, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_pairs(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_pairs(n_pairs, rank, world)
    hdr = np.zeros(2 * (hi - lo), dtype=capi.alignment_dt)
    hdr["read_id"] = np.arange(2 * lo, 2 * hi)
    hdr["score"] = 100 + rank
    t = torch.from_numpy(hdr.view(np.uint8).reshape(-1, 32).copy())
    parts = shard.gather_headers(t, rank, world)
    if rank == 0:
        merged = np.concatenate([p.numpy().reshape(-1).view(capi.alignment_dt) for p in parts])
        q.put((merged["read_id"].tolist(), merged["score"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_headers_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_pairs = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    ids, scores = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == list(range(2 * n_pairs))
    assert scores == [100] * 8 + [101] * 6


def _frag_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    f = capi.FragmentDistribution(1000, 1000, 0.95)
    if rank == 0:
        rng = np.random.default_rng(4)
        for v in rng.normal(410, 55, size=1000).astype(np.int64).tolist():
            f.register_fragment_length(v)
    mean, sd = shard.share_fragment_distribution(f, rank, world)
    q.put((rank, mean, sd, f.mean(), f.std_dev(), f.is_finalized(), f.curr_sample_size()))
    dist.barrier()
    dist.destroy_process_group()


def test_fragment_distribution_broadcast_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_frag_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, s0, fm0, fs0, fin0, n0), (r1, m1, s1, fm1, fs1, fin1, n1) = got
    assert (m0, s0) == (m1, s1) == (fm0, fs0) == (fm1, fs1)        # bit-identical on both ranks
    assert fin0 and fin1 and n0 == 1000 and n1 == 0                # rank 1 is forced, it never sampled
    assert 400 < m0 < 420 and 45 < s0 < 65
