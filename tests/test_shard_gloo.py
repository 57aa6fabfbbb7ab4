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


def _records_worker(rank, world, port, q):
    """Each rank maps its shard of a paired batch on the CPU oracle (the test has no GPU), packs the dense record pools the
    way gb_map_paired_batch returns them, and the ranks gather the WHOLE records on rank 0."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as H
    from vg_b200 import synth
    g = synth.make_variant_graph(length=20000, n_snp=30, n_ins=4, n_del=4, n_haps=4, seed=9)
    index = g.build_index()
    rs = synth.simulate_pairs(g, 31, sub_rate=0.01, seed=5, indel_rate=0.002)
    p = H.paired_params()
    lo, hi = shard.shard_pairs(rs.n // 2, rank, world)
    reads, quals = rs.reads[2 * lo: 2 * hi], rs.quals[2 * lo: 2 * hi]
    aln, maps, edits, status, _ = H.oracle_map_paired(index, reads, quals, p, threads=2)
    # dense pools (the oracle writes fixed-stride pools; compact them like the library does)
    dm, de = [], []
    aln = aln.copy()
    for a in aln:
        m0, e0 = int(a["mapping_off"]), int(a["edit_off"])
        a["mapping_off"], a["edit_off"] = len(dm), len(de)
        dm += list(maps[m0: m0 + int(a["n_mappings"])]); de += list(edits[e0: e0 + int(a["n_edits"])])
    dm = np.array(dm, dtype=capi.mapping_dt) if dm else np.zeros(0, dtype=capi.mapping_dt)
    de = np.array(de, dtype=np.uint32)
    th = torch.from_numpy(aln.view(np.uint8).reshape(-1, 32).copy())
    tm = torch.from_numpy(dm.view(np.uint8).reshape(-1, 8).copy())
    te = torch.from_numpy(de.view(np.int32).copy())
    reqs, parts = shard.gather_records(th, tm, te, rank, world)
    for r in reqs:
        r.wait()
    if rank == 0:
        # rank 0 emits ONE GAM stream for the whole batch from the merged records
        for i, (h, m, e) in enumerate(parts):
            base = 2 * shard.shard_pairs(rs.n // 2, i, world)[0]
            hv = h.numpy().reshape(-1).view(capi.alignment_dt); hv["read_id"] += np.uint32(base)
        ga, gm, ge = shard.merge_records([(h.numpy(), m.numpy(), e.numpy()) for h, m, e in parts])
        rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
        gam = capi.emit_text("gam", index.view, ga, gm, ge, rbuf, qbuf, read_off)
        whole = H.oracle_map_paired(index, rs.reads, rs.quals, p, threads=2)
        q.put((gam, [H.decode_alignment(whole[0][i], whole[1], whole[2]) for i in range(rs.n)], [len(x[0]) for x in parts]))
    dist.barrier()
    dist.destroy_process_group()


def test_whole_records_gathered_on_rank0_world_size_2_gloo():
    """Rank 0's GAM holds every rank's alignments WITH their paths (mappings and edits), equal to a single-process run."""
    import test_gam as TG
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_records_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gam, want, sizes = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sizes == [32, 30]                                   # 16 + 15 pairs, exact, unpadded
    msgs, _ = TG.read_stream(gam)
    assert len(msgs) == 62
    n_paths = 0
    for i, m in enumerate(msgs):
        a = TG.decode_alignment(m)
        score, mapq, path = want[i]
        assert a["name"] == f"read{i}" and a["score"] == score
        assert [(mp["node_id"], mp["offset"], mp["is_reverse"]) for mp in a["path"]] == [(n >> 1, o, bool(n & 1)) for n, o, _ in path]
        assert [len(mp["edits"]) for mp in a["path"]] == [len(e) for _, _, e in path]
        n_paths += bool(a["path"])
    assert n_paths >= 58                                       # the second rank's reads carry their paths too
