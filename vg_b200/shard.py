"""Read sharding and record emission across ranks (one process per GPU).

Giraffe reads are independent once the fragment-length distribution is fixed
(giraffe_main.cpp:2416-2459), so rank r maps pairs [lo, hi) of the batch with its own replica of
the index; the only exchange is the gather of every rank's whole records (headers, mappings, edits:
gather_records) on rank 0 for emission, plus two doubles when the distribution is
learned (share_fragment_distribution).  With max_multimaps > 1 the header array simply holds
n * max_multimaps records per rank (absent ranks included; the emitters skip them)."""
from __future__ import annotations


def shard_pairs(n_pairs: int, rank: int, world: int):
    """Contiguous, balanced pair range of `rank` (first n_pairs % world ranks get one extra pair)."""
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def gather_headers(headers, rank: int, world: int, dst: int = 0):
    """Headers only (round 1's gather, kept for callers that need just scores / MAPQs; gather_records moves whole records):
    gather per-rank [n_i, 32] uint8 header tensors on `dst`, padded to the largest shard.
    Returns the list of per-rank tensors (trimmed) on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    n = torch.tensor([headers.shape[0]], dtype=torch.int64, device=headers.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(int(s.item()) for s in sizes))
    padded = headers
    if headers.shape[0] < m:
        padded = torch.zeros((m, headers.shape[1]), dtype=headers.dtype, device=headers.device)
        padded[: headers.shape[0]] = headers
    out = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, out, dst=dst)
    if rank != dst:
        return None
    return [o[: int(s.item())] for o, s in zip(out, sizes)]


def gather_records(headers, mappings, edits, rank: int, world: int, dst: int = 0, counts_group=None, out=None):
    """The emission gather of SURVEY §8(e): every rank's WHOLE records — [n_i, 32] uint8 headers, [m_i, 8] uint8
    mappings, [e_i] int32 edits, the dense pools gb_map_*_batch returns — land on `dst`, exact sizes, no padding.

    Sizes travel once as one 3-number all_gather (on `counts_group` when given: a host-side gloo group keeps that
    exchange off the CUDA stream, so nothing here synchronises the device); the payload is one batched isend / irecv per
    array, enqueued on the CURRENT stream and returned as outstanding requests, so the caller can run the next batch while
    the records of this one move (NCCL over NVLink for CUDA tensors, gloo for CPU tensors).

    Returns (requests, parts): `parts` on dst is a list of (headers, mappings, edits) per rank whose header.mapping_off /
    edit_off still index that rank's own pools (merge_records rebases them); None elsewhere.  `out`, on dst, optionally
    supplies preallocated receive buffers [(headers, mappings, edits)] per rank (capacity >= the incoming sizes)."""
    import torch
    import torch.distributed as dist
    dev = headers.device
    mine = torch.tensor([headers.shape[0], mappings.shape[0], edits.shape[0]], dtype=torch.int64)
    if counts_group is not None:
        sizes = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, mine, group=counts_group)
    else:
        sizes = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, mine.to(dev))
        sizes = [s.cpu() for s in sizes]
    sizes = [tuple(int(x) for x in s) for s in sizes]
    ops, parts = [], None
    if rank == dst:
        parts = []
        for r in range(world):
            nh, nm, ne = sizes[r]
            if r == dst:
                parts.append((headers, mappings, edits))
                continue
            if out is not None:
                bh, bm, be = out[r]
                bufs = (bh[:nh], bm[:nm], be[:ne])
            else:
                bufs = (torch.empty((nh, 32), dtype=torch.uint8, device=dev), torch.empty((nm, 8), dtype=torch.uint8, device=dev),
                        torch.empty((ne,), dtype=torch.int32, device=dev))
            parts.append(bufs)
            ops += [dist.P2POp(dist.irecv, b, r) for b in bufs if b.numel()]
    else:
        ops += [dist.P2POp(dist.isend, t, dst) for t in (headers, mappings, edits) if t.numel()]
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return reqs, parts


def merge_records(parts):
    """Concatenate the per-rank records of gather_records into one set of pools with globally valid offsets (numpy in,
    numpy out: (alignment records, mappings, edits)).  read_id stays the rank-local read index; the caller adds its shard
    base if it wants global ids."""
    import numpy as np
    from . import capi
    hs, ms, es = [], [], []
    mbase = ebase = 0
    for h, m, e in parts:
        h = np.ascontiguousarray(h).reshape(-1).view(capi.alignment_dt).copy()
        m = np.ascontiguousarray(m).reshape(-1).view(capi.mapping_dt)
        e = np.ascontiguousarray(e).reshape(-1).view(np.uint32)
        h["mapping_off"] += np.uint32(mbase); h["edit_off"] += np.uint32(ebase)
        hs.append(h); ms.append(m); es.append(e)
        mbase += len(m); ebase += len(e)
    return np.concatenate(hs), np.concatenate(ms), np.concatenate(es)


def share_fragment_distribution(distribution, rank: int, world: int, src: int = 0, device=None):
    """The one piece of shared state of a paired job (minimizer_mapper.hpp:696): rank `src` learns the
    fragment length distribution from the head of the input (gb_map_paired_job stops training after
    maximum_sample_size unambiguous pairs) and broadcasts (mean, stdev) as two doubles; every other rank
    forces its own gb_fragment_distribution to them (force_parameters, mapper.cpp:5250), after which all
    ranks map their shards independently.  Returns (mean, stdev).  Raises if `src` has not finalized."""
    import torch
    import torch.distributed as dist
    t = torch.zeros(3, dtype=torch.float64, device=device)
    if rank == src:
        t[0], t[1], t[2] = distribution.mean(), distribution.std_dev(), 1.0 if distribution.is_finalized() else 0.0
    dist.broadcast(t, src=src)
    mean, stdev, finalized = (float(x) for x in t.cpu())
    if finalized != 1.0:
        raise RuntimeError("share_fragment_distribution: the source rank's distribution is not finalized")
    if rank != src:
        distribution.force_parameters(mean, stdev)
    return mean, stdev
