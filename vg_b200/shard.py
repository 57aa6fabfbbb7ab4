"""Read sharding and record emission across ranks (one process per GPU).

Giraffe reads are independent once the fragment-length distribution is fixed
(giraffe_main.cpp:2416-2459), so rank r maps pairs [lo, hi) of the batch with its own replica of
the index; the only exchange is the gather of the fixed-width 32-byte alignment headers on rank 0
for emission."""
from __future__ import annotations


def shard_pairs(n_pairs: int, rank: int, world: int):
    """Contiguous, balanced pair range of `rank` (first n_pairs % world ranks get one extra pair)."""
    base, extra = divmod(n_pairs, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def gather_headers(headers, rank: int, world: int, dst: int = 0):
    """Gather per-rank [n_i, 32] uint8 header tensors on `dst`, padded to the largest shard.
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
