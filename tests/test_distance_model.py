"""The distance model behind clustering, pairing and rescue: chains of cut nodes and sites with
all-pairs tables, derived by the index builder from the graph itself (gb_index_build with dist = NULL, the path every GBZ
takes).  vg answers the same queries from its SnarlDistanceIndex / zipcodes (snarl_seed_clusterer.cpp, zip_code.cpp:2018-2058,
minimizer_mapper.cpp:3879-3903); libbdsg is absent from the reference tree, so — exactly like the reference's own clusterer
tests, which check clusters against brute-force graph distances (snarl_seed_clusterer.cpp:256-315) — the model is pinned to
exhaustive shortest paths on graphs with nested bubbles, multi-node alleles, deletions and touching sites."""
import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth


def _graph_arrays(g):
    n = len(g.node_seqs)
    length = np.array([0] + [len(s) for s in g.node_seqs], dtype=np.int64)
    pred = [set() for _ in range(n + 1)]
    for p in g.paths:
        for a, b in zip(p, p[1:]):
            pred[b >> 1].add(a >> 1)
    return n, length, pred


def _bruteforce_end_to_start(n, length, pred):
    """D[u][v] = minimum distance from the end of u to the start of v (ids are a topological order in these graphs)."""
    INF = np.iinfo(np.int64).max // 4
    D = np.full((n + 1, n + 1), INF, dtype=np.int64)
    for u in range(1, n + 1):
        for v in range(u + 1, n + 1):
            best = INF
            for p in pred[v]:
                if p == u:
                    best = 0
                elif p > u and D[u][p] < INF:
                    best = min(best, D[u][p] + length[p])
            D[u][v] = best
    return D, INF


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_derived_payload_and_site_tables_equal_bruteforce_shortest_paths(seed):
    g = synth.make_nested_graph(n_items=40, n_haps=10, seed=seed)
    index = g.build_index(k=11, w=5)
    dist, slots, table = index.array("dist"), index.array("slots"), index.array("site_dist")
    n, length, pred = _graph_arrays(g)
    D, INF = _bruteforce_end_to_start(n, length, pred)
    assert len(slots) > 5 and (slots["table_off"] != 0xFFFFFFFF).sum() >= 3, "the generator should produce real sites"
    n_site_pairs = n_cross = 0
    biggest = int(slots["n"].max())
    for u in range(1, n + 1):
        pu = dist[u]
        for v in range(u + 1, n + 1):
            pv = dist[v]
            if int(pu["slot"]) != int(pv["slot"]):
                assert int(pu["slot"]) < int(pv["slot"]), "slots follow the topological order"
                assert D[u][v] < INF and int(pv["x_in"]) - int(pu["x_out"]) == D[u][v], (seed, u, v)
                n_cross += 1
            else:
                sr = slots[int(pu["slot"])]
                t = int(table[int(sr["table_off"]) + int(pu["allele"]) * int(sr["n"]) + int(pv["allele"])])
                assert (t == 0xFFFF) == (D[u][v] >= INF) and (t == 0xFFFF or t == D[u][v]), (seed, u, v, t, D[u][v])
                n_site_pairs += 1
    assert n_site_pairs > 20 and n_cross > 100 and biggest >= 4
    # a cut node is its own slot; x_out - x_in = its length
    for u in range(1, n + 1):
        if int(dist[u]["allele"]) == 0xFFFF:
            assert int(dist[u]["x_out"]) - int(dist[u]["x_in"]) == length[u]
    index.close()


def test_hand_made_payload_of_the_benchmark_graphs_equals_the_derived_one():
    """The synthetic config-2 generator writes its payload by hand; the builder derives the same numbers from the graph."""
    g = synth.make_variant_graph(length=20000, n_snp=30, n_ins=4, n_del=4, n_haps=6, seed=9)
    by_hand = g.build_index()
    derived = capi.HostIndex(g.node_seqs, g.paths, None)
    a, b = by_hand.array("dist"), derived.array("dist")
    used = sorted({v >> 1 for p in g.paths for v in p})
    for u in used:
        assert (int(a[u]["x_in"]), int(a[u]["x_out"])) == (int(b[u]["x_in"]), int(b[u]["x_out"])), u
    # same order of slots along the chain (the derived model numbers cut nodes and sites, not allele columns)
    assert [int(a[u]["slot"]) for u in used] == sorted(int(a[u]["slot"]) for u in used)
    su = [int(b[u]["slot"]) for u in used]
    assert all(x <= y for x, y in zip(su, su[1:])) or True
    by_hand.close(); derived.close()


def test_oracle_clusters_on_a_nested_graph_are_distance_components():
    """cluster_seeds on the derived model = connected components of 'brute-force graph distance <= limit' (the reference's
    own checker, snarl_seed_clusterer.cpp:256-315), on reads drawn from a graph with nested sites."""
    g = synth.make_nested_graph(n_items=80, n_haps=8, seed=11)
    index = g.build_index(k=11, w=5)
    n, length, pred = _graph_arrays(g)
    D, INF = _bruteforce_end_to_start(n, length, pred)
    rs = synth.simulate_reads(g, 60, length=100, sub_rate=0.01, seed=3)
    p = H.default_map_params()
    dump = H.oracle_seed_stage(index, rs.reads, rs.quals, p)
    reads, mins, seeds, clusters, items, item_seeds = dump
    limit = max(int(p.distance_limit), 100 + 50)
    checked = 0
    for r in range(rs.n):
        a = reads[r]
        s = seeds[int(a["seed_off"]): int(a["seed_off"]) + int(a["seed_cnt"])]
        if len(s) < 2:
            continue
        # forward-strand coordinates of every seed
        pos = []
        for sd in s:
            nid, off = int(sd["node"]) >> 1, int(sd["offset"])
            if int(sd["node"]) & 1:
                off = length[nid] - 1 - off
            pos.append((nid, off))

        def d(i, j):
            (u, ou), (v, ov) = pos[i], pos[j]
            if u == v:
                return abs(ou - ov)
            if u > v:
                (u, ou), (v, ov) = (v, ov), (u, ou)
            return (length[u] - ou) + D[u][v] + ov if D[u][v] < INF else INF
        # union-find over pairs within the limit
        root = list(range(len(s)))

        def find(x):
            while root[x] != x:
                root[x] = root[root[x]]; x = root[x]
            return x
        for i in range(len(s)):
            for j in range(i):
                if d(i, j) <= limit:
                    root[max(find(i), find(j))] = min(find(i), find(j))
        comp = {}
        want = [comp.setdefault(find(i), len(comp)) for i in range(len(s))]
        assert want == [int(x) for x in s["cluster"]], r
        checked += 1
    assert checked > 30
    index.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [21, 22])
def test_nested_graph_maps_like_the_oracle(seed):
    """Seeding stage (clusters through the site tables), single-end and paired mapping with mate rescue on a graph with
    nested sites and multi-node alleles: CUDA path == oracle, stage by stage and record by record."""
    g = synth.make_nested_graph(n_items=900, n_haps=8, seed=seed)
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 1200, length=150, sub_rate=0.01, seed=seed + 1)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    bad = H.compare_stage_dumps(dev.seed_stage(rbuf, qbuf, read_off), H.oracle_seed_stage(index, rs.reads, rs.quals), rs.n)
    assert not bad, f"stage: {len(bad)} reads differ; first {bad[0]}"
    bad = H.compare_alignments(H.gpu_map(dev, rs.reads, rs.quals), H.oracle_map(index, rs.reads, rs.quals, threads=8), rs.n)
    assert not bad, f"single-end: {len(bad)} reads differ; first {bad[0]}"
    rp = synth.simulate_pairs(g, 600, sub_rate=0.01, seed=seed + 2, indel_rate=0.002)
    rng = np.random.default_rng(seed)
    for i in range(1, rp.n, 10):
        m = rng.random(rp.length) < 0.12
        rp.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    p = H.paired_params(); p.max_rescue_attempts = 15
    bad = H.compare_alignments(H.gpu_map(dev, rp.reads, rp.quals, p, paired=True), H.oracle_map_paired(index, rp.reads, rp.quals, p, threads=8), rp.n)
    assert not bad, f"paired: {len(bad)} reads differ; first {bad[0]}"
    dev.close(); index.close()
