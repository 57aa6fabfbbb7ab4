"""The candidate side of the chaining seam: gb_chain_candidates_batch enumerates, for every destination seed, the seeds it can
be reached from within a lookback and the minimum graph distance between the two positions — what the reference's
zip_tree_transition_iterator reads off its zip-code tree (chain_items.cpp:116-260, ZipCodeTree::find_distances).

Pins: the "Check iterator" expectations of the reference's zip-code-tree tests on DAGs (unittest/zip_code_tree.cpp: one node
:220-245, two node chain :344-388, simple bubbles in chains :599-627 / :706-734 / :758-790 / :836-862, nested bubbles
:989-1003), read as sets of (source, destination, distance) — the tests state them for whichever orientation the tree took,
a reversed view (dest d sees source s) being the forward transition d -> s.  The reference validates its trees against
SnarlDistanceIndex::minimum_distance (validate_zip_forest); here every graph, the non-simple DAG of :1246-1275 included, is
also checked against exhaustive shortest paths between positions."""
import ctypes as C
import itertools

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth

CD = capi.chain_candidate_dt
INF = 2 ** 64 - 1


def build(node_lens, edges):
    """Index of a DAG given as node lengths (ids 1.., topological) and edges; haplotypes = every source-to-sink walk."""
    n = len(node_lens)
    succ = {i: [b for a, b in edges if a == i] for i in range(1, n + 1)}
    has_pred = {b for _, b in edges}
    walks = []

    def go(path):
        if not succ[path[-1]]:
            walks.append(list(path)); return
        for b in succ[path[-1]]:
            go(path + [b])
    for s in range(1, n + 1):
        if s not in has_pred:
            go([s])
    rng = np.random.default_rng(len(edges) + n)
    seqs = ["".join(rng.choice(list("ACGT"), size=l)) for l in node_lens]
    return capi.HostIndex(seqs, [[2 * v for v in w] for w in walks], k=5, w=3), succ


def brute(node_lens, succ, seeds, limit):
    """Minimum distance between forward positions by exhaustive search: (from, to, d) for every ordered pair."""
    n = len(node_lens)
    # end-of-u -> start-of-v minimum distances
    D = {}
    for u in range(n, 0, -1):
        D[u] = {}
        for s in succ[u]:
            D[u][s] = 0
            for v, d in D[s].items():
                D[u][v] = min(D[u].get(v, 1 << 60), d + node_lens[s - 1])
    out = set()
    for (i, (ni, oi)), (j, (nj, oj)) in itertools.permutations(enumerate(seeds), 2):
        if ni == nj:
            d = oj - oi if oj >= oi else None
        else:
            d = (node_lens[ni - 1] - oi) + D[ni][nj] + oj if nj in D[ni] else None
        if d is not None and d <= limit:
            out.add((i, j, d))
    return out


def oracle_candidates(index, seeds, limit=INF):
    lib = H.oracle_lib()
    lib.oracle_chain_candidates.restype = C.c_uint64
    lib.oracle_chain_candidates.argtypes = [C.POINTER(capi.FlatIndex), C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
    pos = np.ascontiguousarray(np.array(list(seeds) + [(0, 0)], dtype=np.uint32).reshape(-1))
    cap = len(seeds) * len(seeds) + 1
    out = np.zeros(cap, dtype=CD)
    n = lib.oracle_chain_candidates(C.byref(index.view), len(seeds), capi.ptr(pos), limit, capi.ptr(out), cap)
    return out[:n]


def as_set(c):
    return {(int(x["from"]), int(x["to"]), int(x["graph_distance"])) for x in c}


ONE = ([3], [])
TWO = ([3, 7], [(1, 2)])
BUBBLES = ([3, 7, 3, 3, 3, 3], [(1, 2), (1, 3), (2, 3), (3, 4), (3, 5), (4, 6), (5, 6)])
NESTED = ([3, 7, 7, 10, 3, 3, 13], [(1, 2), (1, 5), (2, 3), (2, 4), (3, 4), (4, 5), (5, 6), (5, 7), (6, 7)])
NON_SIMPLE = ([3, 4, 6, 2, 2, 3, 3, 17], [(1, 2), (1, 3), (2, 3), (3, 4), (3, 7), (3, 8), (4, 5), (4, 6), (5, 6), (6, 7), (7, 8)])
TWO_CHAINS = ([3, 7, 3, 7], [(1, 2), (3, 4)])

# (graph, seeds as (node id, is_reverse, offset), limit, expected transitions or None, extra check)
CASES = {
    "one node, three seeds (:220-245)": (ONE, [(1, 0, 0), (1, 0, 1), (1, 0, 2)], INF, {(0, 1, 1), (1, 2, 1), (0, 2, 2)}),
    "two node chain (:344-368)": (TWO, [(1, 0, 0), (1, 0, 1), (2, 0, 2)], INF, {(0, 1, 1), (1, 2, 4), (0, 2, 5)}),
    "two node chain, distance limit 2 (:370-388)": (TWO, [(1, 0, 0), (1, 0, 1), (2, 0, 2)], 2, {(0, 1, 1)}),
    "two chains: components do not see each other (:446-455)": (TWO_CHAINS, [(1, 0, 0), (3, 0, 0)], INF, set()),
    "bubbles, seeds on chain nodes (:599-623)": (BUBBLES, [(1, 0, 0), (3, 0, 0), (6, 0, 0)], INF, {(0, 1, 3), (1, 2, 6), (0, 2, 9)}),
    "bubbles, one seed on the snarl (:706-730)": (BUBBLES, [(1, 0, 0), (2, 0, 1), (6, 0, 0)], INF, {(0, 1, 4), (1, 2, 12), (0, 2, 9)}),
    "bubbles, reverse strand (:758-770)": (BUBBLES, [(1, 1, 0), (2, 1, 1)], INF, {(1, 0, 6)}),
    "bubbles, reverse strand, distance limit 2 (:777-790)": (BUBBLES, [(1, 1, 0), (2, 1, 1)], 2, set()),
    "bubbles, two children of a snarl (:836-858)": (BUBBLES, [(1, 0, 0), (3, 0, 0), (4, 0, 0), (5, 0, 0), (5, 0, 1), (6, 0, 0)], INF, None),
    "nested bubbles, limit 4 (:989-1001)": (NESTED, [(1, 0, 0), (2, 0, 0), (3, 0, 6), (4, 0, 0), (5, 0, 0)], 4, None),
    "non-simple DAG (:1246-1290)": (NON_SIMPLE, [(1, 0, 0), (2, 0, 0), (3, 0, 0), (3, 0, 1), (4, 0, 0), (5, 0, 0), (6, 0, 0), (7, 0, 1), (8, 0, 0), (8, 0, 2)], INF, None),
}


def positions(graph, seeds):
    lens = graph[0]
    return [(2 * n + r, o) for n, r, o in seeds]


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_candidates_match_the_reference_iterator_expectations(name):
    graph, seeds, limit, expected = CASES[name]
    index, succ = build(*graph)
    got = as_set(oracle_candidates(index, positions(graph, seeds), limit))
    if expected is not None:
        assert got == expected
    if all(r == 0 for _, r, _ in seeds):
        assert got == brute(graph[0], succ, [(n, o) for n, _, o in seeds], limit)
    if name.startswith("bubbles, two children"):
        assert len([t for t in got if t[0] == 1]) == 4                          # 3+0 reaches all four seeds behind it (:846)
        assert {t[1:] for t in got if t[0] == 3} == {(4, 1), (5, 3)}            # 5+0: its own chain, then 6+0; the other child is skipped (:851-857)
    if name.startswith("nested"):
        assert {t for t in got if t[1] == 4} == {(0, 4, 3)}                     # 5+0 goes straight to 1+0 (:994-999)
    index.close()


def random_seeds(g, rng, n):
    seeds = []
    for _ in range(n):
        nid = int(rng.integers(1, len(g.node_seqs) + 1))
        seeds.append((2 * nid + int(rng.integers(0, 2)), int(rng.integers(0, len(g.node_seqs[nid - 1])))))
    return seeds


def test_candidates_feed_the_chain_dp_on_the_oracle():
    """Seeds -> candidates -> find_best_chains on the CPU: exact k-mer anchors sampled along one haplotype chain up completely."""
    g = synth.make_variant_graph(length=4000, n_snp=12, n_ins=2, n_del=2, n_haps=2, seed=4)
    index = g.build_index()
    hap_nodes, hap_off = g.hap_node[0], g.hap_off[0]
    starts = list(range(100, 1300, 40))
    seeds = [(2 * int(hap_nodes[s]), int(hap_off[s])) for s in starts]
    cands = oracle_candidates(index, seeds, 400)
    A = np.zeros(len(starts), capi.chain_anchor_dt)
    for i, s in enumerate(starts):
        A[i] = (s - 100, 15, 0, 0, 15, 0, 15, 15, 0, 0)
    import test_chain_golden as T
    res = T.oracle_chain(A, cands, T.params())
    assert res["chains"][0][1] == list(range(len(starts))) and res["chains"][0][0] == 15 * len(starts)
    index.close()


# ---- GPU --------------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_cuda_candidates_match_reference_cases_and_oracle():
    for name, (graph, seeds, limit, expected) in CASES.items():
        index, _ = build(*graph)
        dev = capi.Device(index, 0)
        pos = positions(graph, seeds)
        got = dev.chain_candidates_batch([pos, pos[:1], []], limit)
        want = oracle_candidates(index, pos, limit)
        assert got[0].tobytes() == want.tobytes(), name
        assert expected is None or as_set(got[0]) == expected, name
        assert len(got[1]) == 0 and len(got[2]) == 0
        dev.close(); index.close()


@pytest.mark.gpu
@pytest.mark.parametrize("limit", [50, 600, INF])
def test_cuda_candidates_parity_on_nested_graphs(limit):
    import test_distance_model as DM
    rng = np.random.default_rng(11)
    for seed in (1, 2, 3):
        g = synth.make_nested_graph(seed=seed) if hasattr(synth, "make_nested_graph") else synth.make_variant_graph(length=3000, n_snp=25, n_ins=4, n_del=4, n_haps=4, seed=seed)
        index = g.build_index() if g.dist is None else capi.HostIndex(g.node_seqs, g.paths, None, k=11, w=5)
        dev = capi.Device(index, 0)
        problems = [random_seeds(g, rng, int(rng.integers(1, 120))) for _ in range(40)] + [random_seeds(g, rng, 700)]
        got = dev.chain_candidates_batch(problems, limit)
        for p, c in zip(problems, got):
            assert c.tobytes() == oracle_candidates(index, p, limit).tobytes()
        with pytest.raises(capi.GbError):
            dev.chain_candidates_batch([[(2, 5000)]], limit)                         # offset outside the node
        dev.close(); index.close()


@pytest.mark.gpu
def test_cuda_seeds_to_chains_end_to_end():
    """candidates and chaining both on the device, against both on the CPU."""
    import test_chain_golden as T
    g = synth.make_variant_graph(length=6000, n_snp=20, n_ins=3, n_del=3, n_haps=3, seed=8)
    index = g.build_index()
    dev = capi.Device(index, 0)
    rng = np.random.default_rng(3)
    problems, anchors = [], []
    for _ in range(30):
        h = int(rng.integers(0, len(g.hap_node)))
        starts = sorted(set(int(x) for x in rng.integers(0, 1500, size=int(rng.integers(2, 60)))))
        base = int(rng.integers(0, len(g.hap_node[h]) - 1600))
        jitter = rng.integers(-3, 4, size=len(starts))
        seeds = [(2 * int(g.hap_node[h][base + s]), int(g.hap_off[h][base + s])) for s in starts]
        A = np.zeros(len(starts), capi.chain_anchor_dt)
        for i, s in enumerate(starts):
            A[i] = (max(0, s + int(jitter[i])), 12, 0, 0, int(rng.integers(5, 13)), 0, 12, 12, 0, 0)
        order = np.argsort(A["read_start"], kind="stable")
        problems.append([seeds[i] for i in order]); anchors.append(A[order])
    cands = dev.chain_candidates_batch(problems, 300)
    p = T.params(max_chains=2)
    got = dev.chain_batch(list(zip(anchors, cands)), p)
    for A, seeds, c, gres in zip(anchors, problems, cands, got):
        want_c = oracle_candidates(index, seeds, 300)
        assert c.tobytes() == want_c.tobytes()
        assert gres == T.oracle_chain(A, want_c, p)
    dev.close(); index.close()


@pytest.mark.parametrize("graph", [ONE, TWO, BUBBLES, NESTED, NON_SIMPLE, TWO_CHAINS], ids=["one", "two", "bubbles", "nested", "non-simple", "two-chains"])
def test_every_position_pair_of_the_reference_graphs_equals_exhaustive_shortest_paths(graph):
    """All positions of every node as seeds.  NESTED is the graph that showed the derived model needs a SIGNED x_out: its
    site {2, 3, 4} sits three bases from the chain start and is bypassed by the edge 1 -> 5, so the way out of the site is
    longer than the exit's own chain coordinate."""
    index, succ = build(*graph)
    seeds = [(n, o) for n, l in enumerate(graph[0], 1) for o in range(l)]
    got = as_set(oracle_candidates(index, [(2 * n, o) for n, o in seeds], INF))
    assert got == brute(graph[0], succ, seeds, INF)
    # and the reverse strand mirrors it: the same walks read backwards
    lens = graph[0]
    rev = [(2 * n + 1, lens[n - 1] - 1 - o) for n, o in seeds]
    got_rev = as_set(oracle_candidates(index, rev, INF))
    assert got_rev == {(j, i, d) for i, j, d in got}
    index.close()


# ---- unittest/minimizer_mapper.cpp:882-1048: "can make correct anchors from minimizers and their zip codes" --------------------------
# A 10 bp read along a 10 bp node, four minimizer hits: 3 bp at the read start (anchored at its first or, as a reverse-strand
# minimizer, its last base), 3 bp at the read end (likewise), a 3 bp one overlapping the first, a 2 bp one abutting the first.
# Expected: read_start == forward_offset, length == minimizer length, and exactly the transitions (0,1) (2,1) (3,1) (0,3), all
# with indel 0 — for both graph strands and all four strand combinations of the outer minimizers.

def anchor_case(graph_reverse_strand, a_rev, b_rev):
    hits = [((2 if a_rev else 0), a_rev, 3), ((9 if b_rev else 7), b_rev, 3), (1, False, 3), (3, False, 2)]      # (offset = pin, is_reverse, length)
    seeds = [(2 * 1 + int(graph_reverse_strand), off) for off, _, _ in hits]          # the graph position equals the read position: same strand walk
    forward_offset = [off - (ln - 1) if rev else off for off, rev, ln in hits]        # Minimizer::forward_offset, minimizer_mapper.hpp:583-592
    return hits, seeds, forward_offset


def oracle_to_anchor(node_len, seed_offset, min_offset, rev, length):
    lib = H.oracle_lib()
    lib.oracle_to_anchor.restype = None
    lib.oracle_to_anchor.argtypes = [C.POINTER(capi.Scores), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p]
    out = np.zeros(1, capi.chain_anchor_dt)
    lib.oracle_to_anchor(C.byref(capi.DEFAULT_SCORES), node_len, seed_offset, min_offset, int(rev), length, 0, capi.ptr(out))
    return out[0]


EXPECTED_TRANSITIONS = {(0, 1): 0, (2, 1): 0, (3, 1): 0, (0, 3): 0}


@pytest.mark.parametrize("graph_reverse_strand", [False, True])
@pytest.mark.parametrize("a_rev", [False, True])
@pytest.mark.parametrize("b_rev", [False, True])
def test_anchors_and_transitions_match_the_reference_minimizer_mapper_test(graph_reverse_strand, a_rev, b_rev):
    import test_chain_golden as T
    index = capi.HostIndex(["AAAAAAAAAA"], [[2]], k=5, w=3)
    hits, seeds, fwd = anchor_case(graph_reverse_strand, a_rev, b_rev)
    anchors = capi.chain_anchors(index, seeds, [h[0] for h in hits], [h[1] for h in hits], [h[2] for h in hits])
    for i, (off, rev, ln) in enumerate(hits):
        assert int(anchors[i]["read_start"]) == fwd[i] and int(anchors[i]["length"]) == ln          # :1003-1006
        assert anchors[i].tobytes() == oracle_to_anchor(10, seeds[i][1], off, rev, ln).tobytes()
        assert int(anchors[i]["score"]) == ln and int(anchors[i]["start_hint_offset"]) == (ln - 1 if rev else 0)
    cands = oracle_candidates(index, seeds, INF)
    got = T.oracle_chain(anchors, cands, T.params(max_indel_bases=65535), transitions=True)
    assert got == EXPECTED_TRANSITIONS                                                               # :1024-1045
    index.close()


def test_anchors_are_cut_at_node_ends():
    """to_anchor keeps the part of the match on the seed's node (:3997, :4015); the margins record the rest."""
    index = capi.HostIndex(["ACGTACGT", "TTGCA"], [[2, 4]], k=5, w=3)
    a = capi.chain_anchors(index, [(2, 6), (4, 1)], [20, 40], [0, 1], [5, 5])
    assert (int(a[0]["length"]), int(a[0]["margin_before"]), int(a[0]["margin_after"]), int(a[0]["read_start"])) == (2, 0, 3, 20)
    assert (int(a[1]["length"]), int(a[1]["margin_before"]), int(a[1]["margin_after"]), int(a[1]["read_start"])) == (2, 3, 0, 39)
    assert int(a[0]["score"]) == 5 and int(a[1]["end_hint_offset"]) == 1 and int(a[1]["base_seed_length"]) == 5
    with pytest.raises(capi.GbError):
        capi.chain_anchors(index, [(2, 8)], [0], [0], [5])
    index.close()


@pytest.mark.gpu
def test_cuda_anchors_candidates_and_transitions_match_the_reference_minimizer_mapper_test():
    import test_chain_golden as T
    index = capi.HostIndex(["AAAAAAAAAA"], [[2]], k=5, w=3)
    dev = capi.Device(index, 0)
    for graph_reverse_strand in (False, True):
        for a_rev in (False, True):
            for b_rev in (False, True):
                hits, seeds, fwd = anchor_case(graph_reverse_strand, a_rev, b_rev)
                anchors = capi.chain_anchors(index, seeds, [h[0] for h in hits], [h[1] for h in hits], [h[2] for h in hits])
                cands = dev.chain_candidates_batch([seeds])[0]
                order = np.argsort(anchors["read_start"], kind="stable")                 # gb_chain_batch takes anchors in read order
                rank = np.empty(len(order), np.uint32); rank[order] = np.arange(len(order))
                c = cands.copy(); c["from"] = rank[cands["from"]]; c["to"] = rank[cands["to"]]
                res = dev.chain_batch([(anchors[order], c)], T.params(max_indel_bases=65535), transitions=True)[0]
                back = {(int(order[f]), int(order[t])): i for (f, t), i in res["transitions"].items()}
                assert back == EXPECTED_TRANSITIONS
                assert res["dp"] == T.oracle_chain(anchors[order], c, T.params(max_indel_bases=65535))["dp"]
    dev.close(); index.close()
