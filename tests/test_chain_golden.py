"""Chaining-route seam: algorithms::find_best_chains (chain_items.cpp:733-800) through gb_chain_batch.

The oracle (oracle/chain.cpp) restates add_transition_if_legal, chain_items_dp and chain_items_traceback in the
reference's order of operations; it is pinned by the four find_best_chain cases of src/unittest/chain_items.cpp:96-155.
Those cases run the zip-code tree over seeds on linear graphs: there the tree lists, for every destination seed, every
seed to its left with the distance between the two seed positions (zip_code_tree.cpp find_distances), which
`linear_candidates` reproduces.  The CUDA kernel evaluates the same recurrence as a lexicographic maximum per
destination in parallel; the CPU tests check the property it relies on (the result does not depend on the order of the
candidates), the GPU tests compare every DP cell and every chain with the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth

A, CD = capi.chain_anchor_dt, capi.chain_candidate_dt


def params(**kw):
    p = capi.ChainParams()
    p.item_bonus, p.recombination_penalty, p.consistency_bonus, p.max_chains = 0, 0, 0, 1
    p.gap_scale, p.max_indel_bases, p.max_read_lookback_bases = 1.0, 100, 2 ** 64 - 1          # chain_items.hpp:407-418, :583-584
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def oracle_chain(anchors, cands, p, transitions=False):
    lib = H.oracle_lib()
    lib.oracle_chain.restype = C.c_int
    lib.oracle_chain.argtypes = [C.POINTER(capi.ChainParams), C.c_uint32, C.c_void_p, C.c_uint64] + [C.c_void_p] * 11
    anchors = np.ascontiguousarray(anchors, dtype=A); cands = np.ascontiguousarray(cands, dtype=CD)
    n = len(anchors); k = int(p.max_chains)
    dps = np.zeros(n + 1, np.int32); src = np.zeros(n + 1, np.uint32); pth = np.zeros(n + 1, np.uint64); rec = np.zeros(n + 1, np.uint32)
    nch = np.zeros(1, np.uint32); cs = np.zeros(k + n + 1, np.int32); cb = np.zeros(k + n + 1, np.uint32); cc = np.zeros(k + n + 1, np.uint32)
    items = np.zeros(n + 1, np.uint32)
    a = np.concatenate([anchors, np.zeros(1, A)]); c = np.concatenate([cands, np.zeros(1, CD)])
    indel = np.zeros(len(cands) + 1, np.uint32)
    rc = lib.oracle_chain(C.byref(p), n, capi.ptr(a), len(cands), capi.ptr(c), capi.ptr(dps), capi.ptr(src), capi.ptr(pth), capi.ptr(rec),
                          capi.ptr(nch), capi.ptr(cs), capi.ptr(cb), capi.ptr(cc), capi.ptr(items), capi.ptr(indel))
    assert rc == 0
    if transitions:
        return {(int(x["from"]), int(x["to"])): int(i) for x, i in zip(cands, indel[:len(cands)]) if i != 0xffffffff}
    return {"dp": [(int(dps[i]), int(src[i]), int(pth[i]), int(rec[i])) for i in range(n)],
            "chains": [(int(cs[c]), [int(x) for x in items[int(cb[c]): int(cb[c]) + int(cc[c])]]) for c in range(int(nch[0]))]}


def make_anchors(rows, paths=0):
    """unittest/chain_items.cpp make_anchors: Anchor(read_start, pos, length, 0, 0, score, seed i) — the hint is the seed
    itself at the anchor's start (start offset 0, end offset = length), base seed length = length (chain_items.hpp:231-244)."""
    a = np.zeros(len(rows), A)
    for i, (read_start, _coord, length, score) in enumerate(rows):
        a[i] = (read_start, length, 0, 0, score, 0, length, length, paths, paths)
    return a


def linear_candidates(rows):
    """What zip_tree_transition_iterator offers on a linear graph: every seed to the left of the destination seed, with
    the distance between the seed positions (all reads forward)."""
    out = [(i, j, rows[j][1] - rows[i][1]) for j in range(len(rows)) for i in range(len(rows)) if i != j and rows[i][1] <= rows[j][1]]
    return np.array(out, dtype=CD) if out else np.zeros(0, CD)


# (read start, graph coordinate, length, score); coordinate = 32 * (node - 1) + offset on make_long_graph(nodes, 32) etc.
REFERENCE_CASES = {
    "abutting in read and graph": ([(1, 1, 9, 9), (10, 10, 9, 9)], 18, [0, 1]),                 # chain_items.cpp:96-108
    "gap in graph": ([(1, 1, 9, 9), (10, 11, 9, 9)], 18, [0, 1]),                               # :110-123
    "gap in read": ([(1, 1, 9, 9), (11, 10, 9, 9)], 18, [0, 1]),                                # :125-138
    "leaves the main diagonal": ([(10, 0, 10, 10), (41, 30, 10, 10), (61, 50, 10, 10), (100, 90, 10, 10)], None, [0, 1, 2, 3]),   # :140-155
}


@pytest.mark.parametrize("name", list(REFERENCE_CASES))
def test_oracle_matches_reference_find_best_chain_cases(name):
    rows, score, chain = REFERENCE_CASES[name]
    got = oracle_chain(make_anchors(rows), linear_candidates(rows), params())
    assert got["chains"][0][1] == chain
    if score is not None:
        assert got["chains"][0][0] == score


def random_problem(rng, n, paths=False, margins=False):
    """Anchors scattered around a few diagonals of a linear graph, candidates = all left-to-right seed pairs within a
    lookback, shuffled, a few duplicated."""
    rows = []
    for _ in range(n):
        rs = int(rng.integers(0, 2000)); ln = int(rng.integers(5, 40))
        diag = int(rng.choice([0, 0, 0, 300, -150])) + int(rng.integers(-6, 7))
        rows.append((rs, max(0, rs + 1000 + diag), ln, int(rng.integers(1, ln + 1))))
    rows.sort(key=lambda r: (r[0], -r[2]))                                                       # sort_anchor_indexes, chain_items.cpp:98-110
    a = make_anchors(rows)
    if margins:
        a["margin_before"] = np.minimum(a["read_start"], rng.integers(0, 4, n)); a["margin_after"] = rng.integers(0, 4, n)
        a["start_hint_offset"] = rng.integers(0, 3, n); a["end_hint_offset"] = a["length"] - a["start_hint_offset"]
        a["base_seed_length"] = a["margin_before"] + a["length"] + a["margin_after"]
    if paths:
        a["start_paths"] = rng.integers(1, 16, n).astype(np.uint64)
        same = rng.random(n) < 0.8
        a["end_paths"] = np.where(same, a["start_paths"], rng.integers(1, 16, n).astype(np.uint64))
    c = [(i, j, rows[j][1] - rows[i][1]) for j in range(n) for i in range(n) if i != j and 0 <= rows[j][1] - rows[i][1] <= 600]
    c = np.array(c, dtype=CD) if c else np.zeros(0, CD)
    if len(c):
        c = np.concatenate([c, c[rng.integers(0, len(c), 3)]])
        c = c[rng.permutation(len(c))]
    return a, c


def test_oracle_result_does_not_depend_on_candidate_order():
    """The property the kernel's per-destination maximum rests on: chain_items_dp keeps, per destination, the
    lexicographic maximum of (evaluation value, score, source), so any order of the candidates gives the same table."""
    rng = np.random.default_rng(5)
    for trial in range(30):
        a, c = random_problem(rng, int(rng.integers(2, 60)), paths=trial % 2 == 0, margins=trial % 3 == 0)
        p = params(max_chains=4, item_bonus=int(trial % 3), recombination_penalty=int(trial % 5), consistency_bonus=int((trial % 4) * 3), gap_scale=1.0 + 0.5 * (trial % 3))
        base = oracle_chain(a, c, p)
        for _ in range(3):
            assert oracle_chain(a, c[rng.permutation(len(c))], p) == base
        assert sum(len(ch) for _, ch in base["chains"]) <= len(a) and base["chains"][0][0] == max(s for s, *_ in base["dp"])
        for score, chain in base["chains"]:
            assert chain == sorted(chain) and all(base["dp"][y][1] == x for x, y in zip(chain, chain[1:]))


def test_chain_entry_refuses_bad_input_without_a_device():
    lib = capi.load_library()
    p = params()
    assert lib.gb_chain_batch(None, C.byref(p), 0, None, None, None, None, None, None, None, None, None, None, None, None, None) == capi.GB_ERR_ARG


# ---- GPU --------------------------------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def dev():
    g = synth.make_tiny_graph()
    d = capi.Device(g.build_index(), 0)
    yield d
    d.close()


@pytest.mark.gpu
def test_cuda_chain_matches_reference_cases_and_oracle(dev):
    problems = [(make_anchors(rows), linear_candidates(rows)) for rows, _, _ in REFERENCE_CASES.values()]
    problems.append((np.zeros(0, A), np.zeros(0, CD)))                                           # no anchors: the empty chain
    problems.append((make_anchors([(5, 5, 10, 10)]), np.zeros(0, CD)))                           # one anchor, no transitions
    got = dev.chain_batch(problems, params())
    for (rows, score, chain), g, (a, c) in zip(REFERENCE_CASES.values(), got, problems):
        assert g["chains"][0][1] == chain and (score is None or g["chains"][0][0] == score)
        assert g == oracle_chain(a, c, params())
    assert got[-2] == {"dp": [], "chains": []}
    assert got[-1] == {"dp": [(10, 0xffffffff, 0, 0)], "chains": [(10, [0])]}


@pytest.mark.gpu
@pytest.mark.parametrize("scheme", [dict(), dict(max_chains=3, item_bonus=1), dict(max_chains=8, recombination_penalty=4, consistency_bonus=6, gap_scale=1.5),
                                    dict(max_chains=2, max_indel_bases=20, max_read_lookback_bases=150)])
def test_cuda_chain_parity_random_problems(dev, scheme):
    rng = np.random.default_rng(17)
    problems = [random_problem(rng, int(rng.integers(1, 200)), paths=i % 2 == 0, margins=i % 3 == 0) for i in range(120)]
    problems.append(random_problem(rng, 1500))                                                    # one large problem among small ones
    p = params(**scheme)
    got = dev.chain_batch(problems, p)
    for i, ((a, c), g) in enumerate(zip(problems, got)):
        want = oracle_chain(a, c, p)
        assert g["dp"] == want["dp"], f"problem {i}: DP table differs"
        assert g["chains"] == want["chains"], f"problem {i}: chains differ"
    assert dev.launches() > 0


@pytest.mark.gpu
def test_cuda_chain_refuses_unsorted_anchors_and_foreign_candidates(dev):
    a = make_anchors([(20, 20, 5, 5), (10, 10, 5, 5)])
    with pytest.raises(capi.GbError):
        dev.chain_batch([(a, np.zeros(0, CD))], params())
    a = make_anchors([(10, 10, 5, 5), (20, 20, 5, 5)])
    with pytest.raises(capi.GbError):
        dev.chain_batch([(a, np.array([(0, 2, 10)], dtype=CD))], params())
