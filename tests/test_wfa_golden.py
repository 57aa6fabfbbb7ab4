"""Pin the oracle's WFAExtender restatement (oracle/wfa.cpp) against the reference's own unit vectors
(src/unittest/gbwt_extender.cpp:1531-2650 -> tests/golden/wfa.json, transcribed by
scripts/extract_wfa_vectors.py), and the CUDA seam gb_wfa_batch against the oracle."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi

GOLD = H.load_golden("wfa.json")
SCORES = (1, 4, 6, 1, 5)
MODE = {"connect": 0, "suffix": 1, "prefix": 2}
OPS = "MXID"      # match, mismatch, insertion, deletion


def graph_index(name):
    g = GOLD["graphs"][name]
    ids = sorted(int(k) for k in g["nodes"])
    assert ids == list(range(1, len(ids) + 1))
    return capi.HostIndex([g["nodes"][str(i)] for i in ids], [[2 * i for i in p] for p in g["paths"]], None, k=5, w=3)


_INDEX = {}


def index_of(name):
    if name not in _INDEX:
        _INDEX[name] = graph_index(name)
    return _INDEX[name]


def oriented(pos):
    return (0, 0) if pos is None else (2 * pos[0] + (1 if pos[1] else 0), pos[2])


def oracle_wfa(index, case):
    lib = H.oracle_lib()
    lib.oracle_wfa.restype = C.c_int
    lib.oracle_wfa.argtypes = [C.POINTER(capi.FlatIndex), C.POINTER(capi.Scores), C.c_void_p, C.c_int, C.c_void_p, C.c_uint32,
                               C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] + [C.c_void_p] * 5 + [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    seq = case["sequence"]
    q = np.frombuffer(seq.encode() + b"\0", dtype=np.uint8).copy()
    em = None if case["error_model"] is None else np.asarray(case["error_model"], dtype=np.float64)
    fn, fo = oriented(case["from"]); tn, to = oriented(case["to"])
    ok, score = C.c_int32(), C.c_int32(); noff, soff, length, npath, nedits = (C.c_uint32() for _ in range(5))
    path = np.zeros(4096, dtype=np.uint32); edits = np.zeros(4096, dtype=np.uint32)
    sc = capi.Scores(*SCORES)
    rc = lib.oracle_wfa(C.byref(index.view), C.byref(sc), None if em is None else capi.ptr(em), MODE[case["call"]], capi.ptr(q), len(seq),
                        fn, fo, tn, to, C.byref(ok), C.byref(score), C.byref(noff), C.byref(soff), C.byref(length),
                        capi.ptr(path), 4096, C.byref(npath), capi.ptr(edits), 4096, C.byref(nedits))
    assert rc == 0
    return {"ok": bool(ok.value), "score": score.value, "node_offset": noff.value, "seq_offset": soff.value, "length": length.value,
            "path": [int(x) for x in path[: npath.value]], "edits": [(OPS[int(w) & 3], int(w) >> 2) for w in edits[: nedits.value]]}


def check_alignment(aln, case, index):
    """Python port of check_alignment (src/unittest/gbwt_extender.cpp:1422-1527)."""
    seq = case["sequence"]
    g = GOLD["graphs"][case["graph"]]
    node_seq = lambda v: g["nodes"][str(v >> 1)] if not (v & 1) else H_revcomp(g["nodes"][str(v >> 1)])
    edges = set()
    for p in g["paths"]:
        for a, b in zip(p, p[1:]):
            edges.add((2 * a, 2 * b)); edges.add((2 * b + 1, 2 * a + 1))
    frm, to = case["from"], case["to"]
    check_from, check_to = case["expect"]["check_alignment"]
    assert aln["ok"]
    assert aln["seq_offset"] + aln["length"] <= len(seq)
    assert not check_from or aln["seq_offset"] == 0
    assert not check_to or aln["seq_offset"] + aln["length"] == len(seq)
    assert sum(n for op, n in aln["edits"] if op != "D") == aln["length"]
    path = aln["path"]
    for a, b in zip(path, path[1:]):
        assert (a, b) in edges
    final_offset = aln["node_offset"] + sum(n for op, n in aln["edits"] if op != "I") - sum(len(node_seq(v)) for v in path[:-1])
    if path:
        assert aln["node_offset"] < len(node_seq(path[0]))
        if check_from:
            fh, fo = oriented(frm)
            if path[0] == fh and aln["node_offset"] > 0:
                assert aln["node_offset"] == fo + 1
            else:
                assert fo + 1 == len(node_seq(fh)) and (fh, path[0]) in edges and aln["node_offset"] == 0
        assert final_offset > 0
        if check_to:
            th, to_off = oriented(to)
            if path[-1] == th and final_offset < len(node_seq(path[-1])):
                assert final_offset == to_off
            else:
                assert to_off == 0 and (path[-1], th) in edges and final_offset == len(node_seq(path[-1]))
    for (a, _), (b, _) in zip(aln["edits"], aln["edits"][1:]):
        assert a != b
    # the alignment itself, base by base
    so, no, po = aln["seq_offset"], aln["node_offset"], 0
    masked = "".join(c if c in "ACGT" else "X" for c in seq)
    for op, n in aln["edits"]:
        if op == "I":
            so += n; continue
        end = no + n
        while end > no:
            assert po < len(path)
            ns = node_seq(path[po])
            ln = min(end, len(ns)) - no
            if op == "M":
                assert masked[so:so + ln] == ns[no:no + ln]; so += ln
            elif op == "X":
                assert all(masked[so + i] != ns[no + i] for i in range(ln)); so += ln
            no += ln
            if no >= len(ns):
                no = 0; end -= len(ns); po += 1
    if path:
        assert po == len(path) - 1 or (po == len(path) and no == 0)


def H_revcomp(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def expected_score(terms):
    matches, mismatches, gaps, gap_length, ends = terms
    m, x, go, ge, flb = SCORES
    return matches * m - mismatches * x - gaps * go - (gap_length - gaps) * ge + ends * flb


def verify(case, aln):
    e = case["expect"]
    if e.get("fail"):
        assert not aln["ok"]
        return
    if e.get("unlocalized_insertion"):
        assert aln["ok"] and aln["path"] == [] and aln["edits"] == [("I", len(case["sequence"]))]
        assert aln["seq_offset"] == 0 and aln["length"] == len(case["sequence"])
        assert aln["score"] == -SCORES[2] - (aln["length"] - 1) * SCORES[3]
    if "score_terms" in e:
        assert aln["score"] == expected_score(e["score_terms"]), (aln, e)
    if "score_expr" in e:
        assert aln["score"] == e["score_expr"]
    if "check_alignment" in e:
        check_alignment(aln, case, index_of(case["graph"]))


@pytest.mark.parametrize("case", GOLD["cases"], ids=[f'{c["line"]}:{c["section"]}' for c in GOLD["cases"]])
def test_oracle_wfa_matches_reference_vectors(case):
    verify(case, oracle_wfa(index_of(case["graph"]), case))


def _problem_of(case):
    frm = None if case["from"] is None else oriented(case["from"])
    to = None if case["to"] is None else oriented(case["to"])
    return (MODE[case["call"]], case["sequence"].encode(), frm, to)


@pytest.mark.gpu
def test_cuda_wfa_matches_reference_vectors_and_oracle():
    by_key = {}
    for case in GOLD["cases"]:
        key = (case["graph"], None if case["error_model"] is None else tuple(case["error_model"]))
        by_key.setdefault(key, []).append(case)
    for (graph, em), cases in by_key.items():
        index = index_of(graph)
        dev = capi.Device(index, scores=capi.Scores(*SCORES))
        got = dev.wfa_batch([_problem_of(c) for c in cases], error_model=None if em is None else list(em))
        for c, g in zip(cases, got):
            want = oracle_wfa(index, c)
            assert g == want, (c["line"], c["section"], g, want)
            verify(c, g)
        dev.close()


@pytest.mark.gpu
def test_cuda_wfa_parity_random_problems_on_a_variant_graph():
    from vg_b200 import synth
    g = synth.make_variant_graph(length=20000, n_snp=60, n_ins=8, n_del=8, n_haps=6, seed=9)
    index = g.build_index()
    rng = np.random.default_rng(17)
    problems, cases = [], []
    for _ in range(400):
        h = int(rng.integers(0, len(g.paths)))
        hs = g.hap_seq[h]
        a = int(rng.integers(50, len(hs) - 300)); ln = int(rng.integers(0, 90))
        seq = hs[a + 1:a + 1 + ln].copy()
        for i in range(len(seq)):
            if rng.random() < 0.03:
                seq[i] = synth.BASES[int(rng.integers(0, 4))]
        seq = bytes(seq)
        if rng.random() < 0.2 and len(seq) > 10:
            k = int(rng.integers(2, len(seq) - 2)); seq = seq[:k] + seq[k + int(rng.integers(1, 4)):]
        if rng.random() < 0.2 and len(seq) > 10:
            k = int(rng.integers(2, len(seq) - 2)); seq = seq[:k] + b"ACGT"[: int(rng.integers(1, 4))] + seq[k:]
        frm = (2 * int(g.hap_node[h][a]), int(g.hap_off[h][a]))
        b = a + 1 + ln
        to = (2 * int(g.hap_node[h][b]), int(g.hap_off[h][b]))
        mode = int(rng.integers(0, 3))
        if rng.random() < 0.15:       # other strand
            seq = H_revcomp(seq.decode()).encode()
            nlen = lambda v: len(g.node_seqs[v // 2 - 1])
            frm, to = (to[0] ^ 1, nlen(to[0]) - 1 - to[1]), (frm[0] ^ 1, nlen(frm[0]) - 1 - frm[1])
        problems.append((mode, seq, frm if mode != 2 else None, to if mode != 1 else None))
        cases.append({"call": ["connect", "suffix", "prefix"][mode], "sequence": seq.decode(), "error_model": None,
                      "from": None if mode == 2 else [frm[0] >> 1, bool(frm[0] & 1), frm[1]], "to": None if mode == 1 else [to[0] >> 1, bool(to[0] & 1), to[1]]})
    dev = capi.Device(index, scores=capi.Scores(*SCORES))
    got = dev.wfa_batch(problems)
    n_ok = 0
    for c, gt in zip(cases, got):
        want = oracle_wfa(index, c)
        assert gt == want, (c, gt, want)
        n_ok += want["ok"]
    assert n_ok > 200
    dev.close()
