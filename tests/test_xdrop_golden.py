"""Pin the oracle's pinned X-drop aligner against the reference's own unit vectors
(src/unittest/xdrop_aligner.cpp, pinned cases; tests/golden/xdrop_pinned.json)."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi

GOLD = H.load_golden("xdrop_pinned.json")


def oracle_xdrop(index, parents, nodes, root_trim, read, scores, max_gap):
    lib = H.oracle_lib()
    lib.oracle_xdrop_pinned.restype = C.c_int
    lib.oracle_xdrop_pinned.argtypes = [C.POINTER(capi.FlatIndex), C.POINTER(capi.Scores), C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p,
                                        C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    par = np.asarray(parents, dtype=np.int32)
    nd = np.asarray(nodes, dtype=np.uint32)
    q = np.frombuffer(read.encode() + b"\0", dtype=np.uint8).copy()
    score = C.c_int32()
    maps = np.zeros(256, dtype=H.mapping_dt)
    edits = np.zeros(1024, dtype=np.uint32)
    nm, ne = C.c_uint32(), C.c_uint32()
    rc = lib.oracle_xdrop_pinned(C.byref(index.view), C.byref(scores), capi.ptr(par), capi.ptr(nd), len(nd), root_trim,
                                 capi.ptr(q), len(read), max_gap, C.byref(score), capi.ptr(maps), 256, C.byref(nm),
                                 capi.ptr(edits), 1024, C.byref(ne))
    assert rc == 0
    path, e = [], 0
    for i in range(nm.value):
        ed = []
        for _ in range(int(maps[i]["n_edits"])):
            w = int(edits[e]); e += 1
            ed.append(["MSID"[w & 3], w >> 4, "ACGT"[(w >> 2) & 3] if (w & 3) == 1 else ""])
        path.append([int(maps[i]["node"]), int(maps[i]["offset"]), ed])
    return score.value, path


def case_index(case):
    seqs = case["nodes"]
    # one haplotype per root-to-leaf walk so every node has a GBWT record
    parents = case["parents"]
    children = {i: [] for i in range(len(seqs))}
    for i, p in enumerate(parents):
        if p >= 0:
            children[p].append(i)
    paths = []
    for leaf in range(len(seqs)):
        if children[leaf]:
            continue
        walk = []
        x = leaf
        while x >= 0:
            walk.append(2 * (x + 1))
            x = parents[x]
        paths.append(walk[::-1])
    return capi.HostIndex(seqs, paths, None, k=5, w=3)


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_xdrop_matches_reference_vectors(case):
    index = case_index(case)
    sc = capi.Scores(*case["scores"])
    nodes = [2 * (i + 1) for i in range(len(case["nodes"]))]
    score, path = oracle_xdrop(index, case["parents"], nodes, 0, case["read"], sc, max(case["max_gap"], 1))
    want = case["score"] if "score" in case else eval(case["score_expr"], {"len": len(case["read"])})
    assert score == want
    if "path" in case:
        assert path == case["path"]
    # every alignment accounts for the whole query
    qlen = sum(ed[1] for m in path for ed in m[2] if ed[0] in "MSI")
    assert qlen == len(case["read"])
