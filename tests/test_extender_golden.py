"""Pin the oracle against the reference's own GaplessExtender unit vectors
(src/unittest/gbwt_extender.cpp:822-1156, transcribed in tests/golden/gapless_extender.json),
and — on a GPU — the CUDA path against the same vectors through the C-ABI."""
import numpy as np
import pytest

import helpers as H
from vg_b200 import capi

GOLD = H.load_golden("gapless_extender.json")


def _correct_score(e, read_len_unused=None):
    # correct_score(), src/unittest/gbwt_extender.cpp:124-130 with Aligner defaults 1/4/5
    length = e["read_hi"] - e["read_lo"]
    mm = len(e["mismatches"])
    return (length - mm) * 1 - mm * 4 + 5 * e["left_full"] + 5 * e["right_full"]


def _check_case(case, result, index):
    node_len = lambda v: len(index.node_seqs[(v >> 1) - 1])
    read = case["read"]
    bound = case["error_bound"]
    kind = case["kind"]
    correct = case["correct"]
    if kind == "full_length_match" and not correct:
        for e in result:
            if e["left_full"] and e["right_full"]:
                assert len(e["mismatches"]) > bound
        return
    if kind == "full_length_match":
        assert len(result) == 1
    else:
        assert len(result) == len(correct)
    for i, e in enumerate(result):
        assert e["read_hi"] > e["read_lo"]
        if kind.startswith("full_length"):
            assert e["left_full"] and e["right_full"]
            assert len(e["mismatches"]) <= bound
            assert e["score"] == _correct_score(e)
        else:
            if e["left_full"] and e["right_full"]:
                assert len(e["mismatches"]) > bound
            assert e["read_lo"] == case["correct_offsets"][i]
        got = H.extension_to_mappings(e, node_len, read)
        want = [(H.enc(n, r), off, H.parse_edit_string(s)) for n, r, off, s in correct[i]]
        assert got == want, f"{case['name']}: extension {i}"
    if case.get("check_seeds"):
        # contains(): every seed lies on the extension's path at its diagonal
        e = result[0]
        diags = set()
        ro, no = e["read_lo"], e["offset"]
        for h in e["path"]:
            diags.add((h, ro - no))
            ro += min(node_len(h) - no, e["read_hi"] - ro)
            no = 0
        for s in case["seeds"]:
            assert H.to_seed(*s) in diags


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_matches_reference_vectors(case):
    index = H.golden_graph_index(GOLD["graphs"][case["graph"]])
    seeds = [H.to_seed(*s) for s in case["seeds"]]
    result = H.oracle_extend(index, case["read"], seeds, max_mismatches=case["error_bound"],
                             overlap_threshold=case.get("overlap_threshold", 0.8))
    _check_case(case, result, index)


@pytest.mark.parametrize("case", GOLD["seed_normalisation"], ids=lambda c: f"line{c['line']}")
def test_seed_normalisation(case):
    # "Redundant seeds are removed from a cluster", src/unittest/gbwt_extender.cpp:822-864
    seeds = {H.to_seed(*s) for s in case["seeds"]}
    got = sorted((node >> 1, bool(node & 1), (-d if d < 0 else 0), (d if d >= 0 else 0)) for node, d in seeds)
    want = sorted((n, r, o, ro) for n, r, o, ro in case["correct"])
    assert got == want


@pytest.mark.gpu
def test_cuda_matches_reference_vectors():
    for gname, spec in GOLD["graphs"].items():
        index = H.golden_graph_index(spec)
        dev = capi.Device(index)
        cases = [c for c in GOLD["cases"] if c["graph"] == gname]
        for case in cases:
            seeds = [H.to_seed(*s) for s in case["seeds"]]
            out = dev.extend_batch([case["read"]], [(0, seeds)], max_mismatches=case["error_bound"],
                                   overlap_threshold=case.get("overlap_threshold", 0.8), max_ext=16)
            result = H.gpu_extensions(*out, max_ext=16)[0]
            _check_case(case, result, index)
            want = H.oracle_extend(index, case["read"], seeds, max_mismatches=case["error_bound"],
                                   overlap_threshold=case.get("overlap_threshold", 0.8))
            assert result == want, case["name"]
        dev.close()


# ---- GaplessExtension helpers on hand-built extensions (unittest/gbwt_extender.cpp:576-820, the "toy" graph) ---------
def _ext_args(path_ids, offset, lo, hi):
    import ctypes as C
    p = np.array([2 * i for i in path_ids], dtype=np.uint32)
    return p, (capi.ptr(p), len(p), offset, lo, hi)


@pytest.mark.parametrize("path,offset,interval,start,tail", [
    ([1, 4], 0, (0, 4), (1, 0), (4, 3)),        # starts and ends at node boundaries   :583-599
    ([4, 5], 1, (0, 3), (4, 1), (5, 1)),        # starts in the middle                 :601-617
    ([1, 4], 0, (0, 3), (1, 0), (4, 2)),        # ends in the middle                   :619-635
    ([4], 1, (0, 1), (4, 1), (4, 2)),           # starts and ends in the middle        :637-652
])
def test_extension_positions_reference_vectors(path, offset, interval, start, tail):
    import ctypes as C
    index = H.golden_graph_index(GOLD["graphs"]["toy"])
    lib = H.oracle_lib()
    lib.oracle_extension_positions.argtypes = [C.POINTER(capi.FlatIndex), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.oracle_extension_positions.restype = None
    keep, args = _ext_args(path, offset, *interval)
    out = np.zeros(4, dtype=np.uint32)
    lib.oracle_extension_positions(C.byref(index.view), *args, capi.ptr(out))
    assert (int(out[0]) >> 1, int(out[1])) == start and (int(out[2]) >> 1, int(out[3])) == tail and not (out[0] & 1) and not (out[2] & 1)


@pytest.mark.parametrize("a,b,expected", [
    (([1, 4], 0, (0, 4)), ([5, 6, 8, 9], 0, (0, 4)), 0),            # unrelated extensions          :660-685
    (([1, 4], 0, (0, 4)), ([1, 4], 0, (0, 4)), 4),                  # identical extensions          :686-701
    (([5, 6, 7, 9], 0, (0, 4)), ([5, 6, 8, 9], 0, (0, 4)), 3),      # one difference                :702-731
    (([4, 5, 6, 8], 2, (0, 4)), ([5, 6, 8, 9], 0, (1, 5)), 3),      # partial overlap               :732-760
    (([4, 5, 6, 8, 9], 2, (0, 5)), ([5, 6, 8, 9], 0, (0, 4)), 0),   # shifted by one                :761-790
    (([1, 2, 4, 5], 0, (0, 6)), ([1, 4, 5], 0, (1, 6)), 4),         # paths of different lengths    :791-819
])
def test_extension_overlap_reference_vectors(a, b, expected):
    import ctypes as C
    index = H.golden_graph_index(GOLD["graphs"]["toy"])
    lib = H.oracle_lib()
    lib.oracle_extension_overlap.argtypes = [C.POINTER(capi.FlatIndex)] + [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32] * 2
    lib.oracle_extension_overlap.restype = C.c_uint64
    ka, aa = _ext_args(a[0], a[1], *a[2]); kb, ab = _ext_args(b[0], b[1], *b[2])
    assert lib.oracle_extension_overlap(C.byref(index.view), *aa, *ab) == expected
    assert lib.oracle_extension_overlap(C.byref(index.view), *ab, *aa) == expected          # symmetric
