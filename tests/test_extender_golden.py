"""Pin the oracle against the reference's own GaplessExtender unit vectors
(src/unittest/gbwt_extender.cpp:822-1156, transcribed in tests/golden/gapless_extender.json),
and — on a GPU — the CUDA path against the same vectors through the C-ABI."""
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
