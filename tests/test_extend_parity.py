"""GaplessExtender::extend parity: CUDA (through the C-ABI) vs the oracle on seeded synthetic
graphs, plus CPU-side sanity of the oracle and of the flat GBWT."""
import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth


def _items_for(g, rs, **kw):
    return synth.truth_seeds(g, rs, **kw)


def _oracle_all(index, rs, items, **kw):
    out = []
    for i, seeds in items:
        out.append(H.oracle_extend(index, bytes(rs.reads[i]), seeds, **kw))
    return out


def test_oracle_clean_reads_are_full_length_exact():
    g = synth.make_tiny_graph()
    index = g.build_index()
    rs = synth.simulate_reads(g, 50, length=150, sub_rate=0.0, seed=3)
    items = _items_for(g, rs)
    for (i, seeds), exts in zip(items, _oracle_all(index, rs, items)):
        assert len(exts) >= 1
        e = exts[0]
        assert e["left_full"] and e["right_full"] and e["mismatches"] == []
        assert e["score"] == 150 + 10


def test_oracle_substitutions_counted():
    g = synth.make_variant_graph(length=20000, n_snp=40, n_ins=5, n_del=5, n_haps=4, seed=7)
    index = g.build_index()
    rs = synth.simulate_reads(g, 200, length=150, sub_rate=0.01, seed=9)
    items = _items_for(g, rs)
    n_full = 0
    for (i, seeds), exts in zip(items, _oracle_all(index, rs, items)):
        for e in exts:
            length = e["read_hi"] - e["read_lo"]
            mm = len(e["mismatches"])
            assert e["score"] == length - 5 * mm + 5 * e["left_full"] + 5 * e["right_full"]
            assert e["mismatches"] == sorted(e["mismatches"])
        if exts and exts[0]["left_full"] and exts[0]["right_full"]:
            n_full += 1
    assert n_full > 150


def test_flat_gbwt_matches_bruteforce_haplotype_scan():
    """follow_paths over the flat GBWT == scanning the haplotype paths for occurrences."""
    import ctypes as C
    g = synth.make_variant_graph(length=3000, n_snp=12, n_ins=3, n_del=3, n_haps=5, seed=13)
    index = g.build_index()
    lib = H.oracle_lib()
    # all oriented sequences (forward and reverse of each path)
    seqs = []
    for p in g.paths:
        seqs.append(list(p))
        seqs.append([v ^ 1 for v in reversed(p)])

    def occurrences(pattern):
        n = 0
        for s in seqs:
            for i in range(len(s) - len(pattern) + 1):
                if s[i:i + len(pattern)] == pattern:
                    n += 1
        return n

    rng = np.random.default_rng(1)
    state = np.zeros(6, dtype=np.int64)
    out = np.zeros(6 * 16, dtype=np.int64)
    for _ in range(60):
        s = seqs[int(rng.integers(0, len(seqs)))]
        start = int(rng.integers(0, len(s) - 6))
        pattern = [s[start]]
        lib.oracle_bd_state(C.byref(index.view), pattern[0], capi.ptr(state))
        assert state[2] - state[1] + 1 == occurrences(pattern)
        for step in range(5):
            backward = bool(rng.integers(0, 2))
            n = lib.oracle_follow_paths(C.byref(index.view), capi.ptr(state), int(backward), capi.ptr(out), 16)
            assert n <= 16
            if n == 0:
                # only legal at a path end: no haplotype continues the pattern
                for sq in seqs:
                    for i in range(len(sq) - len(pattern) + 1):
                        if sq[i:i + len(pattern)] == pattern:
                            assert (i == 0) if backward else (i + len(pattern) == len(sq))
                break
            total = 0
            succ_sizes = {}
            for j in range(n):
                st = out[6 * j: 6 * j + 6]
                size = int(st[2] - st[1] + 1)
                assert size == int(st[5] - st[4] + 1)
                node = int(st[3]) ^ 1 if backward else int(st[0])
                pat = ([node] + pattern) if backward else (pattern + [node])
                assert size == occurrences(pat), (pattern, node, backward)
                succ_sizes[node] = st.copy()
                total += size
            assert total <= occurrences(pattern)
            # continue along a random successor
            node = list(succ_sizes)[int(rng.integers(0, len(succ_sizes)))]
            state[:] = succ_sizes[node]
            pattern = ([node] + pattern) if backward else (pattern + [node])


def _compare(dev, index, g, rs, items, max_ext=16, **kw):
    out = dev.extend_batch([bytes(r) for r in rs.reads], items, max_ext=max_ext, path_cap=512, mism_cap=256, **kw)
    got = H.gpu_extensions(*out, max_ext=max_ext)
    want = _oracle_all(index, rs, items, **kw)
    bad = [i for i in range(len(items)) if got[i] != want[i]]
    assert not bad, f"{len(bad)} of {len(items)} items differ; first: item {bad[0]}\n got={got[bad[0]]}\nwant={want[bad[0]]}"
    return got


@pytest.mark.gpu
def test_cuda_extend_parity_tiny_graph():
    g = synth.make_tiny_graph()
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 1000, length=150, sub_rate=0.01, seed=11)
    _compare(dev, index, g, rs, _items_for(g, rs))
    _compare(dev, index, g, rs, _items_for(g, rs, false_seeds=3), max_ext=24)
    dev.close()


@pytest.mark.gpu
@pytest.mark.parametrize("sub_rate,trim,max_mm", [(0.002, True, 4), (0.03, True, 4), (0.06, False, 4), (0.03, True, 1)])
def test_cuda_extend_parity_variant_graph(sub_rate, trim, max_mm):
    g = synth.make_variant_graph(length=100000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=21)
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 3000, length=150, sub_rate=sub_rate, seed=23)
    items = _items_for(g, rs, read_offsets=(0, 20, 51, 77, 111, 140), false_seeds=2)
    _compare(dev, index, g, rs, items, max_ext=24, trim=trim, max_mismatches=max_mm)
    dev.close()


@pytest.mark.gpu
def test_cuda_extend_parity_branchy_graph():
    g = synth.make_branchy_graph(n_layers=2000, n_haps=16, seed=4)
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 2000, length=150, sub_rate=0.005, seed=44)
    items = _items_for(g, rs, read_offsets=(3, 60, 100, 149))
    _compare(dev, index, g, rs, items, max_ext=24)
    dev.close()


@pytest.mark.gpu
def test_cuda_extend_edge_cases():
    g = synth.make_tiny_graph()
    index = g.build_index()
    dev = capi.Device(index)
    reads = [b"ACGT", b"", b"N" * 40, bytes(g.hap_seq[0][:33]), bytes(g.hap_seq[0][100:101])]
    n0 = int(g.hap_node[0][0]); n100 = int(g.hap_node[0][100]); o100 = int(g.hap_off[0][100])
    items = [(0, []), (1, [(2 * n0, 0)]), (2, [(2 * n0, 0), (2 * n0 + 1, 5)]), (3, [(2 * n0, 0)] * 5),
             (4, [(2 * n100, -o100)])]
    out = dev.extend_batch(reads, items, max_ext=8)
    got = H.gpu_extensions(*out, max_ext=8)
    for (r, seeds), g_ext in zip(items, got):
        want = H.oracle_extend(index, reads[r], seeds) if len(reads[r]) and seeds else []
        assert g_ext == want
    dev.close()
