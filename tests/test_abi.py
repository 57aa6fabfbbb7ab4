"""The C-ABI shared library loads without a GPU, exports every entry point declared in
include/giraffe_b200.h, and refuses compute loudly when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "giraffe_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = capi.load_library()
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/giraffe_b200.h but not exported"


def test_struct_sizes_match_the_header():
    assert capi.node_rec_dt.itemsize == 16 and capi.hit_dt.itemsize == 24 and capi.extension_dt.itemsize == 64
    assert capi.alignment_dt.itemsize == 32 and capi.mapping_dt.itemsize == 8
    assert C.sizeof(capi.MapParams) == C.sizeof(H.MapParams)
    p = capi.default_map_params()
    q = H.MapParams()
    H.oracle_lib().oracle_map_params_default.argtypes = [C.POINTER(H.MapParams)]
    H.oracle_lib().oracle_map_params_default(C.byref(q))
    assert bytes(p) == bytes(q), "library and oracle disagree on the MinimizerMapper defaults"
    assert (p.hit_cap, p.hard_hit_cap, p.max_extension_mismatches, p.max_alignments, p.distance_limit) == (10, 500, 4, 8, 200)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    g = synth.make_tiny_graph()
    index = g.build_index()
    with pytest.raises(capi.GbError) as e:
        capi.Device(index)
    assert e.value.code == capi.GB_ERR_NO_DEVICE


def test_index_builder_arrays_are_consistent():
    g = synth.make_variant_graph(length=5000, n_snp=20, n_ins=4, n_del=4, n_haps=4, seed=5)
    index = g.build_index()
    nodes = index.array("nodes")
    assert nodes["len"][2::2].tolist() == [len(s) for s in g.node_seqs]
    seq = index.array("seq")
    for nid in (1, 2, len(g.node_seqs)):
        v = 2 * nid
        fwd = bytes(seq[nodes["seq_off"][v]: nodes["seq_off"][v] + nodes["len"][v]])
        rev = bytes(seq[nodes["seq_off"][v + 1]: nodes["seq_off"][v + 1] + nodes["len"][v + 1]])
        assert fwd.decode() == g.node_seqs[nid - 1]
        assert rev == bytes(synth.revcomp_bytes(np.frombuffer(fwd, dtype=np.uint8)))
    # every haplotype visit is in exactly one record, both orientations
    assert int(nodes["size"].sum()) == 2 * sum(len(p) for p in g.paths)
    table = index.array("table")
    hits = index.array("hits")
    used = table[table["key"] != np.uint64(0xFFFFFFFFFFFFFFFF)]
    assert int(used["hit_cnt"].sum()) == len(hits)


def test_distance_payload_matches_bruteforce_shortest_paths():
    """The 16-byte payload reproduces minimum graph distances (Dijkstra over the graph edges)."""
    import heapq
    g = synth.make_variant_graph(length=2500, n_snp=10, n_ins=4, n_del=4, n_haps=6, seed=17)
    lens = [0] + [len(s) for s in g.node_seqs]
    # edges of the variation graph itself (every allele, as the distance index sees it; the
    # haplotypes may not use all of them)
    succ = {}
    frontier = set()
    for alleles in g.slots:
        nxt = set(a for a in alleles if a)
        for v in nxt:
            for u in frontier:
                succ.setdefault(u, set()).add(v)
        if 0 in alleles:
            nxt |= frontier
        frontier = nxt
    dist = g.dist

    def brute(u):
        # distance from the END of u to the START of every reachable node
        best = {}
        pq = [(0, v) for v in succ.get(u, ())]
        heapq.heapify(pq)
        while pq:
            d, v = heapq.heappop(pq)
            if v in best:
                continue
            best[v] = d
            for w in succ.get(v, ()):
                if w not in best:
                    heapq.heappush(pq, (d + lens[v], w))
        return best

    rng = np.random.default_rng(2)
    for u in rng.integers(1, len(g.node_seqs) + 1, size=60):
        u = int(u)
        best = brute(u)
        for v, d in best.items():
            if dist[v]["slot"] > dist[u]["slot"]:
                assert int(dist[v]["x_in"]) - int(dist[u]["x_out"]) == d, (u, v)


def test_flat_index_file_roundtrip(tmp_path):
    """gb_index_save / gb_index_load: the arrays come back bit-identical, a loaded index maps like the built one,
    and files that are not flat indexes are refused."""
    import helpers as H
    from vg_b200 import synth
    g = synth.make_variant_graph(length=20000, n_snp=30, n_ins=4, n_del=4, n_haps=4, seed=9)
    built = g.build_index()
    path = tmp_path / "graph.gbflat"
    built.save(path)
    loaded = capi.HostIndex.load(path)
    assert (loaded.view.n_nodes, loaded.view.k, loaded.view.w, loaded.view.n_paths) == (built.view.n_nodes, built.view.k, built.view.w, built.view.n_paths)
    for name in ("nodes", "seq", "gbwt", "dist", "table", "hits"):
        a, b = built.array(name), loaded.array(name)
        assert a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes(), name
    rs = synth.simulate_reads(g, 200, length=150, sub_rate=0.01, seed=3)
    want = H.oracle_map(built, rs.reads, rs.quals, threads=4)
    got = H.oracle_map(loaded, rs.reads, rs.quals, threads=4)
    assert not H.compare_alignments(got, want, rs.n, mapq_tol=0)
    # refused: a truncated file, a foreign file, a missing file
    raw = path.read_bytes()
    (tmp_path / "short.gbflat").write_bytes(raw[: len(raw) // 2])
    (tmp_path / "foreign.gbflat").write_bytes(b"GBZ" + raw[3:])
    for bad in ("short.gbflat", "foreign.gbflat", "missing.gbflat"):
        with pytest.raises(capi.GbError):
            capi.HostIndex.load(tmp_path / bad)
    # refused: an offset pointing outside the hit array
    corrupt = bytearray(raw)
    header = np.frombuffer(raw[:64], dtype=np.uint8)
    n_hits = int(np.frombuffer(raw[48:56], dtype=np.uint64)[0])
    assert n_hits == built.view.n_hits
    corrupt[48:56] = np.uint64(max(1, n_hits // 2)).tobytes()          # claim fewer hits than the table refers to
    (tmp_path / "corrupt.gbflat").write_bytes(bytes(corrupt))
    with pytest.raises(capi.GbError):
        capi.HostIndex.load(tmp_path / "corrupt.gbflat")
    # refused, not crashed: headers whose counts would wrap the size sum or ask for absurd allocations (each count is
    # bounded by the file size on its own), a zero window, node records reaching into the sequence padding
    for field, value in (((40, 48), 1 << 60), ((24, 32), (1 << 64) - 16), ((32, 40), 1 << 62), ((48, 56), (1 << 64) // 24)):
        crafted = bytearray(raw); crafted[field[0]:field[1]] = np.uint64(value).tobytes()
        (tmp_path / "crafted.gbflat").write_bytes(bytes(crafted))
        with pytest.raises(capi.GbError):
            capi.HostIndex.load(tmp_path / "crafted.gbflat")
    (tmp_path / "tiny.gbflat").write_bytes(raw[:64])                       # header only, table_cells = 2^60
    tiny = bytearray(raw[:64]); tiny[40:48] = np.uint64(1 << 60).tobytes()
    (tmp_path / "tiny.gbflat").write_bytes(bytes(tiny))
    with pytest.raises(capi.GbError):
        capi.HostIndex.load(tmp_path / "tiny.gbflat")
    zero_w = bytearray(raw); zero_w[16:20] = np.uint32(0).tobytes()
    (tmp_path / "w0.gbflat").write_bytes(bytes(zero_w))
    with pytest.raises(capi.GbError):
        capi.HostIndex.load(tmp_path / "w0.gbflat")
    loaded.close(); built.close()
