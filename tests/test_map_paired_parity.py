"""MinimizerMapper::map_paired parity (forced fragment distribution, with and without mate rescue):
gb_map_paired_batch on the GPU vs the oracle restatement, BASELINE.json configs[1] family."""
import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth


def _run(g, rs, params):
    index = g.build_index()
    dev = capi.Device(index)
    got = H.gpu_map(dev, rs.reads, rs.quals, params, paired=True)
    want = H.oracle_map_paired(index, rs.reads, rs.quals, params, threads=8)
    bad = H.compare_alignments(got, want, rs.n)
    dev.close()
    assert not bad, f"{len(bad)} of {rs.n} reads differ; first: read {bad[0][0]}\n got={bad[0][1]}\nwant={bad[0][2]}"
    return got, want


def test_oracle_pairs_map_to_their_origin():
    g = synth.make_variant_graph(length=50000, n_snp=80, n_ins=10, n_del=10, n_haps=4, seed=3)
    index = g.build_index()
    rs = synth.simulate_pairs(g, 200, sub_rate=0.0, seed=7)
    aln, maps, edits, status, counters = H.oracle_map_paired(index, rs.reads, rs.quals, H.paired_params())
    assert (aln["flags"] & capi.GB_EXT_LEFT_FULL).all()          # bit 0 = mapped
    assert (aln["score"] == 160).all()
    assert (aln["mapq"] == 60).mean() > 0.95
    # mates land on opposite strands (mate 2 is reported in input orientation)
    strands = np.array([maps[int(a["mapping_off"])]["node"] & 1 for a in aln])
    assert (strands[0::2] != strands[1::2]).all()


def _wrecked_pairs(n_pairs=300, seed=61):
    g = synth.make_variant_graph(length=100000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=5)
    rs = synth.simulate_pairs(g, n_pairs, sub_rate=0.01, seed=seed)
    rng = np.random.default_rng(2)
    wrecked = np.arange(1, rs.n, 6)                 # mate 2 of every third pair gets 12 % substitutions: few or no seeds
    for i in wrecked:
        m = rng.random(rs.length) < 0.12
        rs.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    return g, rs, wrecked


def test_oracle_fix_end_deletions_reference_vector():
    """unittest/minimizer_mapper.cpp:1130-1181 "can fix up alignments with deletions on the ends": node 1 (+3) [2D],
    node 2 (+0) [2D 1M 1D], node 3 (+0) [1D]  ->  node 2 (+2) [1M]."""
    import ctypes as C
    lib = H.oracle_lib()
    lib.oracle_fix_end_deletions.argtypes = [C.c_uint32] + [C.c_void_p] * 5
    lib.oracle_fix_end_deletions.restype = C.c_uint32
    node = np.array([1, 2, 3], dtype=np.uint32); offset = np.array([3, 0, 0], dtype=np.uint32)
    count = np.array([1, 3, 1], dtype=np.uint32)
    frm = np.array([2, 2, 1, 1, 1], dtype=np.uint32); to = np.array([0, 0, 1, 0, 0], dtype=np.uint32)
    n = lib.oracle_fix_end_deletions(3, capi.ptr(node), capi.ptr(offset), capi.ptr(count), capi.ptr(frm), capi.ptr(to))
    assert n == 1 and node[0] == 2 and offset[0] == 2 and count[0] == 1 and (frm[0], to[0]) == (1, 1)
    # an alignment that is all deletion is cleared
    node = np.array([1], dtype=np.uint32); offset = np.array([0], dtype=np.uint32); count = np.array([1], dtype=np.uint32)
    frm = np.array([4], dtype=np.uint32); to = np.array([0], dtype=np.uint32)
    assert lib.oracle_fix_end_deletions(1, capi.ptr(node), capi.ptr(offset), capi.ptr(count), capi.ptr(frm), capi.ptr(to)) == 0


def test_oracle_score_contiguous_alignment_reference_vector():
    """unittest/aligner.cpp:347-369 "Full-length bonus is applied to both ends by rescoring": the 12-mapping alignment with a
    3-bp deletion split over two mappings and a 3-bp insertion scores 129 without and 139 with a bonus of 5
    (fix_dozeu_score re-scores rescued alignments with this function, minimizer_mapper.cpp:3502-3517)."""
    import ctypes as C
    lib = H.oracle_lib()
    lib.oracle_score_contiguous.argtypes = [C.POINTER(capi.Scores), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_score_contiguous.restype = C.c_int32
    mappings = [[(4, 4)], [(1, 1)], [(3, 3)], [(1, 1)], [(32, 32)], [(32, 32)], [(8, 8)], [(1, 1)], [(24, 24)], [(1, 0)],
                [(2, 0), (3, 3), (0, 3), (27, 27)], [(9, 9)]]
    count = np.array([len(m) for m in mappings], dtype=np.uint32)
    frm = np.array([e[0] for m in mappings for e in m], dtype=np.uint32); to = np.array([e[1] for m in mappings for e in m], dtype=np.uint32)
    sub = np.zeros(len(frm), dtype=np.uint8)
    assert int(to.sum()) == 148                          # the read of the vector
    for bonus, want in ((0, 129), (5, 139)):
        sc = capi.Scores(1, 4, 6, 1, bonus)
        assert lib.oracle_score_contiguous(C.byref(sc), len(mappings), capi.ptr(count), capi.ptr(frm), capi.ptr(to), capi.ptr(sub)) == want


def test_oracle_rescue_recovers_mates_without_seeds():
    """attempt_rescue (minimizer_mapper.cpp:3264-3482) in the oracle: mates too noisy to seed are found next to
    their partner; the rescued records are internally consistent."""
    g, rs, wrecked = _wrecked_pairs()
    index = g.build_index()
    p0 = H.paired_params(); p15 = H.paired_params(); p15.max_rescue_attempts = 15
    a0 = H.oracle_map_paired(index, rs.reads, rs.quals, p0, threads=8)
    a15 = H.oracle_map_paired(index, rs.reads, rs.quals, p15, threads=8)
    m0, m15 = (a0[0]["flags"] & 1), (a15[0]["flags"] & 1)
    assert m0[wrecked].mean() < 0.5 and m15[wrecked].mean() > 0.85
    rescued = np.nonzero(a15[0]["flags"] & capi.GB_ALN_RESCUED)[0]
    assert len(rescued) >= 50 and a15[4]["rescues"] >= len(rescued)
    for i in rescued:
        score, mapq, path = H.decode_alignment(a15[0][i], a15[1], a15[2])
        if not path:
            continue
        assert sum(e[1] for m in path for e in m[2] if e[0] in "MSI") == rs.length
        if score < 50:
            continue          # short chance matches pass the reference's likelihood filter too (:3451-3481)
        # a real rescue: the partner is mapped on the other strand nearby
        mate = i ^ 1
        assert a15[0][mate]["flags"] & 1
        _, _, mpath = H.decode_alignment(a15[0][mate], a15[1], a15[2])
        assert (path[0][0] & 1) != (mpath[0][0] & 1)
        assert abs((path[0][0] >> 1) - (mpath[0][0] >> 1)) < 200          # node ids grow along the chain
    # pairs that needed no rescue come out the same with and without it
    clean = [i for i in range(rs.n) if (i | 1) not in set(wrecked.tolist())]
    same = sum(H.decode_alignment(a0[0][i], a0[1], a0[2]) == H.decode_alignment(a15[0][i], a15[1], a15[2]) for i in clean)
    assert same >= 0.97 * len(clean)


@pytest.mark.gpu
@pytest.mark.parametrize("sub_rate,seed", [(0.002, 22), (0.02, 23), (0.06, 24)])
def test_map_paired_parity_variant_graph(sub_rate, seed):
    g = synth.make_variant_graph(length=200000, n_snp=320, n_ins=40, n_del=40, n_haps=8, seed=2)
    rs = synth.simulate_pairs(g, 2000, sub_rate=sub_rate, seed=seed)
    _run(g, rs, H.paired_params())


@pytest.mark.gpu
def test_map_paired_parity_with_indels_and_odd_fragments():
    g = synth.make_variant_graph(length=100000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=21)
    rs = synth.simulate_pairs(g, 1500, frag_mean=400, frag_sd=150, sub_rate=0.01, seed=9)   # many improper fragments
    # sprinkle indels and a few garbage mates
    rng = np.random.default_rng(3)
    for i in rng.integers(0, rs.n, size=150):
        p = int(rng.integers(10, 140))
        rs.reads[i, p:-1] = rs.reads[i, p + 1:]
    for i in rng.integers(0, rs.n, size=40):
        rs.reads[i] = synth.BASES[rng.integers(0, 4, size=rs.length)]
    _run(g, rs, H.paired_params())


@pytest.mark.gpu
def test_map_paired_parity_branchy_graph():
    g = synth.make_branchy_graph(n_layers=4000, n_haps=16, seed=4)
    rs = synth.simulate_pairs(g, 1000, sub_rate=0.005, seed=44)
    _run(g, rs, H.paired_params())


@pytest.mark.gpu
@pytest.mark.parametrize("attempts", [15, 2])
def test_map_paired_parity_with_mate_rescue(attempts):
    """max_rescue_attempts != 0: attempt_rescue (minimizer_mapper.cpp:3264-3482) and the rescue branch of
    map_paired (:2288-2457, multiplicities :2661-2690) on the GPU vs the oracle."""
    g, rs, wrecked = _wrecked_pairs()
    p = H.paired_params(); p.max_rescue_attempts = attempts
    got, want = _run(g, rs, p)
    assert ((got[0]["flags"] & capi.GB_ALN_RESCUED) == (want[0]["flags"] & capi.GB_ALN_RESCUED)).all()
    assert (got[0]["flags"] & capi.GB_ALN_RESCUED).sum() >= 50


@pytest.mark.gpu
def test_map_paired_rescue_parity_with_indels_odd_fragments_and_garbage_mates():
    """Rescue alignments that need the gapped aligner (indels in the rescued mate), improper fragments
    (rescue window misses) and random mates (rescue finds nothing / chance matches)."""
    g = synth.make_variant_graph(length=100000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=21)
    rs = synth.simulate_pairs(g, 1500, frag_mean=400, frag_sd=150, sub_rate=0.01, seed=9)
    rng = np.random.default_rng(3)
    for i in rng.integers(0, rs.n, size=400):
        p = int(rng.integers(10, 140))
        rs.reads[i, p:-1] = rs.reads[i, p + 1:]
        m = rng.random(rs.length) < 0.08                       # and too noisy to seed
        rs.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    for i in rng.integers(0, rs.n, size=60):
        rs.reads[i] = synth.BASES[rng.integers(0, 4, size=rs.length)]
    p = H.paired_params(); p.max_rescue_attempts = 15
    got, want = _run(g, rs, p)
    assert (got[0]["flags"] & capi.GB_ALN_RESCUED).sum() >= 100


@pytest.mark.gpu
def test_map_paired_rescue_parity_branchy_graph():
    """Short nodes, many alleles: big rescue subgraphs.  Pairs whose subgraph exceeds the per-warp rescue
    workspace report GB_ITEM_OUT_FULL (the caller re-maps those on the CPU); everything else is identical."""
    g = synth.make_branchy_graph(n_layers=4000, n_haps=16, seed=4)
    rs = synth.simulate_pairs(g, 600, sub_rate=0.005, seed=44)
    rng = np.random.default_rng(5)
    for i in range(1, rs.n, 4):
        m = rng.random(rs.length) < 0.12
        rs.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    index = g.build_index()
    dev = capi.Device(index)
    p = H.paired_params(); p.max_rescue_attempts = 15
    got = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
    want = H.oracle_map_paired(index, rs.reads, rs.quals, p, threads=8)
    dev.close()
    assert set(np.unique(got[3]).tolist()) <= {0, capi.GB_ITEM_OUT_FULL}
    ok = np.nonzero(got[3] == 0)[0]
    assert len(ok) >= 0.9 * rs.n
    bad = H.compare_alignments(got, want, rs.n, indices=ok.tolist())
    assert not bad, f"{len(bad)} of {len(ok)} reads differ; first: read {bad[0][0]}\n got={bad[0][1]}\nwant={bad[0][2]}"
    assert (got[0]["flags"][ok] & capi.GB_ALN_RESCUED).sum() >= 20


@pytest.mark.gpu
def test_rescue_seed_limit_beyond_the_workspace_is_refused_loudly():
    g = synth.make_tiny_graph()
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_pairs(g, 4, frag_mean=300, frag_sd=20, sub_rate=0.0, seed=1)
    p = H.paired_params(300, 20)
    p.max_rescue_attempts = 15; p.rescue_seed_limit = 500
    with pytest.raises(capi.GbError):
        H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
    dev.close()


@pytest.mark.gpu
def test_map_paired_rescue_on_clean_pairs_matches_no_rescue_path():
    """Pairs whose mates both cluster go through the thread-per-pair fast path even with rescue enabled."""
    g = synth.make_variant_graph(length=200000, n_snp=320, n_ins=40, n_del=40, n_haps=8, seed=2)
    rs = synth.simulate_pairs(g, 2000, sub_rate=0.02, seed=23)
    p = H.paired_params(); p.max_rescue_attempts = 15
    _run(g, rs, p)


@pytest.mark.gpu
@pytest.mark.parametrize("tables", ["16,1", "8,2", "64,16,8"])
def test_map_paired_parity_through_the_seeding_retry_pass(tables, monkeypatch):
    """First-pass seeding tables too small for most pairs: they are retried at full size by the
    second launch and must come out identical.  A third number caps the seed sets clustered in
    shared memory, sending the others through label propagation over the HBM records."""
    monkeypatch.setenv("GIRAFFE_B200_SEED_TABLES", tables)
    g = synth.make_variant_graph(length=100000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=5)
    rs = synth.simulate_pairs(g, 1200, sub_rate=0.01, seed=51)
    _run(g, rs, H.paired_params())


@pytest.mark.gpu
def test_map_paired_parity_across_host_chunks(monkeypatch):
    """Several double-buffered chunks per call: headers must index the caller's whole pools."""
    monkeypatch.setenv("GIRAFFE_B200_MAP_CHUNK", "256")
    g = synth.make_variant_graph(length=100000, n_snp=160, n_ins=20, n_del=20, n_haps=8, seed=5)
    rs = synth.simulate_pairs(g, 1100, sub_rate=0.02, seed=53)
    _run(g, rs, H.paired_params())


@pytest.mark.gpu
def test_map_paired_edge_cases():
    """Ragged pair lengths, mates without minimizers, N runs, an empty read, identical mates."""
    g = synth.make_variant_graph(length=50000, n_snp=80, n_ins=10, n_del=10, n_haps=4, seed=3)
    index = g.build_index()
    dev = capi.Device(index)
    hs = g.hap_seq[0]
    rc = lambda a: bytes(synth.revcomp_bytes(np.frombuffer(a, dtype=np.uint8)))
    f = lambda a, b: bytes(hs[a:b])
    pairs = [
        (f(1000, 1150), rc(f(1250, 1400))),                     # proper pair
        (f(2000, 2100), rc(f(2300, 2420))),                     # ragged lengths
        (f(3000, 3150), b"ACGT"),                               # mate 2 too short for a minimizer
        (b"N" * 150, rc(f(4200, 4350))),                        # mate 1 all N
        (f(5000, 5150).replace(b"A", b"N", 4), rc(f(5300, 5450)).replace(b"C", b"N", 2)),
        (f(6000, 6150), rc(f(26000, 26150))),                   # far apart: no shared fragment cluster
        (f(7000, 7150), f(7000, 7150)),                         # same strand, same place
        (b"", rc(f(8200, 8350))),                               # empty mate
        (f(9000, 9039), rc(f(9100, 9139))),                     # exactly k + w - 1 bases
    ]
    reads = [r for p in pairs for r in p]
    quals = [bytes([30] * len(r)) for r in reads]
    params = H.paired_params()
    got = H.gpu_map(dev, reads, quals, params, paired=True)
    want = H.oracle_map_paired(index, reads, quals, params)
    bad = H.compare_alignments(got, want, len(reads))
    assert not bad, bad[0]
    got = H.gpu_map(dev, reads, None, params, paired=True)
    want = H.oracle_map_paired(index, reads, None, params)
    bad = H.compare_alignments(got, want, len(reads))
    assert not bad, bad[0]
    dev.close()


@pytest.mark.gpu
def test_unusable_fragment_distribution_falls_back_to_single_end():
    """fragment limit (mean + 2 sd) below the read limit max(200, L + 50): map_paired maps both ends
    single-ended and emits them as a pair (minimizer_mapper.cpp:1469-1496)."""
    g = synth.make_variant_graph(length=60000, n_snp=100, n_ins=10, n_del=10, n_haps=4, seed=8)
    rs = synth.simulate_pairs(g, 300, sub_rate=0.01, seed=5)
    index = g.build_index()
    dev = capi.Device(index)
    p = H.paired_params(100.0, 20.0)                       # limit 140 < 200
    got = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
    want = H.oracle_map_paired(index, rs.reads, rs.quals, p, threads=8)
    assert not H.compare_alignments(got, want, rs.n)
    se = H.gpu_map(dev, rs.reads, rs.quals, H.default_map_params())
    assert not H.compare_alignments(got, se, rs.n, mapq_tol=0)
    assert ((got[0]["flags"] & capi.GB_ALN_PAIRED) != 0).all() and ((se[0]["flags"] & capi.GB_ALN_PAIRED) == 0).all()
    dev.close()


def _repeat_graph():
    return synth.make_variant_graph(length=24000, n_snp=40, n_ins=4, n_del=4, n_haps=4, seed=23, repeat_unit=600, repeat_copies=6)


@pytest.mark.gpu
@pytest.mark.parametrize("rescue", [0, 15])
def test_pairs_in_repeats_draw_the_pair_rng_in_reference_order(rescue):
    """Six exact copies of a 600-bp unit: reads inside it have six tied clusters per mate, six tied fragment clusters and
    six tied extension sets, so every tie shuffle of map_paired fires.  The shuffles share one LazyRNG per pair and the
    reference interleaves them read by read (clusters of read 1, its extension sets and tails, then clusters of read 2,
    minimizer_mapper.cpp:1723-2043); which copy becomes the primary placement depends on that order."""
    g = _repeat_graph()
    rs = synth.simulate_pairs(g, 1500, sub_rate=0.01, seed=35)
    p = H.paired_params(); p.max_rescue_attempts = rescue
    got, want = _run(g, rs, p)
    multi = int((want[0]["mapq"] < 60).sum())
    assert multi > 50                                   # the repeat really produced ambiguous placements


@pytest.mark.gpu
def test_single_end_reads_in_repeats():
    g = _repeat_graph()
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 2000, length=150, sub_rate=0.01, seed=36)
    got = H.gpu_map(dev, rs.reads, rs.quals)
    want = H.oracle_map(index, rs.reads, rs.quals, threads=8)
    bad = H.compare_alignments(got, want, rs.n)
    assert not bad, f"{len(bad)} of {rs.n} reads differ; first: read {bad[0][0]}\n got={bad[0][1]}\nwant={bad[0][2]}"
    dev.close(); index.close()
