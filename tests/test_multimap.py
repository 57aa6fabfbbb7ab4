"""max_multimaps > 1 (minimizer_mapper.hpp:410-411): MinimizerMapper::map returns the best max_multimaps alignments in score
order (process_until_threshold_a over the alignments, minimizer_mapper.cpp:1087-1130; secondaries flagged :1199-1206),
map_paired the best max_multimaps pairs (:2505-2598, both reads of a later pair secondary :2552-2557).  Only the primary
carries a MAPQ.  Records are rank-major: record j * n_reads + read is mapping j of the read (GB_ALN_ABSENT when there is none)."""
import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth

SEC, ABSENT, MAPPED, PAIRED = capi.GB_ALN_SECONDARY, capi.GB_ALN_ABSENT, capi.GB_ALN_MAPPED, capi.GB_ALN_PAIRED


def repeat_graph():
    # six exact copies of a 600-bp unit: reads inside it have up to six equally good placements
    return synth.make_variant_graph(length=24000, n_snp=40, n_ins=4, n_del=4, n_haps=4, seed=23, repeat_unit=600, repeat_copies=6)


def params(k, paired=False, rescue=0):
    p = H.paired_params() if paired else H.default_map_params()
    p.max_multimaps = k
    if paired:
        p.max_rescue_attempts = rescue
    return p


def check_layout(res, n, k, paired):
    aln, maps, edits, status = res[:4]
    assert len(aln) == n * k and (status == 0).all()
    assert (aln[:n]["flags"] & (SEC | ABSENT) == 0).all()                                        # rank 0: primaries
    for j in range(1, k):
        rank = aln[j * n:(j + 1) * n]
        present = (rank["flags"] & ABSENT) == 0
        assert ((rank["flags"][present] & SEC) != 0).all() and (rank["mapq"][present] == 0).all()
        assert (rank["n_mappings"][~present] == 0).all() and (rank["read_id"] == np.arange(n)).all()
        prev = aln[(j - 1) * n: j * n]
        assert ((prev["flags"][present] & ABSENT) == 0).all()                                    # no gaps in a read's list
        if not paired:
            assert (rank["score"][present] <= prev["score"][present]).all()                      # score order
        else:
            assert ((rank["flags"][present] & PAIRED) != 0).all()
            assert (present[0::2] == present[1::2]).all()                                        # a pair is secondary as a whole


def test_oracle_multimaps_single_end():
    g = repeat_graph(); index = g.build_index()
    rs = synth.simulate_reads(g, 600, length=150, sub_rate=0.01, seed=36)
    one = H.oracle_map(index, rs.reads, rs.quals, params(1), threads=4)
    for k in (2, 4):
        many = H.oracle_map(index, rs.reads, rs.quals, params(k), threads=4)
        check_layout(many, rs.n, k, False)
        # the primary is the mapping max_multimaps = 1 reports (same candidates, same tie shuffle), MAPQ included
        assert not H.compare_alignments((many[0][:rs.n],) + many[1:4], one, rs.n, mapq_tol=0)
        n_sec = int(((many[0][rs.n:2 * rs.n]["flags"] & SEC) != 0).sum())
        assert n_sec > 50                                                                        # the repeat really has several placements
        # a secondary is a different placement of the same read
        for r in np.flatnonzero((many[0][rs.n:2 * rs.n]["flags"] & SEC) != 0)[:50]:
            assert H.decode_alignment(many[0][r], many[1], many[2])[2] != H.decode_alignment(many[0][rs.n + r], many[1], many[2])[2]
    index.close()


def test_oracle_multimaps_paired():
    g = repeat_graph(); index = g.build_index()
    rs = synth.simulate_pairs(g, 300, sub_rate=0.01, seed=35)
    for rescue in (0, 15):
        one = H.oracle_map_paired(index, rs.reads, rs.quals, params(1, True, rescue), threads=4)
        many = H.oracle_map_paired(index, rs.reads, rs.quals, params(3, True, rescue), threads=4)
        check_layout(many, rs.n, 3, True)
        assert not H.compare_alignments((many[0][:rs.n],) + many[1:4], one, rs.n, mapq_tol=0)
        assert int(((many[0][rs.n:2 * rs.n]["flags"] & SEC) != 0).sum()) > 30
    index.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [2, 3, 8])
def test_cuda_multimaps_single_end_parity(k):
    g = repeat_graph(); index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 2000, length=150, sub_rate=0.01, seed=36)
    got = H.gpu_map(dev, rs.reads, rs.quals, params(k))
    want = H.oracle_map(index, rs.reads, rs.quals, params(k), threads=8)
    check_layout(got, rs.n, k, False)
    bad = H.compare_alignments(got, want, rs.n, k=k)
    assert not bad, f"{len(bad)} of {rs.n * k} records differ; first: record {bad[0][0]}\n got={bad[0][1]}\nwant={bad[0][2]}"
    dev.close(); index.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k,rescue", [(2, 0), (3, 15)])
def test_cuda_multimaps_paired_parity(k, rescue):
    g = repeat_graph(); index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_pairs(g, 1200, sub_rate=0.01, seed=35)
    rng = np.random.default_rng(2)
    for i in range(1, rs.n, 10):                         # mates too noisy to seed: found by rescue when it is on
        m = rng.random(rs.length) < 0.12
        rs.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    p = params(k, True, rescue)
    got = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
    want = H.oracle_map_paired(index, rs.reads, rs.quals, p, threads=8)
    check_layout(got, rs.n, k, True)
    bad = H.compare_alignments(got, want, rs.n, k=k)
    assert not bad, f"{len(bad)} of {rs.n * k} records differ; first: record {bad[0][0]}\n got={bad[0][1]}\nwant={bad[0][2]}"
    both = (got[0]["flags"] & (capi.GB_ALN_RESCUED | ABSENT)) == (want[0]["flags"] & (capi.GB_ALN_RESCUED | ABSENT))
    assert both.all()
    dev.close(); index.close()


@pytest.mark.gpu
def test_cuda_multimaps_on_the_config2_family_and_across_chunks(monkeypatch):
    """Unique placements: every secondary rank is absent, the primaries equal the max_multimaps = 1 run; several host
    chunks land in the right rank-major places."""
    monkeypatch.setenv("GIRAFFE_B200_MAP_CHUNK", "512")
    g = synth.make_variant_graph(length=60000, n_snp=100, n_ins=10, n_del=10, n_haps=4, seed=5)
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_pairs(g, 900, sub_rate=0.005, seed=9)
    p1, p3 = params(1, True, 15), params(3, True, 15)
    one = H.gpu_map(dev, rs.reads, rs.quals, p1, paired=True)
    many = H.gpu_map(dev, rs.reads, rs.quals, p3, paired=True)
    want = H.oracle_map_paired(index, rs.reads, rs.quals, p3, threads=8)
    check_layout(many, rs.n, 3, True)
    assert not H.compare_alignments((many[0][:rs.n],) + many[1:4], one, rs.n, mapq_tol=0)
    assert not H.compare_alignments(many, want, rs.n, k=3)
    dev.close(); index.close()


@pytest.mark.gpu
def test_cuda_multimap_bounds_are_checked():
    g = synth.make_tiny_graph(); index = g.build_index(); dev = capi.Device(index)
    rs = synth.simulate_reads(g, 8, length=150, sub_rate=0.0, seed=3)
    for k in (0, capi.GB_MAX_MULTIMAPS + 1):
        with pytest.raises(capi.GbError):
            H.gpu_map(dev, rs.reads, rs.quals, params(k))
    dev.close(); index.close()


def test_emitters_skip_absent_ranks_and_mark_secondaries():
    """GAF / JSON / GAM of a multi-mapping oracle run: one line (message) per record that exists, secondaries carry
    is_secondary (vg.proto field 15; `vg view -aj` prints "is_secondary": true, test/t/07_vg_map.t:54, :102)."""
    import json
    import test_gam as G
    g = repeat_graph(); index = g.build_index()
    rs = synth.simulate_reads(g, 60, length=150, sub_rate=0.01, seed=36)
    res = H.oracle_map(index, rs.reads, rs.quals, params(3), threads=2)
    aln = res[0]
    n_present = int(((aln["flags"] & ABSENT) == 0).sum()); n_sec = int(((aln["flags"] & SEC) != 0).sum())
    assert 60 < n_present < 180 and n_sec == n_present - 60
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    text = capi.emit_text("json", index.view, aln, res[1], res[2], rbuf, qbuf, read_off)
    recs = [json.loads(line) for line in text.splitlines()]
    assert len(recs) == n_present and sum(1 for r in recs if r.get("is_secondary")) == n_sec
    assert all("mapping_quality" not in r for r in recs if r.get("is_secondary"))
    gaf = capi.emit_text("gaf", index.view, aln, res[1], res[2], rbuf, qbuf, read_off)
    assert len(gaf.splitlines()) == n_present
    msgs, tagged = G.read_stream(capi.emit_text("gam", index.view, aln, res[1], res[2], rbuf, qbuf, read_off))
    assert len(msgs) == n_present and tagged == 1
    assert sum(1 for m in msgs if G._one(G._fields(m), 15) == 1) == n_sec                        # Alignment.is_secondary
    index.close()
