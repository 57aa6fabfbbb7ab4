"""Stage-level parity of the seeding kernels (SURVEY §8 rows a1-a7, a19's fragment clusters): everything the seeding stage
hands to the extension kernel — minimizers (hash, score bits, offsets, agglomeration window, hits) in score order after the
tie shuffle, seeds, read clusters (members, score, coverage), fragment clusters, which clusters are kept and in which order,
and the (node, diagonal) seeds of every work item — equals what the oracle computes at the same points of
MinimizerMapper::map / map_paired (minimizer_mapper.cpp:3918-4517, :4738-4850, :655-832, :1568-1883,
snarl_seed_clusterer.cpp:28-145), field by field and bit for bit.  End-to-end record equality cannot see a divergence here
that leaves the winning alignment unchanged; these tests can."""
import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth


def _oracle_dump_is_consistent(index, rs, dump, paired):
    """CPU-side sanity of the oracle's dump itself (runs without a GPU)."""
    reads, mins, seeds, clusters, items, item_seeds = dump
    for r in range(rs.n):
        a = reads[r]
        m = mins[int(a["min_off"]): int(a["min_off"]) + int(a["min_cnt"])]
        s = seeds[int(a["seed_off"]): int(a["seed_off"]) + int(a["seed_cnt"])]
        c = clusters[int(a["cluster_off"]): int(a["cluster_off"]) + int(a["cluster_cnt"])]
        it = items[int(a["item_off"]): int(a["item_off"]) + int(a["item_cnt"])]
        assert (np.diff(m["score"]) <= 0).all()                                   # score order
        assert (s["source"] < max(1, len(m))).all() and (np.diff(s["source"].astype(np.int64)) >= 0).all()
        assert int(c["n_seeds"].sum()) == len(s) and (np.diff(c["first_seed"].astype(np.int64)) > 0).all()
        kept = sorted(int(x) for x in c["kept_rank"] if x != 0xffffffff)
        assert kept == list(range(len(it)))
        for t in range(len(it)):
            assert int(c[int(it[t]["cluster"])]["kept_rank"]) == t and int(it[t]["seed_cnt"]) == int(c[int(it[t]["cluster"])]["n_seeds"])


def test_oracle_stage_dump_is_self_consistent():
    g = synth.make_variant_graph(length=30000, n_snp=48, n_ins=6, n_del=6, n_haps=4, seed=3)
    index = g.build_index()
    rs = synth.simulate_reads(g, 60, length=150, sub_rate=0.01, seed=5)
    _oracle_dump_is_consistent(index, rs, H.oracle_seed_stage(index, rs.reads, rs.quals), False)
    rp = synth.simulate_pairs(g, 40, sub_rate=0.01, seed=6)
    _oracle_dump_is_consistent(index, rp, H.oracle_seed_stage(index, rp.reads, rp.quals, H.paired_params(), paired=True), True)
    index.close()


def _graphs():
    yield "variants", synth.make_variant_graph(length=60000, n_snp=120, n_ins=12, n_del=12, n_haps=6, seed=21), dict(length=150, sub_rate=0.01)
    yield "repeats", synth.make_variant_graph(length=24000, n_snp=40, n_ins=4, n_del=4, n_haps=4, seed=23, repeat_unit=600, repeat_copies=6), dict(length=150, sub_rate=0.02)
    yield "branchy", synth.make_branchy_graph(n_layers=2500, n_haps=8, seed=44), dict(length=150, sub_rate=0.005)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["variants", "repeats", "branchy"])
def test_seed_stage_single_end_equals_the_oracle(which):
    """a1 find_minimizers, a2 sort_minimizers_by_score + LazyRNG, a3 find_seeds, a4 cluster_seeds, a5 score_cluster,
    a6 process_until_threshold (cluster selection), a7 extend_seed_group packing — single-end."""
    name, g, kw = next(x for x in _graphs() if x[0] == which)
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 1500, seed=31, **kw)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    got = dev.seed_stage(rbuf, qbuf, read_off)
    want = H.oracle_seed_stage(index, rs.reads, rs.quals)
    bad = H.compare_stage_dumps(got, want, rs.n)
    assert not bad, f"{len(bad)} of {rs.n} reads differ at a stage; first: {bad[0]}"
    assert int(want[0]["item_cnt"].sum()) > rs.n // 2 and int(want[0]["cluster_cnt"].max()) >= 1
    dev.close(); index.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["variants", "repeats"])
def test_seed_stage_paired_equals_the_oracle(which):
    """The same rows through map_paired: joint clustering with fragment clusters (a19), per-read selection under
    found_paired_cluster / has_pair, the pair's shared LazyRNG."""
    name, g, kw = next(x for x in _graphs() if x[0] == which)
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_pairs(g, 1000, sub_rate=kw["sub_rate"], seed=33)
    # a few mates from elsewhere: pairs whose ends fall into different fragment clusters
    rng = np.random.default_rng(2)
    other = synth.simulate_pairs(g, 1000, sub_rate=kw["sub_rate"], seed=34)
    for i in range(1, rs.n, 14):
        rs.reads[i] = other.reads[i]
    p = H.paired_params()
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    got = dev.seed_stage(rbuf, qbuf, read_off, p, paired=True)
    want = H.oracle_seed_stage(index, rs.reads, rs.quals, p, paired=True)
    bad = H.compare_stage_dumps(got, want, rs.n)
    assert not bad, f"{len(bad)} of {rs.n} reads differ at a stage; first: {bad[0]}"
    assert int(want[3]["fragment"].max()) >= 1                      # some pair really had two fragment clusters
    dev.close(); index.close()


@pytest.mark.gpu
def test_pool_overflow_reruns_the_chunk_and_results_do_not_change(monkeypatch):
    """Intermediate pools far too small for the batch (GIRAFFE_B200_POOL_SCALE): the host entry points notice, grow the
    pools and redo the chunk; records equal those of a run with roomy pools and no read reports GB_ITEM_OUT_FULL."""
    g = synth.make_variant_graph(length=40000, n_snp=60, n_ins=6, n_del=6, n_haps=4, seed=9)
    index = g.build_index()
    rs = synth.simulate_pairs(g, 3000, sub_rate=0.01, seed=12)
    p = H.paired_params()
    monkeypatch.setenv("GIRAFFE_B200_MAP_CHUNK", "2048")
    dev = capi.Device(index)
    want = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
    dev.close()
    monkeypatch.setenv("GIRAFFE_B200_POOL_SCALE", "0.02")
    dev = capi.Device(index)
    got = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
    assert (got[3] == 0).all()
    assert not H.compare_alignments(got, want, rs.n, mapq_tol=0)
    se = H.gpu_map(dev, rs.reads, rs.quals)
    assert (se[3] == 0).all()
    dev.close(); index.close()
