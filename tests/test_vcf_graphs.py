"""Graphs built from the reference's own VCF test inputs (tests/golden/vcf/, copied by scripts/extract_vcf_fixtures.py):
  * test/t/50_vg_giraffe.t:48 / :93 — the whole-pipeline pin SURVEY §8(c) names: the 63-bp read reads/small.middle.ref.fq
    mapped to the graph of small/x.fa + x.vcf.gz scores 73 (63 matches + both full-length bonuses of 5), and 63 with
    --full-l-bonus 0;
  * test/1mb1kgp (1000 Genomes sites, first 100 kbp): touching SNPs, multi-allelic sites, indels, alleles chopped into several
    nodes — the graphs the round-1 index model refused.  It loads through the derived chain model and maps GPU == oracle."""
import numpy as np
import pytest

import helpers as H
import vcf_graph as V
from vg_b200 import capi, synth


def _read_fq():
    lines = (V.GOLD / "small.middle.ref.fq").read_text().split("\n")
    return lines[1].strip()


def test_giraffe_t_end_to_end_pin_on_the_oracle():
    g, kept = V.small_x()
    assert kept >= 60                                    # nearly all of x.vcf.gz's 75 sites fit (overlapping records are skipped)
    index = g.build_index()
    read = _read_fq()
    assert len(read) == 63
    reads = np.frombuffer(read.encode(), dtype=np.uint8)[None, :].copy()
    quals = np.full(reads.shape, 30, dtype=np.uint8)
    for bonus, want in ((5, 73), (0, 63)):               # 50_vg_giraffe.t:48 and :93
        res = H.oracle_map(index, reads, quals, scores=capi.Scores(1, 4, 6, 1, bonus))
        score, mapq, path = H.decode_alignment(res[0][0], res[1], res[2])
        assert score == want and path and sum(l for _, _, ed in path for op, l, _ in ed if op == "M") == 63
    index.close()


@pytest.mark.gpu
def test_giraffe_t_end_to_end_pin_on_the_gpu():
    g, _ = V.small_x()
    index = g.build_index()
    reads = np.frombuffer(_read_fq().encode(), dtype=np.uint8)[None, :].copy()
    quals = np.full(reads.shape, 30, dtype=np.uint8)
    for bonus, want in ((5, 73), (0, 63)):
        dev = capi.Device(index, scores=capi.Scores(1, 4, 6, 1, bonus))
        got = H.gpu_map(dev, reads, quals)
        assert int(got[0][0]["score"]) == want and int(got[3][0]) == 0
        assert not H.compare_alignments(got, H.oracle_map(index, reads, quals, scores=capi.Scores(1, 4, 6, 1, bonus)), 1)
        dev.close()
    index.close()


def test_1000_genomes_slice_loads_through_the_chain_model():
    g, kept = V.kgp_100k()
    assert kept > 2500
    index = g.build_index()
    slots, dist = index.array("slots"), index.array("dist")
    sites = slots[slots["table_off"] != 0xFFFFFFFF]
    assert len(sites) > 2000 and int(sites["n"].max()) >= 4          # touching / multi-allelic sites merge into larger sites
    assert max(len(s) for s in g.node_seqs) <= 32
    # reads from the haplotypes map back exactly on the CPU restatement
    rs = synth.simulate_reads(g, 200, length=150, sub_rate=0.0, seed=2)
    res = H.oracle_map(index, rs.reads, rs.quals, threads=4)
    assert (res[0]["flags"] & 1).all() and (res[0]["score"] == 160).mean() > 0.97
    index.close()


@pytest.mark.gpu
def test_1000_genomes_slice_maps_like_the_oracle():
    g, _ = V.kgp_100k()
    index = g.build_index()
    dev = capi.Device(index)
    rs = synth.simulate_reads(g, 3000, length=150, sub_rate=0.01, seed=7)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    bad = H.compare_stage_dumps(dev.seed_stage(rbuf, qbuf, read_off), H.oracle_seed_stage(index, rs.reads, rs.quals), rs.n)
    assert not bad, f"stage: {len(bad)} reads differ; first {bad[0]}"
    bad = H.compare_alignments(H.gpu_map(dev, rs.reads, rs.quals), H.oracle_map(index, rs.reads, rs.quals, threads=8), rs.n)
    assert not bad, f"single-end: {len(bad)} of {rs.n} reads differ; first {bad[0]}"
    rp = synth.simulate_pairs(g, 1500, sub_rate=0.01, seed=8, indel_rate=0.002)
    rng = np.random.default_rng(4)
    for i in range(1, rp.n, 12):
        m = rng.random(rp.length) < 0.12
        rp.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    p = H.paired_params(); p.max_rescue_attempts = 15
    bad = H.compare_alignments(H.gpu_map(dev, rp.reads, rp.quals, p, paired=True), H.oracle_map_paired(index, rp.reads, rp.quals, p, threads=8), rp.n)
    assert not bad, f"paired: {len(bad)} of {rp.n} reads differ; first {bad[0]}"
    dev.close(); index.close()
