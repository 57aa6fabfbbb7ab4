"""FragmentLengthDistribution (mapper.hpp:83-139, mapper.cpp:5231-5333) and the training branch of
MinimizerMapper::map_paired(aln1, aln2, ambiguous_pair_buffer) (minimizer_mapper.cpp:1303-1395) driven as
giraffe_main.cpp:2246-2400 drives it.  The oracle is pinned to the reference's own fixture
(unittest/minimizer_mapper.cpp:37-108 -> tests/golden/fragment_lengths.json, scripts/extract_fragment_vectors.py)
and to independent numpy / scipy restatements; the library's host-side gb_fragment_* must agree with it
exactly; gb_map_paired_job is compared with the oracle job on the GPU."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import helpers as H
from vg_b200 import capi, synth

GOLDEN = json.loads((Path(__file__).parent / "golden" / "fragment_lengths.json").read_text())


def _numpy_estimate(sample, fraction):
    """estimate_distribution restated independently: trimmed moments + truncated-normal correction."""
    from scipy.stats import norm
    x = np.sort(np.asarray(sample, dtype=np.float64))
    skip = int(len(x) * (1.0 - fraction) * 0.5)
    t = x[skip:len(x) - skip]
    mu = t.sum() / len(t)
    raw_var = (t * t).sum() / len(t) - mu * mu
    a = norm.ppf(1.0 - 0.5 * (1.0 - fraction))
    return mu, np.sqrt(raw_var / (1.0 - 2.0 * a * norm.pdf(a)))


def test_phi_inv_matches_scipy():
    from scipy.stats import norm
    lib = H.oracle_lib()
    lib.oracle_phi_inv.argtypes = [C.c_double]; lib.oracle_phi_inv.restype = C.c_double
    for p in [1e-300, 1e-20, 1e-12, 1e-6, 0.01, 0.0749, 0.075, 0.0751, 0.3, 0.5, 0.7, 0.925, 0.9251, 0.975, 0.999999, 1 - 1e-12]:
        want = norm.ppf(p)
        assert abs(lib.oracle_phi_inv(p) - want) <= 1e-12 * max(1.0, abs(want)), p


def test_reference_fixture_gives_a_reasonable_distribution():
    """unittest/minimizer_mapper.cpp:96-107: register every distance <= max_fragment_length; REQUIRE(std_dev() <= 400)."""
    cfg = GOLDEN["distribution"]
    kept = [d for d in GOLDEN["distances"] if d <= GOLDEN["max_fragment_length"]]
    mean, sd, finalized, n = H.oracle_fragment_estimate(kept, cfg["maximum_sample_size"], cfg["reestimation_frequency"], cfg["robust_estimation_fraction"])
    assert finalized and n == cfg["maximum_sample_size"]          # 1056 eligible distances, the first 1000 are used
    assert sd <= GOLDEN["require"]["std_dev_at_most"]
    want_mean, want_sd = _numpy_estimate(kept[:cfg["maximum_sample_size"]], cfg["robust_estimation_fraction"])
    assert abs(mean - want_mean) < 1e-9 and abs(sd - want_sd) < 1e-9
    # the outlier tail (up to 654834) must not have moved it
    assert 150 < mean < 400


@pytest.mark.parametrize("max_n,freq,frac", [(1000, 1000, 0.95), (200, 50, 0.9), (64, 7, 0.5), (5000, 100, 0.99)])
def test_library_distribution_equals_oracle(max_n, freq, frac):
    """gb_fragment_* (host-side, no device needed) vs the oracle, state after every prefix that changes it."""
    rng = np.random.default_rng(max_n)
    sample = np.concatenate([rng.normal(420, 60, size=1500).astype(np.int64), rng.integers(2000, 10**6, size=40)])
    rng.shuffle(sample)
    f = capi.FragmentDistribution(max_n, freq, frac)
    assert not f.is_finalized() and f.mean() == 0.0 and f.std_dev() == 1.0          # mapper.hpp:135-136
    for i, v in enumerate(sample.tolist(), 1):
        f.register_fragment_length(v)
        if i % freq == 0 or i == max_n or i == len(sample):
            mean, sd, fin, n = H.oracle_fragment_estimate(sample[:i], max_n, freq, frac)
            assert (f.mean(), f.std_dev(), f.is_finalized(), f.curr_sample_size()) == (mean, sd, fin, n), i
    assert f.is_finalized() == (len(sample) >= max_n)
    f.close()


def test_forced_and_early_finalized_distributions():
    f = capi.FragmentDistribution()
    f.register_fragment_length(300); f.register_fragment_length(500)
    f.finalize()                                    # finalize_fragment_length_distr: keeps the running estimate (none yet: 0 +- 1)
    assert f.is_finalized() and (f.mean(), f.std_dev()) == (0.0, 1.0)
    f.register_fragment_length(700)
    assert f.curr_sample_size() == 2               # registrations after finalization are ignored (mapper.cpp:5259)
    g = capi.FragmentDistribution()
    g.force_parameters(400.0, 50.0)
    assert g.is_finalized() and (g.mean(), g.std_dev()) == (400.0, 50.0)
    with pytest.raises(capi.GbError):
        capi.FragmentDistribution(10, 0, 0.95)
    f.close(); g.close()


def _job_inputs(n_pairs, seed):
    g = synth.make_variant_graph(length=150000, n_snp=240, n_ins=30, n_del=30, n_haps=8, seed=12)
    rs = synth.simulate_pairs(g, n_pairs, frag_mean=380, frag_sd=45, sub_rate=0.004, seed=seed)
    rng = np.random.default_rng(seed)
    for i in rng.integers(0, rs.n, size=n_pairs // 8):          # noisy mates: not "perfect", so their pairs are buffered
        m = rng.random(rs.length) < 0.06
        rs.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
    for i in rng.integers(0, rs.n, size=n_pairs // 40):         # garbage mates
        rs.reads[i] = synth.BASES[rng.integers(0, 4, size=rs.length)]
    return g, rs


def test_oracle_job_routes_and_records_are_consistent_with_its_parts():
    """The job restatement against its own building blocks: training pairs are exactly the first `max_n` pairs (input
    order) whose ends both map uniquely and near-perfectly, their records are the single-end results; every other pair
    carries the paired result under the learned distribution; pairs seen while training that did not qualify are buffered."""
    n_pairs, max_n = 700, 200
    g, rs = _job_inputs(n_pairs, seed=77)
    index = g.build_index()
    p = H.default_map_params(); p.fragment_mean = 0.0; p.fragment_stdev = 0.0
    aln, maps, edits, status, route, (mean, sd, n) = H.oracle_map_paired_job(index, rs.reads, rs.quals, p, max_n, 50, 0.95, threads=8)
    assert n == max_n and (route == capi.GB_PAIR_TRAINING).sum() == max_n
    last_training = int(np.nonzero(route == capi.GB_PAIR_TRAINING)[0][-1])
    assert (route[last_training + 1:] == capi.GB_PAIR_PAIRED).all()                 # nothing is trained on or buffered after finalization
    assert set(np.unique(route[:last_training + 1]).tolist()) <= {capi.GB_PAIR_TRAINING, capi.GB_PAIR_BUFFERED}
    assert (route[:last_training + 1] == capi.GB_PAIR_BUFFERED).sum() >= 10
    se = H.oracle_map(index, rs.reads, rs.quals, threads=8)
    p2 = H.default_map_params(); p2.fragment_mean = mean; p2.fragment_stdev = sd
    pe = H.oracle_map_paired(index, rs.reads, rs.quals, p2, threads=8)
    got = (aln, maps, edits, status)
    train_reads = np.nonzero(np.repeat(route == capi.GB_PAIR_TRAINING, 2))[0].tolist()
    other_reads = np.nonzero(np.repeat(route != capi.GB_PAIR_TRAINING, 2))[0].tolist()
    assert not H.compare_alignments(got, se, rs.n, mapq_tol=0, indices=train_reads)
    assert not H.compare_alignments(got, pe, rs.n, mapq_tol=0, indices=other_reads)
    # the qualifying rule itself (minimizer_mapper.cpp:1316-1322): MAPQ 60 and score >= 0.85 * match * L on both ends
    for pair in range(last_training + 1):
        perfect = all((se[0][2 * pair + r]["flags"] & 1) and se[0][2 * pair + r]["mapq"] == 60 and se[0][2 * pair + r]["score"] >= 0.85 * rs.length
                      for r in range(2))
        if route[pair] == capi.GB_PAIR_TRAINING:
            assert perfect
    assert abs(mean - 380) < 20 and abs(sd - 45) < 15


@pytest.mark.gpu
@pytest.mark.parametrize("n_pairs,max_n,freq,window", [(2500, 1000, 1000, 0), (1500, 300, 100, 128), (400, 1000, 1000, 64)])
def test_paired_job_with_fragment_length_training(n_pairs, max_n, freq, window):
    """Same routes, same estimated distribution, same records as the oracle job; the third case ends before the
    sample fills up (forced finalization with the initial 0 +- 1, as the reference does)."""
    g, rs = _job_inputs(n_pairs, seed=70 + n_pairs)
    index = g.build_index()
    p = H.default_map_params()
    p.fragment_mean = 0.0; p.fragment_stdev = 0.0               # not forced: learn it
    want = H.oracle_map_paired_job(index, rs.reads, rs.quals, p, max_n, freq, 0.95, threads=8)
    dev = capi.Device(index)
    f = capi.FragmentDistribution(max_n, freq, 0.95)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    got = dev.map_paired_job(rbuf, qbuf, read_off, f, p, training_window=window)
    assert (got[4] == want[4]).all(), "pair routes differ"
    assert (f.mean(), f.std_dev(), f.curr_sample_size()) == want[5]
    assert f.is_finalized()
    if n_pairs > 1000:
        assert (got[4] == capi.GB_PAIR_TRAINING).sum() == max_n and abs(f.mean() - 380) < 15 and abs(f.std_dev() - 45) < 10
    bad = H.compare_alignments(got, want, rs.n)
    assert not bad, f"{len(bad)} of {rs.n} reads differ; first: read {bad[0][0]}\n got={bad[0][1]}\nwant={bad[0][2]}"
    # every record of the job is mate-linked: training pairs go through pair_all too (minimizer_mapper.cpp:1345-1350)
    assert ((got[0]["flags"] & capi.GB_ALN_PAIRED) != 0).all()
    # a second call with the now finalized distribution maps everything paired
    got2 = dev.map_paired_job(rbuf, qbuf, read_off, f, p)
    assert (got2[4] == capi.GB_PAIR_PAIRED).all()
    p2 = H.default_map_params(); p2.fragment_mean = f.mean(); p2.fragment_stdev = f.std_dev()
    direct = H.gpu_map(dev, rs.reads, rs.quals, p2, paired=True)
    assert not H.compare_alignments(got2, direct, rs.n, mapq_tol=0)
    dev.close(); f.close()
