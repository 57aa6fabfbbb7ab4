"""Randomised GPU-vs-oracle soak (not part of the pytest suite: minutes of CPU time): many graph seeds x error models x
single-end / paired-end (with rescue) x max_multimaps, every record compared.  usage: python tests/tools/soak_parity.py [reads] [rounds]"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
from vg_b200 import capi, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
total = bad_total = 0
t0 = time.time()
for rnd in range(rounds):
    graphs = [
        ("variants", synth.make_variant_graph(length=150000, n_snp=400, n_ins=50, n_del=50, n_haps=8, seed=100 + rnd)),
        ("repeats", synth.make_variant_graph(length=40000, n_snp=60, n_ins=8, n_del=8, n_haps=4, seed=200 + rnd, repeat_unit=500 + 50 * rnd, repeat_copies=5)),
        ("nested", synth.make_nested_graph(n_items=400, n_haps=8, seed=300 + rnd)),
        ("branchy", synth.make_branchy_graph(n_layers=2500, n_haps=16, seed=400 + rnd)),
    ]
    for name, g in graphs:
        if g is None:
            continue
        index = g.build_index()
        dev = capi.Device(index, 0)
        for mode in ("se", "pe"):
            for k in (1, 3):
                for sub, indel in ((0.002, 0.0002), (0.02, 0.004)):
                    seed = 1000 * rnd + 17 * k + int(sub * 1e4)
                    if mode == "se":
                        rs = synth.simulate_reads(g, n_reads, length=150, sub_rate=sub, ins_rate=indel, del_rate=indel, seed=seed)
                        p = H.default_map_params(); p.max_multimaps = k
                        got = H.gpu_map(dev, rs.reads, rs.quals, p)
                        want = H.oracle_map(index, rs.reads, rs.quals, p, threads=16)
                    else:
                        rs = synth.simulate_pairs(g, n_reads // 2, sub_rate=sub, seed=seed, indel_rate=indel)
                        p = H.paired_params(); p.max_rescue_attempts = 15; p.max_multimaps = k
                        got = H.gpu_map(dev, rs.reads, rs.quals, p, paired=True)
                        want = H.oracle_map_paired(index, rs.reads, rs.quals, p, threads=16)
                    st_bad = int((got[3] != 0).sum())
                    ok = np.flatnonzero(got[3] == 0)
                    idx = [j * rs.n + int(i) for j in range(k) for i in ok]
                    bad = H.compare_alignments(got, want, rs.n, indices=idx, k=k)
                    total += len(idx); bad_total += len(bad)
                    print(f"round {rnd} {name:9s} {mode} k={k} sub={sub}: {len(idx)} records, {len(bad)} differ, {st_bad} status != 0" + (f"; first {bad[0]}" if bad else ""), flush=True)
        dev.close(); index.close()
print(f"SOAK: {total} records compared, {bad_total} differ, {time.time() - t0:.0f}s")
