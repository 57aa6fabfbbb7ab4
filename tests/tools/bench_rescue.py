"""Secondary workload: config-2 pairs with a share of mates too noisy to seed, mate rescue on
(vg giraffe default --rescue-attempts 15).  Kernel-time throughput, CPU oracle rate, parity count.
usage: python tests/tools/bench_rescue.py [n_pairs] [wrecked_every]"""
import os, sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
import bench
from vg_b200 import capi, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
every = int(sys.argv[2]) if len(sys.argv) > 2 else 10
threads, note = bench.usable_cpus()
g = synth.make_variant_graph(); index = g.build_index()
rs = synth.simulate_pairs(g, n_pairs, sub_rate=0.002, seed=23)
rng = np.random.default_rng(2)
for i in range(1, rs.n, 2 * every):
    m = rng.random(rs.length) < 0.12
    rs.reads[i, m] = synth.BASES[rng.integers(0, 4, size=int(m.sum()))]
dev = capi.Device(index)
rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
out = {}
for attempts in (0, 15):
    p = H.paired_params(); p.max_rescue_attempts = attempts
    best = None
    for rep in range(3):
        got = dev.map_arrays(rbuf, qbuf, read_off, p, paired=True)
        ms = dev.kernel_ms(); best = ms if best is None else min(best, ms)
    sub = min(rs.n, 40000)
    t = time.time(); want = H.oracle_map_paired(index, rs.reads[:sub], rs.quals[:sub], p, threads=threads); dt = time.time() - t
    ok = np.nonzero(got[3][:sub] == 0)[0]
    bad = H.compare_alignments(got, want, sub, indices=ok.tolist())
    out[f"rescue_attempts_{attempts}"] = {"reads": rs.n, "wrecked_mate_every_n_pairs": every, "gpu_kernel_ms": best, "gpu_reads_per_s": rs.n / (best / 1e3),
        "stage_ms_last_chunk": dev.stage_times(), "status_errors": int((got[3] != 0).sum()), "mapped_fraction": float((got[0]["flags"] & 1).mean()),
        "rescued_records": int(((got[0]["flags"] & capi.GB_ALN_RESCUED) != 0).sum()), "cpu_reads_per_s": sub / dt, "cpu_threads": threads,
        "parity_mismatches": len(bad), "parity_compared": int(len(ok))}
    print(attempts, json.dumps(out[f"rescue_attempts_{attempts}"]), flush=True)
dev.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_rescue.json", "w"), indent=1)
