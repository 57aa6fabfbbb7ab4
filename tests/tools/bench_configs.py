"""Secondary workloads of BASELINE.json (not the bench line): config 4 (branchy graph, 150 bp SE)
and config 5 (config-2 graph, 250 bp SE with 5 % errors: every read takes the tail-DP path).
Prints kernel-time throughput, stage times, the CPU oracle rate on the same reads and a parity count.
usage: python tests/tools/bench_configs.py [n_reads]"""
import os, sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
import bench
from vg_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
only = sys.argv[2] if len(sys.argv) > 2 else ""
threads, note = bench.usable_cpus()
out = {}
for name, make_graph, kw in [
    ("config4_branchy_150bp_SE", lambda: synth.make_branchy_graph(), dict(length=150, sub_rate=0.005, seed=44)),
    ("config5_tailDP_250bp_SE", lambda: synth.make_variant_graph(), dict(length=250, sub_rate=0.03, ins_rate=0.01, del_rate=0.01, seed=55)),
    ("config2_graph_150bp_SE", lambda: synth.make_variant_graph(), dict(length=150, sub_rate=0.002, seed=23)),
]:
    if only and only not in name:
        continue
    g = make_graph(); index = g.build_index()
    rs = synth.simulate_reads(g, n, **kw)
    dev = capi.Device(index)
    rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
    best = None
    for rep in range(3):
        got = dev.map_arrays(rbuf, qbuf, read_off)
        ms = dev.kernel_ms()
        best = ms if best is None else min(best, ms)
    stages = dev.stage_times()
    kernels = dev.kernel_times()
    sub = min(n, 100000)
    t = time.time()
    want = H.oracle_map(index, rs.reads[:sub], rs.quals[:sub], threads=threads)
    dt = time.time() - t
    bad = H.compare_alignments(got, want, min(sub, 20000))
    out[name] = {"reads": n, "gpu_kernel_ms": best, "gpu_reads_per_s": n / (best / 1e3), "stage_ms_last_chunk": stages, "kernel_ms": [(k, round(v, 3)) for k, v in kernels], "plan": dev.plan_stats(),
                 "status_errors": int((got[3] != 0).sum()), "mapped_fraction": float((got[0]["flags"] & 1).mean()),
                 "cpu_reads_per_s": sub / dt, "cpu_threads": threads, "oracle_counters": {k: int(v) for k, v in want[4].items()},
                 "parity_mismatches_of_20000": len(bad)}
    print(name, json.dumps(out[name]), flush=True)
    dev.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/bench_configs.json", "w"), indent=1)
