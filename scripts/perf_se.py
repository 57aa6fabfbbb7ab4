import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import helpers as H
from vg_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
t = time.time()
g = synth.make_variant_graph()   # config 2 graph: 1 Mbp, 1000 variants, 8 haplotypes
print("graph", time.time() - t, len(g.node_seqs)); t = time.time()
index = g.build_index()
print("index", time.time() - t, index.view.n_hits, index.view.table_cells); t = time.time()
rs = synth.simulate_pairs(g, n // 2, sub_rate=0.002, seed=22)
print("reads", time.time() - t); t = time.time()
dev = capi.Device(index)
rbuf, qbuf, read_off = H.pack_reads(rs.reads, rs.quals)
for rep in range(3):
    t = time.time()
    got = dev.map_arrays(rbuf, qbuf, read_off)
    wall = time.time() - t
    print(f"gpu rep {rep}: wall {wall:.3f}s kernels {dev.kernel_ms():.1f} ms -> {n / (dev.kernel_ms() / 1e3) / 1e6:.2f} M reads/s (kernel), {n / wall / 1e6:.2f} M reads/s (wall)")
print("mapped frac", (got[0]["flags"] & 1).mean(), "mapq60", (got[0]["mapq"] == 60).mean(), "status bad", (got[3] != 0).sum())
import os
nt = os.cpu_count()
sub = min(n, 400000)
t = time.time()
want = H.oracle_map(index, rs.reads[:sub], rs.quals[:sub], threads=nt)
dt = time.time() - t
print(f"oracle {sub} reads on {nt} threads: {dt:.2f}s -> {sub / dt / 1e6:.3f} M reads/s", want[4])
bad = H.compare_alignments([x[:sub * (len(x) // n)] if False else x for x in got], want, min(sub, 20000))
print("bad in first 20000:", len(bad))
